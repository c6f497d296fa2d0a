#!/usr/bin/env python
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers as H
for kind in ("mx", "nv"):
    for m in (1, 7, 16, 17, 32, 33, 64, 65, 128):
        for (n, k) in ((504, 4096), (1000, 2176)):
            aq, asf = H.random_fp4_operand(m, k, kind, seed=m, sf_mode="narrow")
            bq, bsf = H.random_fp4_operand(n, k, kind, seed=n, sf_mode="narrow")
            want = H.gemm_oracle_bits(aq, asf, bq, bsf, kind)
            got = H.run_gemm(aq, asf, bq, bsf, kind, 1.0, cfg=(1, 128))
            mism, rel = H.compare_bits(got, want)
            got0 = H.run_gemm(aq, asf, bq, bsf, kind, 1.0, cfg=(0, 0))
            m0, r0 = H.compare_bits(got0, want)
            print(json.dumps(dict(kind=kind, m=m, n=n, k=k, mismatch=mism, rel=rel, auto_mismatch=m0)), flush=True)
