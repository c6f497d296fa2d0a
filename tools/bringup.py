#!/usr/bin/env python
"""GPU bring-up ladder: runs each check in its own subprocess (a trapped kernel kills only that
process) and appends one JSON line per check to gpurun_out/bringup.jsonl.

  python tools/bringup.py            # whole ladder
  python tools/bringup.py --one gemm mx 1 128 128 128 256 narrow   # a single GEMM check (internal)
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "gpurun_out")


def one_gemm(kind, cg, bn, m, n, k, sf_mode):
    import numpy as np
    import helpers as H
    cg, bn, m, n, k = int(cg), int(bn), int(m), int(n), int(k)
    aq, asf = H.random_fp4_operand(m, k, kind, seed=1, sf_mode=sf_mode)
    bq, bsf = H.random_fp4_operand(n, k, kind, seed=2, sf_mode=sf_mode)
    if sf_mode == "ones":      # all values 1.0, scales 1.0 -> D = K
        aq[:] = 0x22
        bq[:] = 0x22
        asf[:] = 127 if kind == "mx" else 0x38
        bsf[:] = 127 if kind == "mx" else 0x38
    if sf_mode == "sfa":       # only A scales vary
        bsf[:] = 127 if kind == "mx" else 0x38
    if sf_mode == "sfb":
        asf[:] = 127 if kind == "mx" else 0x38
    want = H.gemm_oracle_bits(aq, asf, bq, bsf, kind, 1.0)
    got = H.run_gemm(aq, asf, bq, bsf, kind, 1.0, cfg=(cg, bn))
    mism, rel = H.compare_bits(got, want)
    res = dict(check="gemm", kind=kind, cg=cg, bn=bn, m=m, n=n, k=k, sf=sf_mode, mismatch=mism, max_rel=rel)
    if mism > 0:
        import oracle as O
        bad = np.argwhere(got != want)
        res["first_bad"] = [[int(r), int(c), float(O.bf16_from_bits(got[r, c])), float(O.bf16_from_bits(want[r, c]))]
                            for r, c in bad[:12]]
        res["bad_rows"] = int(len(set(bad[:, 0].tolist())))
        res["bad_cols"] = int(len(set(bad[:, 1].tolist())))
        # which 32-row / 32-col groups are wrong (layout hints)
        res["bad_row_groups"] = sorted(set((bad[:, 0] // 32).tolist()))[:16]
        res["bad_col_groups"] = sorted(set((bad[:, 1] // 32).tolist()))[:16]
    print("RESULT " + json.dumps(res), flush=True)


def one_quant(kind, method, had, rows, k):
    import numpy as np
    import torch
    import helpers as H
    import oracle as O
    import qutlass_b200 as Q
    had, rows, k = int(had), int(rows), int(k)
    x = H.random_bf16((rows, k), seed=3)
    R = O.hadamard_matrix(had)
    xt = H.bf16_tensor_from_f32(x)
    Rt = H.bf16_tensor_from_f32(R)
    res = dict(check="quant", kind=kind, method=method, had=had, rows=rows, k=k)
    if kind == "mx":
        ref = O.quantize_mx(x, R, method, arithmetic="kernel")
        out = Q.fusedQuantizeMx(xt, Rt, method=method, return_mask=(method == "quest"))
        q, sf = out[0], out[1]
        torch.cuda.synchronize()
        cols = k // 32
        sf_rm = H.u8_of(sf).reshape(-1)[:rows * cols].reshape(rows, cols)  # kernel writes the flat group stream
        dq = O.dequant_mx(H.u8_of(q), sf_rm)
        dq_ref = O.dequant_mx(ref["q"].reshape(rows, k // 2), ref["sf"].reshape(rows, cols))
        res["sf_mismatch"] = float((sf_rm != ref["sf"].reshape(rows, cols)).mean())
        if method == "quest":
            res["mask_mismatch"] = float((H.u8_of(out[2]).reshape(-1).view(np.uint32) != ref["mask"]).mean())
        pr, pc = O.padded_sf_shape(rows, cols)
        want_blk = H.blocked_sf(ref["sf"].reshape(rows, cols))
        got_blk = H.u8_of(Q.to_blocked(sf))
        # compare blocked only where sf matches (positions are what we test)
        res["blocked_mismatch"] = float((got_blk != H.blocked_sf(sf_rm)).mean())
    else:
        gs = 6.0
        ref = O.quantize_nv(x, R, gs, method, arithmetic="kernel")
        gst = torch.tensor([gs], dtype=torch.float32, device="cuda")
        q, sf = Q.fusedQuantizeNv(xt, Rt, gst, method=method)
        torch.cuda.synchronize()
        cols = k // 16
        sf_rm = H.u8_of(sf).reshape(-1)[:rows * cols].reshape(rows, cols)  # kernel writes the flat group stream
        dq = O.dequant_nv(H.u8_of(q), sf_rm)
        dq_ref = O.dequant_nv(ref["q"].reshape(rows, k // 2), ref["sf"].reshape(rows, cols))
        res["sf_mismatch"] = float((sf_rm != ref["sf"].reshape(rows, cols)).mean())
        got_blk = H.u8_of(Q.to_blocked(sf))
        res["blocked_mismatch"] = float((got_blk != H.blocked_sf(sf_rm)).mean())
    res["dq_mismatch"] = float((dq != dq_ref).mean())
    # generic (non-Hadamard) rotation: identity
    print("RESULT " + json.dumps(res), flush=True)


def one_quant_generic(kind, had, rows, k):
    import numpy as np
    import torch
    import helpers as H
    import oracle as O
    import qutlass_b200 as Q
    had, rows, k = int(had), int(rows), int(k)
    x = H.random_bf16((rows, k), seed=4)
    rng = np.random.default_rng(5)
    R = O.bf16_round(rng.standard_normal((had, had)).astype(np.float32) * had ** -0.5)
    xt, Rt = H.bf16_tensor_from_f32(x), H.bf16_tensor_from_f32(R)
    if kind == "mx":
        ref = O.quantize_mx(x, R, "abs_max", arithmetic="kernel")
        q, sf = Q.fusedQuantizeMx(xt, Rt, method="abs_max")
        torch.cuda.synchronize()
        cols = k // 32
        dq = O.dequant_mx(H.u8_of(q), H.u8_of(sf).reshape(-1)[:rows * cols].reshape(rows, cols))
        dq_ref = O.dequant_mx(ref["q"].reshape(rows, k // 2), ref["sf"].reshape(rows, cols))
    else:
        ref = O.quantize_nv(x, R, 1.0, "abs_max", arithmetic="kernel")
        gst = torch.tensor([1.0], dtype=torch.float32, device="cuda")
        q, sf = Q.fusedQuantizeNv(xt, Rt, gst)
        torch.cuda.synchronize()
        cols = k // 16
        dq = O.dequant_nv(H.u8_of(q), H.u8_of(sf).reshape(-1)[:rows * cols].reshape(rows, cols))
        dq_ref = O.dequant_nv(ref["q"].reshape(rows, k // 2), ref["sf"].reshape(rows, cols))
    print("RESULT " + json.dumps(dict(check="quant_generic", kind=kind, had=had, rows=rows, k=k,
                                      dq_mismatch=float((dq != dq_ref).mean()))), flush=True)


GEMM_CASES = lambda cg, bn: ((128 * cg, max(bn, 128), 256, "ones"), (128 * cg, max(bn, 128), 256, "one"),
                             (128 * cg, max(bn, 128), 1024, "one"), (256, 512, 1024, "narrow"),
                             (504, 504, 2048, "narrow"), (1, 504, 4096, "narrow"), (300, 1000, 2176, "narrow"),
                             (130, 100, 96, "narrow"), (1000, 1336, 512, "narrow"))


def gemm_cfg(kind, cg, bn):
    cg, bn = int(cg), int(bn)
    for (m, n, k, sf) in GEMM_CASES(cg, bn):
        one_gemm(kind, cg, bn, m, n, k, sf)


def quant_all():
    for kind, hads in (("mx", (32, 64, 128)), ("nv", (16, 32, 64, 128))):
        for method in ("abs_max", "quest"):
            for had in hads:
                one_quant(kind, method, had, 256, 1024)
    one_quant("mx", "abs_max", 32, 1, 4096)
    one_quant("mx", "quest", 128, 300, 4096)
    one_quant("nv", "abs_max", 16, 77, 96)
    one_quant("mx", "abs_max", 64, 5, 96 * 2)
    one_quant_generic("mx", 32, 64, 512)
    one_quant_generic("mx", 128, 64, 512)
    one_quant_generic("nv", 16, 64, 512)


def run_sub(args, timeout=300):
    t0 = time.time()
    recs = []
    try:
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", *map(str, args)],
                           capture_output=True, text=True, timeout=timeout)
        stdout, stderr, rc = r.stdout, r.stderr, r.returncode
    except subprocess.TimeoutExpired as e:
        stdout = e.stdout.decode(errors="replace") if isinstance(e.stdout, bytes) else (e.stdout or "")
        stderr = e.stderr.decode(errors="replace") if isinstance(e.stderr, bytes) else (e.stderr or "")
        rc = "timeout"
    for l in stdout.splitlines():
        if l.startswith("RESULT "):
            recs.append(json.loads(l[7:]))
    if rc != 0:
        recs.append(dict(check="process", args=list(map(str, args)), rc=rc, stdout=stdout[-1500:], stderr=stderr[-3000:]))
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, "bringup.jsonl"), "a") as f:
        for rec in recs:
            f.write(json.dumps(rec) + "\n")
    for rec in recs:
        short = {k: v for k, v in rec.items() if k not in ("stdout", "stderr", "first_bad")}
        print(json.dumps(short), flush=True)
        if "stderr" in rec:
            print(rec["stderr"][-1500:], flush=True)
    print(f"# {args} took {time.time() - t0:.1f}s", flush=True)
    return recs


def ladder(which):
    if "quant" in which:
        run_sub(["quantall"])
    if "gemm" in which:
        for kind in ("mx", "nv"):
            for (cg, bn) in ((1, 128), (1, 64), (1, 192), (1, 256), (2, 128), (2, 192), (2, 256)):
                run_sub(["gemmcfg", kind, cg, bn])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--one":
        what = sys.argv[2]
        if what == "gemm":
            one_gemm(*sys.argv[3:])
        elif what == "gemmcfg":
            gemm_cfg(*sys.argv[3:])
        elif what == "quantall":
            quant_all()
        elif what == "quant":
            one_quant(*sys.argv[3:])
        elif what == "quant_generic":
            one_quant_generic(*sys.argv[3:])
        sys.exit(0)
    which = sys.argv[1:] or ["quant", "gemm"]
    ladder(which)
