#!/usr/bin/env python
"""Per-case mismatch of OUR quantisers against the reference's own kernels (oracle/_ref/qutlass_ref_C.so) in ONE process:
the reference library is loaded first, so its `_qutlass_C` ops win the namespace and qutlass_b200's own registration is
skipped (its Python functions call the C-ABI directly and do not need it).  Diagnostic tool, prints one JSON line per case."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from oracle import ref_gpu
torch.ops.load_library(ref_gpu.LIB)
ops = torch.ops._qutlass_C
import oracle as O
import helpers as H
import qutlass_b200 as Q

k = 4096
CASES = [("mx", "abs_max", 32, 1.0, 512), ("mx", "abs_max", 128, 1.0, 512), ("mx", "quest", 64, 1.0, 512),
         ("mx", "quest", 128, 1.0, 512), ("mx", "abs_max", 128, 1.0, 2048), ("nv", "abs_max", 16, 6.0, 512),
         ("nv", "abs_max", 128, 6.0, 512), ("nv", "quest", 64, 6.0, 512)]
for i, (fmt, method, had, gs, rows) in enumerate(CASES):
    try:
        x = H.bf16_tensor_from_f32(H.random_bf16((rows, k), seed=1000 + had + i))
        R = H.bf16_tensor_from_f32(O.hadamard_matrix(had))
        gst = torch.tensor([gs], device="cuda")
        if fmt == "mx":
            q, sf = Q.fusedQuantizeMx(x, R, method=method)
        else:
            q, sf = Q.fusedQuantizeNv(x, R, gst, method=method)
        rq, rsf = ref_gpu._quantize(torch, ops, fmt, method, x, R, gst)
        torch.cuda.synchronize()
        cols = k // (32 if fmt == "mx" else 16)
        sf_o = H.u8_of(sf).reshape(-1, sf.shape[-1])[:rows, :cols]
        sf_r = H.u8_of(rsf)[:rows, :cols]
        dq = O.dequant_mx if fmt == "mx" else O.dequant_nv
        a, b = dq(H.u8_of(q), sf_o), dq(H.u8_of(rq).reshape(rows, -1), sf_r)
        print(json.dumps({"case": [fmt, method, had, gs, rows], "dq_mismatch": float((a != b).mean()),
                          "sf_mismatch": float((sf_o != sf_r).mean()),
                          "code_mismatch": float((H.u8_of(q) != H.u8_of(rq).reshape(rows, -1)).mean())}), flush=True)
    except Exception as e:   # noqa: BLE001
        print(json.dumps({"case": [fmt, method, had, gs, rows], "error": f"{type(e).__name__}: {e}"[:300]}), flush=True)
