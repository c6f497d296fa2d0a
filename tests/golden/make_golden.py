#!/usr/bin/env python
"""Generate golden vectors by RUNNING THE REFERENCE'S OWN test helpers on CPU.

Runs only in the build container (needs /root/reference); the GPU box never
reads /root/reference -- it uses the committed ``*.npz`` files next to this script.

What is lifted (function bodies are exec'd from the reference sources, nothing is
copied into this repo):
  * /root/reference/tests/mxfp4_test.py : get_hadamard_matrix, _rtne_fp4, _dq_fp4,
    _unpack_mask, _forward_quantize_ref          (lines 39-184)
  * /root/reference/tests/nvfp4_test.py : the NV flavours of the same (lines 36-170)
  * /root/reference/qutlass/utils.py    : ceil_div, get_padded_shape_mx/_nv, to_blocked
    (torch path, lines 136-193)
  * /root/reference/tests/mxfp8_test.py : _pseudoquant_mxfp8 (lines 27-46, decorator stripped)

  * /root/reference/tests/quartet_test.py : _backward_quantize_ref, _backward_bf16_square_double_mxfp8,
    _mxfp4_transpose_mxfp8 (+ its _rtne_fp4 / _dq_fp4) and qutlass/utils.py pad_to_block  -> backward_vectors.npz

Usage:  python tests/golden/make_golden.py [forward|backward|all]
"""
import ast
import os
import sys

import numpy as np
import torch
from scipy.linalg import hadamard

REF = os.environ.get("QUTLASS_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def lift(path, names, extra=None):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "np": np, "hadamard": hadamard}
    ns.update(extra or {})
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            node.decorator_list = []      # e.g. @torch.compile on _pseudoquant_mxfp8: run it eagerly
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def bits(t):
    return t.contiguous().view(torch.int16).numpy().view(np.uint16).copy()


def u8(t):
    return t.contiguous().view(torch.uint8).numpy().copy()


def main():
    mx = lift(f"{REF}/tests/mxfp4_test.py",
              ["get_hadamard_matrix", "_rtne_fp4", "_dq_fp4", "_unpack_mask", "_forward_quantize_ref"])
    nv = lift(f"{REF}/tests/nvfp4_test.py",
              ["get_hadamard_matrix", "_rtne_fp4", "_dq_fp4", "_unpack_mask", "_forward_quantize_ref"])
    ut = lift(f"{REF}/qutlass/utils.py",
              ["ceil_div", "get_padded_shape_mx", "get_padded_shape_nv", "to_blocked"])
    dev = torch.device("cpu")
    out = {}

    # ---- e2m1 rounding: every tie point, neighbours, saturation, signed zeros
    pts = [0.0, -0.0, 0.25, 0.75, 1.25, 1.75, 2.5, 3.5, 5.0, 6.0, 7.0, 100.0, 1e-30]
    v = []
    for p in pts:
        for d in (0.0, 1e-6, -1e-6, 1e-3, -1e-3):
            v += [p + d, -(p + d)]
    v += list(np.linspace(-7, 7, 1121))
    xv = torch.tensor(v, dtype=torch.float64)
    if xv.numel() % 2:
        xv = torch.cat([xv, torch.zeros(1, dtype=torch.float64)])
    y, packed = mx["_rtne_fp4"](xv)
    out["rtne_in"] = xv.numpy()
    out["rtne_val"] = y.numpy()
    out["rtne_packed"] = packed.numpy()

    # ---- MX quantise reference (float64 emulation) on seeded data
    torch.manual_seed(0)
    x = torch.randn(64, 1024, dtype=torch.bfloat16) * 25.0
    # sprinkle special groups: all-zero group, single spike, constant group
    x[0, :32] = 0
    x[1, :32] = 0
    x[1, 7] = 100.0
    x[2, :128] = 3.0
    out["mx_x_bits"] = bits(x)
    for h in (32, 64, 128):
        H = mx["get_hadamard_matrix"](h, torch.bfloat16, dev)
        out[f"had{h}_bits"] = bits(H)
        for quest in (True, False):
            xh_dq, mask_unpacked, (e2m1, e8m0, mask) = mx["_forward_quantize_ref"](x, H, h, quest=quest)
            tag = f"mx_h{h}_{'quest' if quest else 'absmax'}"
            out[tag + "_e2m1"] = e2m1.numpy()
            out[tag + "_e8m0"] = u8(e8m0)
            out[tag + "_mask"] = mask.numpy()
            out[tag + "_dq"] = xh_dq.numpy()

    # ---- NV quantise reference
    H16 = nv["get_hadamard_matrix"](16, torch.bfloat16, dev)
    out["had16_bits"] = bits(H16)
    for h in (16, 32, 64, 128):
        H = nv["get_hadamard_matrix"](h, torch.bfloat16, dev)
        xh_dq, _, (e2m1, e4m3, mask) = nv["_forward_quantize_ref"](x, H, h)
        tag = f"nv_h{h}"
        out[tag + "_e2m1"] = e2m1.numpy()
        out[tag + "_e4m3"] = u8(e4m3)
        out[tag + "_dq"] = xh_dq.numpy()

    # ---- to_blocked
    g = torch.Generator().manual_seed(1)
    for (r, c) in ((128, 4), (256, 8), (384, 12), (128, 128)):
        sf = torch.randint(0, 255, (r, c), dtype=torch.uint8, generator=g)
        out[f"blk_{r}x{c}_in"] = sf.numpy()
        out[f"blk_{r}x{c}_out"] = ut["to_blocked"](sf).numpy()
    a = torch.empty(3, 200, 4096)
    out["padded_mx_3x200x4096"] = np.array(ut["get_padded_shape_mx"](a))
    out["padded_nv_3x200x4096"] = np.array(ut["get_padded_shape_nv"](a))
    a = torch.empty(1, 96)
    out["padded_mx_1x96"] = np.array(ut["get_padded_shape_mx"](a))
    out["padded_nv_1x96"] = np.array(ut["get_padded_shape_nv"](a))

    # ---- config 0: M=N=K=256 MXFP4 abs_max, emulated quantise + matmul on CPU
    torch.manual_seed(0)
    m = n = k = 256
    H = mx["get_hadamard_matrix"](32, torch.bfloat16, dev)
    a = torch.randn(m, k, dtype=torch.bfloat16) * 25.0
    b = torch.randn(n, k, dtype=torch.bfloat16) * 25.0
    _, _, (a_q, a_s, _) = mx["_forward_quantize_ref"](a, H, 32, quest=False)
    _, _, (b_q, b_s, _) = mx["_forward_quantize_ref"](b, H, 32, quest=False)
    a_dq, *_ = mx["_dq_fp4"](a_q, a_s, alpha=1.0)
    b_dq, *_ = mx["_dq_fp4"](b_q, b_s, alpha=1.0)
    ref64 = (a_dq @ b_dq.transpose(-2, -1))
    ref32 = (a_dq.float() @ b_dq.float().transpose(-2, -1))
    out["c0_a_bits"] = bits(a)
    out["c0_b_bits"] = bits(b)
    out["c0_a_q"] = a_q.numpy(); out["c0_a_s"] = u8(a_s)
    out["c0_b_q"] = b_q.numpy(); out["c0_b_s"] = u8(b_s)
    out["c0_out_bits"] = bits(ref64.to(torch.bfloat16))
    out["c0_out32_bits"] = bits(ref32.to(torch.bfloat16))

    # ---- NV GEMM golden (small): m=48, n=80, k=256, global_scale 6
    torch.manual_seed(1)
    m, n, k = 48, 80, 256
    a = torch.randn(m, k, dtype=torch.bfloat16) * 25.0
    b = torch.randn(n, k, dtype=torch.bfloat16) * 25.0
    H = nv["get_hadamard_matrix"](16, torch.bfloat16, dev)
    _, _, (a_q, a_s, _) = nv["_forward_quantize_ref"](a, H, 16)
    _, _, (b_q, b_s, _) = nv["_forward_quantize_ref"](b, H, 16)
    a_dq, *_ = nv["_dq_fp4"](a_q, a_s, alpha=1.0)
    b_dq, *_ = nv["_dq_fp4"](b_q, b_s, alpha=1.0)
    out["nvg_a_bits"] = bits(a); out["nvg_b_bits"] = bits(b)
    out["nvg_a_q"] = a_q.numpy(); out["nvg_a_s"] = u8(a_s)
    out["nvg_b_q"] = b_q.numpy(); out["nvg_b_s"] = u8(b_s)
    out["nvg_out_bits"] = bits((a_dq @ b_dq.transpose(-2, -1)).to(torch.bfloat16))

    # ---- MXFP8 ("next" row): the reference's pseudo-quantiser + GEMM reference (tests/mxfp8_test.py:27-78)
    f8 = lift(f"{REF}/tests/mxfp8_test.py", ["_pseudoquant_mxfp8"])
    torch.manual_seed(2)
    m, n, k = 24, 40, 256
    a = torch.rand(m, k, dtype=torch.bfloat16) * 25.0
    b = torch.rand(n, k, dtype=torch.bfloat16) * 25.0
    a[0, :32] = 0
    a_dq, (a_q, a_s) = f8["_pseudoquant_mxfp8"](a)
    b_dq, (b_q, b_s) = f8["_pseudoquant_mxfp8"](b)
    out["f8_a_bits"] = bits(a); out["f8_b_bits"] = bits(b)
    out["f8_a_q"] = u8(a_q); out["f8_a_s"] = u8(a_s)
    out["f8_b_q"] = u8(b_q); out["f8_b_s"] = u8(b_s)
    out["f8_out_bits"] = bits((a_dq @ b_dq.transpose(-2, -1)).to(torch.bfloat16))
    out["f8_out64_bits"] = bits((a_dq.double() @ b_dq.double().transpose(-2, -1)).to(torch.bfloat16))

    path = os.path.join(OUT, "reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


def main_backward():
    """Golden vectors of the backward re-quantisers from the reference's own test oracles (tests/quartet_test.py)."""
    ut = lift(f"{REF}/qutlass/utils.py", ["pad_to_block"])
    qt = lift(f"{REF}/tests/quartet_test.py",
              ["get_hadamard_matrix", "_rtne_fp4", "_dq_fp4", "_backward_quantize_ref",
               "_backward_bf16_square_double_mxfp8", "_mxfp4_transpose_mxfp8"], extra={"pad_to_block": ut["pad_to_block"]})
    dev = torch.device("cpu")
    out = {}
    H = qt["get_hadamard_matrix"](32, torch.bfloat16, dev)
    out["had32_bits"] = bits(H)

    # ---- backward_t_bf16: x [B, N, M] -> quantise x^T along N (tests/quartet_test.py:220-226)
    torch.manual_seed(3)
    x = torch.randn(2, 64, 96, dtype=torch.bfloat16) * 25.0
    x[0, :32, 5] = 0            # an all-zero group of the transposed matrix
    x[1, 32:, 7] = 0
    x[1, 40, 7] = 1.0           # single spike: amax exactly a power of two
    xh_dq, (e2m1, e8m0) = qt["_backward_quantize_ref"](x.transpose(-2, -1), H)
    out["t_x_bits"] = bits(x)
    out["t_e2m1"] = e2m1.numpy()
    out["t_e8m0"] = u8(e8m0)
    out["t_dq"] = xh_dq.numpy()

    # ---- backward_qt_bf16: MXFP4 input (abs_max forward quantisation emulated by the same oracle), alpha = 3
    #      (tests/quartet_test.py:228-239)
    _, (xq, xs) = qt["_backward_quantize_ref"](x, H)           # [2, 64, 48] bytes, [2, 64, 3] scales
    x_dq = qt["_dq_fp4"](xq, xs, alpha=3.0)[0]
    xh_dq, (e2m1, e8m0) = qt["_backward_quantize_ref"](x_dq.transpose(-2, -1), H)
    out["qt_in_e2m1"] = xq.numpy()
    out["qt_in_e8m0"] = u8(xs)
    out["qt_e2m1"] = e2m1.numpy()
    out["qt_e8m0"] = u8(e8m0)
    out["qt_dq"] = xh_dq.numpy()

    # ---- backward_bf16_square_double_mxfp8 (tests/quartet_test.py:264-291,369-378); every non-zero tile has amax >= 1
    torch.manual_seed(4)
    xs_ = torch.randn(200, 128, dtype=torch.bfloat16) * 25.0
    xs_[32:64, 32:64] = 0
    xs_[64:96, 0:32] = 1.0
    x_fp8, row_s, col_s = qt["_backward_bf16_square_double_mxfp8"](xs_)
    out["sq_x_bits"] = bits(xs_)
    out["sq_fp8"] = u8(x_fp8)
    out["sq_row"] = u8(row_s)
    out["sq_col"] = u8(col_s)
    # the reference test's own input: arange(0, 256) repeated over 2694 rows (only the first 256 rows + pad kept small)
    xa = torch.arange(0, 256, dtype=torch.bfloat16)[None, :].repeat(130, 1)
    x_fp8, row_s, col_s = qt["_backward_bf16_square_double_mxfp8"](xa)
    out["sqa_fp8"] = u8(x_fp8)
    out["sqa_row"] = u8(row_s)
    out["sqa_col"] = u8(col_s)

    # ---- mxfp4_transpose_mxfp8 (tests/quartet_test.py:294-345,380-385): m = 200 -> padded to 256, n = 128
    g = torch.Generator().manual_seed(5)
    m, n = 200, 128
    fp4 = torch.randint(0, 256, (m, n // 2), dtype=torch.uint8, generator=g)
    fp4[10:42, 3] = 0                                  # an all-zero column pair over one 32-row group? (rows 10..41 straddle)
    fp4[32:64, 8] = 0                                  # exactly one group of two output rows
    sc = torch.randint(120, 134, (256, n // 32), dtype=torch.uint8, generator=g)
    sc[m:] = 127                                       # the reference wrapper sets the pad rows' scales to 1.0
    xq8, exps = qt["_mxfp4_transpose_mxfp8"](fp4, sc)
    out["tr_fp4"] = fp4.numpy()
    out["tr_scales"] = sc.numpy()
    out["tr_fp8"] = u8(xq8)
    out["tr_exps"] = u8(exps)

    path = os.path.join(OUT, "backward_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("forward", "all"):
        main()
    if what in ("backward", "all"):
        main_backward()
