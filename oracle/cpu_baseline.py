"""CPU baseline for the reference arm / cpu_baseline leg of bench.py.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference has no CPU
implementation of this path (its tests skip without CUDA); the CPU column is the reference's *test
oracle* path -- unpack e2m1, multiply by the block scale (``_dq_fp4``, tests/mxfp4_test.py:84-120),
``torch.matmul`` in float32, round to bf16 (tests/mxfp4_test.py:229-237) -- executed with torch on
all host threads, the dequant done through a 256-entry byte LUT instead of the fp64 gather.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from .fp4_oracle import E2M1_VALUES, e4m3_decode, e8m0_decode

_LUT = None


def _lut() -> torch.Tensor:
    global _LUT
    if _LUT is None:
        b = np.arange(256)
        pair = np.stack([E2M1_VALUES[b & 0xF], E2M1_VALUES[b >> 4]], axis=-1).astype(np.float32)
        _LUT = torch.from_numpy(pair)  # [256, 2]: low nibble first
    return _LUT


def dequant_torch(q: torch.Tensor, sf: torch.Tensor, kind: str) -> torch.Tensor:
    """q uint8 [rows, K/2], sf uint8 [rows, K/group] row-major -> float32 [rows, K]."""
    group = 32 if kind == "mx" else 16
    vals = _lut()[q.long()].reshape(q.shape[0], -1)
    table = e8m0_decode(np.arange(256, dtype=np.uint8)) if kind == "mx" else e4m3_decode(np.arange(256, dtype=np.uint8))
    table = torch.from_numpy(np.nan_to_num(table, nan=0.0).astype(np.float32))
    s = table[sf.long()]
    return (vals.reshape(q.shape[0], -1, group) * s[..., None]).reshape(q.shape[0], -1)


def linear_cpu(aq, asf, bq, bsf, kind: str, alpha: float = 1.0) -> torch.Tensor:
    a = dequant_torch(aq, asf, kind)
    b = dequant_torch(bq, bsf, kind)
    return (torch.matmul(a, b.t()) * alpha).to(torch.bfloat16)


def time_cpu_path(m: int, n: int, k: int, kind: str = "mx", steps: int = 1, warmup: int = 1, seed: int = 0):
    """Times dequant(A) + dequant(B) + fp32 matmul + bf16 round on the host; returns a dict."""
    g = torch.Generator().manual_seed(seed)
    group = 32 if kind == "mx" else 16
    aq = torch.randint(0, 256, (m, k // 2), dtype=torch.uint8, generator=g)
    bq = torch.randint(0, 256, (n, k // 2), dtype=torch.uint8, generator=g)
    lo, hi = (124, 131) if kind == "mx" else (0x30, 0x48)
    asf = torch.randint(lo, hi, (m, k // group), dtype=torch.uint8, generator=g)
    bsf = torch.randint(lo, hi, (n, k // group), dtype=torch.uint8, generator=g)
    for _ in range(warmup):
        linear_cpu(aq[: min(m, 64)], asf[: min(m, 64)], bq, bsf, kind)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = linear_cpu(aq, asf, bq, bsf, kind)
    dt = (time.perf_counter() - t0) / steps
    return dict(seconds_per_step=dt, tflops=2.0 * m * n * k / dt / 1e12, threads=torch.get_num_threads(),
                checksum=float(out.float().abs().mean()))
