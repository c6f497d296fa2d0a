#!/usr/bin/env python
"""Per-tile timeline of CTA 0 of the GEMM (debug flag 1 << 24): where a tile boundary spends its time.
python tools/trace_probe.py [cg bn]"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from qutlass_b200 import _lib
lib = _lib.load(); dev = torch.device("cuda")
raw = ctypes.CDLL(_lib.LIB if hasattr(_lib, "LIB") else os.path.join(ROOT, "qutlass_b200", "lib", "libb200q.so"))
raw.b200q_debug_read_trace.argtypes = [ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
cg, bn = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2, 256)
M, N, K = 4096, 14336, 4096
sets = []
for i in range(3):
    a = torch.randint(0, 256, (M, K // 2), dtype=torch.uint8, device=dev)
    b = torch.randint(0, 256, (N, K // 2), dtype=torch.uint8, device=dev)
    sfa = torch.randint(126, 129, (M * K // 32,), dtype=torch.uint8, device=dev)
    sfb = torch.randint(126, 129, (N * K // 32,), dtype=torch.uint8, device=dev)
    d = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    sets.append((a, b, sfa, sfb, d))
alpha = torch.ones(1, device=dev)
st = torch.cuda.current_stream().cuda_stream
def go(i):
    a, b, sfa, sfb, d = sets[i % 3]
    rc = lib.b200q_gemm_fp4_cfg(a.data_ptr(), b.data_ptr(), sfa.data_ptr(), sfb.data_ptr(), alpha.data_ptr(), d.data_ptr(), M, N, K, 0, cg, bn, st)
    assert rc == 0, lib.b200q_last_error()
for flags in (1 << 24, (1 << 24) | 1, (1 << 24) | 1 | (3 << 20)):
    os.environ["B200Q_GEMM_DEBUG_FLAGS"] = str(flags)
    for i in range(6): go(i)
    buf = (ctypes.c_ulonglong * 256)()
    assert raw.b200q_debug_read_trace(buf, 256) == 0
    ev = [[buf[t * 8 + e] for e in range(6)] for t in range(14)]
    t0 = ev[0][0]
    print(json.dumps(dict(flags=flags, cg=cg, bn=bn, note="cycles rel. to tile 0 acquire: [acc owned, first k-tile landed, last MMA issued, epi sees full, drained, stores issued]")))
    for t in range(13):
        if ev[t][0] == 0: break
        rel = [x - t0 for x in ev[t]]
        nxt = ev[t + 1][0] - t0 if ev[t + 1][0] else None
        print(json.dumps(dict(tile=t, ev=rel, mma_issue_span=rel[2] - rel[0], full_after_last_issue=rel[3] - rel[2], drain=rel[4] - rel[3],
                              epi_store_span=rel[5] - rel[4], next_acquire_after_drained=(nxt - rel[4]) if nxt else None,
                              period=(nxt - rel[0]) if nxt else None)))
