"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads without a GPU and exports
every symbol include/b200q.h declares, with the argument counts the Python binding assumes."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib_path():
    from qutlass_b200 import build
    return build.build()


def _declared():
    src = open(os.path.join(ROOT, "include", "b200q.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(int|int64_t|const char\*|void|unsigned)\s+(b200q_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(3).strip()
        n = 0 if args in ("void", "") else len([a for a in args.split(",") if a.strip()])
        decls[m.group(2)] = n
    return decls


def test_header_declares_the_hot_path_entry_points():
    d = _declared()
    for name in ("b200q_quantize_mx", "b200q_quantize_nv", "b200q_swizzle_sf", "b200q_gemm_fp4",
                 "b200q_gemm_fp4_cfg", "b200q_last_error", "b200q_abi_version", "b200q_linear_fp4_host",
                 "b200q_linear_workspace_bytes", "b200q_gemm_fp4_launches"):
        assert name in d, name


def test_library_loads_and_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in include/b200q.h but not exported"


def test_python_binding_signatures_match_header(lib_path):
    from qutlass_b200 import _lib
    d = _declared()
    assert set(_lib.SIGNATURES) == set(d)
    for name, (_, args) in _lib.SIGNATURES.items():
        assert len(args) == d[name], name
    assert _lib.load().b200q_abi_version() == 1


def test_no_libcuda_or_torch_link_dependency(lib_path):
    """plain C-ABI: no torch types, and loadable on a box without the driver (cudart is static)."""
    import subprocess
    out = subprocess.run(["ldd", lib_path], capture_output=True, text=True).stdout
    assert "libtorch" not in out and "libcuda.so" not in out and "libc10" not in out


def test_product_never_imports_the_oracle():
    """the oracle is test infrastructure: nothing under qutlass_b200/ (or the qutlass alias) may use it."""
    for pkg in ("qutlass_b200", "qutlass"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), os.path.join(dirpath, f)


def test_gemm_planner_choices_without_a_gpu(lib_path):
    """b200q_gemm_fp4_plan is host-only (no device: 148 SMs assumed).  The choices below are the measured optima on B200
    (profiles/r01_plan_probe.jsonl): weight-streaming single-CTA tiles up to 128 rows, one round of single-CTA 128 x 256
    tiles up to 256 rows, CTA pairs beyond -- and never the dispatch-bound 128-column pair tile on the FFN shapes."""
    from qutlass_b200 import _lib
    lib = _lib.load()

    def plan(m, n, k, kind=0):
        cg, bn = ctypes.c_int(0), ctypes.c_int(0)
        assert lib.b200q_gemm_fp4_plan(m, n, k, kind, ctypes.byref(cg), ctypes.byref(bn)) == 0
        return cg.value, bn.value

    # decode (M <= 32, FP4, K % 256 == 0, activations resident): the swapped-operand weight-streaming kernel, "(1, 16)"
    assert plan(1, 14336, 4096) == (1, 16) and plan(16, 14336, 4096) == (1, 16) and plan(32, 14336, 4096) == (1, 16)
    assert plan(16, 28672, 8192) == (1, 16) and plan(32, 28672, 8192) == (1, 128)   # 32 rows x K = 8192: x no longer fits next to a ring
    assert plan(16, 14336, 4096, 1) == (1, 16)
    assert plan(16, 28672, 8192, 1) == (1, 128)         # NVFP4 at K = 8192: the activation scales do not fit TMEM next to the rest
    assert plan(16, 14336, 4096, 2) == (1, 128)         # MXFP8 keeps the general kernel
    assert plan(16, 14336, 4000) == (1, 128)            # K % 256 != 0
    assert plan(33, 14336, 4096) == (1, 128) and plan(128, 14336, 4096) == (1, 128)
    assert plan(256, 14336, 4096) == (1, 256)
    assert plan(4096, 14336, 4096) == (2, 256) and plan(4096, 14336, 4096, 1) == (2, 256)      # BASELINE configs 1 / 2
    assert plan(16384, 14336, 4096) == (2, 256) and plan(2048, 28672, 8192) == (2, 256)        # config 4 shard
    assert plan(1024, 6144, 4096) == (2, 192)
    for m in (384, 512, 768, 1024, 1536, 2048, 3072, 8192):
        assert plan(m, 14336, 4096)[0] == 2 and plan(m, 14336, 4096)[1] in (192, 256)
    assert plan(4096, 4096, 14336) == (2, 256)
    # MXFP8 has cost constants of its own (profiles/r02_s3_f8_probe.jsonl): 224 tiles of 256 x 256 on 74 CTA pairs are 3.03 rounds at
    # M = 1024 -> (2, 192); the big shapes stay on (2, 256)
    assert plan(1024, 14336, 4096, 2) == (2, 192) and plan(1024, 14336, 4096, 3) == (2, 192)
    assert plan(4096, 14336, 4096, 2) == (2, 256) and plan(16384, 14336, 4096, 2) == (2, 256)
    assert lib.b200q_gemm_fp4_plan(0, 1, 1, 0, ctypes.byref(ctypes.c_int()), ctypes.byref(ctypes.c_int())) != 0


def test_host_entry_slab_schedule_without_a_gpu(lib_path):
    """b200q_linear_host_slabs is host-only: the row slabs b200q_linear_fp4_host pipelines (H2D | kernels | D2H).  Slabs
    tile [0, M) in order, every inner bound is a multiple of 128 (so each slab's blocked scales are self-contained), there
    are at most 16, and beyond one regular slab the first is a single 128-row block (the result copy starts early)."""
    from qutlass_b200 import _lib
    lib = _lib.load()

    def slabs(m):
        b = (ctypes.c_int * 17)()
        n = lib.b200q_linear_host_slabs(m, b, 17)
        assert 1 <= n <= 16, (m, n)
        return list(b[: n + 1])

    assert slabs(1) == [0, 1] and slabs(200) == [0, 200] and slabs(512) == [0, 512]
    assert slabs(513) == [0, 128, 512, 513]
    assert slabs(4096) == [0, 128, 512] + list(range(1024, 4097, 512))        # BASELINE config 1: 9 slabs
    for m in (1, 127, 128, 129, 600, 1100, 4096, 7680, 7681, 16384, 100000, 1 << 20, (1 << 20) + 77):
        b = slabs(m)
        assert b[0] == 0 and b[-1] == m and all(x < y for x, y in zip(b, b[1:]))
        assert all(x % 128 == 0 for x in b[:-1])
    assert lib.b200q_linear_host_slabs(0, (ctypes.c_int * 17)(), 17) < 0
    assert lib.b200q_linear_host_slabs(4096, (ctypes.c_int * 4)(), 4) < 0       # not enough room for the bounds


def test_product_library_has_no_profiling_switches(lib_path, monkeypatch):
    """the product library must not read timing-only switches that corrupt results.  B200Q_GEMM_DEBUG_FLAGS and
    the hybrid (2, 448) tile pairs (measured not faster, profiles/r02_notes.md) exist only in the -DB200Q_PROFILING build."""
    from qutlass_b200 import _lib
    lib = _lib.load()
    assert lib.b200q_profiling_build() == 0

    def plan(m, n, k, kind=0):
        cg, bn = ctypes.c_int(0), ctypes.c_int(0)
        assert lib.b200q_gemm_fp4_plan(m, n, k, kind, ctypes.byref(cg), ctypes.byref(bn)) == 0
        return cg.value, bn.value

    monkeypatch.setenv("B200Q_GEMM_HYBRID", "1")
    monkeypatch.setenv("B200Q_GEMM_DEBUG_FLAGS", "33")
    _lib.reload_env()
    try:
        assert plan(4096, 14336, 4096) == (2, 256)
        assert b"gemm_fp4_hybrid_kernel" not in open(lib_path, "rb").read()
    finally:
        monkeypatch.delenv("B200Q_GEMM_HYBRID")
        monkeypatch.delenv("B200Q_GEMM_DEBUG_FLAGS")
        _lib.reload_env()


def test_environment_switches_are_cached_until_reload(lib_path, monkeypatch):
    """The library reads its switches once (no getenv per launch); b200q_reload_env re-reads them."""
    from qutlass_b200 import _lib
    lib = _lib.load()

    def launches():
        return lib.b200q_gemm_fp4_launches(4096, 14336 + 256, 4096, 0)

    monkeypatch.delenv("B200Q_TAIL_SPLIT", raising=False)
    _lib.reload_env()
    base = launches()
    monkeypatch.setenv("B200Q_TAIL_SPLIT", "1")
    assert launches() == base              # not re-read yet
    _lib.reload_env()
    after = launches()
    monkeypatch.delenv("B200Q_TAIL_SPLIT")
    _lib.reload_env()
    assert launches() == base
    assert after in (1, 2)




def test_compiled_op_layer_shares_the_c_plus_plus_runtime_with_torch(lib_path):
    """lib/b200q_torch_ops.so throws C++ exceptions into libtorch (argument checks).  Built with a g++ whose libstdc++.so is
    missing it silently links libstdc++.a, carries its own __cxa_throw, and a failed check then segfaults on the GPU box
    instead of raising RuntimeError (observed in round 2).  The build verifies this; so does this test."""
    import subprocess
    ops = os.path.join(os.path.dirname(lib_path), "b200q_torch_ops.so")
    assert os.path.exists(ops)
    syms = subprocess.run(["nm", "-D", ops], capture_output=True, text=True).stdout
    lines = [l.split() for l in syms.splitlines() if "__cxa_throw" in l or "__gxx_personality_v0" in l]
    assert lines and all(l[0] == "U" for l in lines), lines
    import torch
    import qutlass_b200 as Q
    assert Q._COMPILED_OPS
    with pytest.raises(RuntimeError):          # CPU tensors: the dispatcher (not our code) rejects them -- still an exception, not a crash
        torch.ops._qutlass_C.matmul_mxf4_bf16_tn(*[torch.zeros(4, 16, dtype=torch.uint8)] * 4, torch.ones(1))
