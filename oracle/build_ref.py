#!/usr/bin/env python
"""Compile the UNMODIFIED reference (IST-DASLab/qutlass) CUDA extension for sm_100a straight from /root/reference.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the result, oracle/_ref/qutlass_ref_C.so, is the reference's own
`_qutlass_C` torch op library (bindings.cpp + gemm.cu + fused_quantize_*.cu + ..., CUTLASS v4.3.0 from the reference's
third_party/ tree).  It is git-ignored but travels to the GPU box, where tools/ref_compare.py loads it in a separate
process (torch.ops.load_library) as the GPU-side oracle and as the kernel to beat.  Nothing under qutlass_b200/ uses it.

No reference source is copied: nvcc / g++ read the files where they lie; only objects and the .so are written, under
oracle/_ref/.  The reference's own build system (setup.py + cmake; needs a GPU at import, setup.py:45-51,112-118) is
not run; flags follow setup.py:60-92,158-176 (sm_100a only, TARGET_CUDA_ARCH=100).

It also stages the reference's own tests/ and sm_100 benchmarks/ unmodified under oracle/_ref/ref_suite/ (stage_suite()).

usage: python oracle/build_ref.py [--jobs N] [--force] [--suite-only]     (minutes: CUTLASS template instantiation)
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("QUTLASS_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
OBJ = os.path.join(OUT, "build")
LIB = os.path.join(OUT, "qutlass_ref_C.so")
CSRC = os.path.join(REF, "qutlass", "csrc")
SOURCES = ["bindings.cpp", "gemm.cu", "gemm_ada.cu", "fused_quantize_mx.cu", "fused_quantize_mx_mask.cu",
           "fused_quantize_nv.cu", "fused_quantize_mx_sm100.cu", "fused_quantize_nv_sm100.cu", "quartet_bwd_sm120.cu"]
TORCH_TARGET_VERSION = "0x%016XULL" % ((2 << 56) | (11 << 48))      # setup.py:54-57


def _includes():
    import torch
    from torch.utils import cpp_extension as ce
    inc = [os.path.join(CSRC, "include"), os.path.join(CSRC, "include", "cutlass_extensions"),
           os.path.join(REF, "third_party", "cutlass", "include"),
           os.path.join(REF, "third_party", "cutlass", "tools", "util", "include"),
           *ce.include_paths(), sysconfig.get_paths()["include"], "/usr/local/cuda/include"]
    return ["-I" + i for i in inc], os.path.join(os.path.dirname(torch.__file__), "lib")


def build(jobs: int = 4, force: bool = False) -> str:
    if not os.path.isdir(CSRC):
        raise RuntimeError(f"{CSRC} not found: the reference checkout is only present in the build container")
    os.makedirs(OBJ, exist_ok=True)
    srcs_mtime = max(os.path.getmtime(os.path.join(CSRC, s)) for s in SOURCES)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= srcs_mtime:
        return LIB                                   # up to date: nothing to do (what __graft_entry__.build() hits)
    inc, torch_lib = _includes()
    common = ["-DUSE_CUDA", "-DTORCH_TARGET_VERSION=" + TORCH_TARGET_VERSION, "-DTARGET_CUDA_ARCH=100",
              "-DTORCH_EXTENSION_NAME=_CUDA", "-DPy_LIMITED_API=0x03090000", "-std=c++17", "-O3", "-DNDEBUG"]
    nvcc_flags = ["-gencode", "arch=compute_100a,code=sm_100a", "--expt-relaxed-constexpr", "--use_fast_math",
                  "-Xcompiler", "-fPIC", "-Xcompiler", "-funroll-loops", "-Xcompiler", "-ffast-math",
                  "-Xcompiler", "-finline-functions"]

    def compile_one(src):
        obj = os.path.join(OBJ, src.rsplit(".", 1)[0] + ".o")
        log = obj + ".log"
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= os.path.getmtime(os.path.join(CSRC, src)):
            return obj
        if src.endswith(".cu"):
            cmd = ["nvcc", *common, *nvcc_flags, *inc, "-c", os.path.join(CSRC, src), "-o", obj]
        else:
            cmd = ["g++", *common, "-fPIC", *inc, "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        open(log, "w").write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"compiling {src} failed, see {log}\n{r.stderr[-3000:]}")
        print("compiled", src, flush=True)
        return obj

    with ThreadPoolExecutor(max_workers=jobs) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = ["nvcc", "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
           "-L" + torch_lib, "-ltorch", "-ltorch_cpu", "-lc10", "-lcudart", "-lcuda",
           "-Xlinker", "-rpath," + torch_lib]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


SUITE = os.path.join(OUT, "ref_suite")
SUITE_FILES = ["tests/__init__.py", "tests/mxfp4_test.py", "tests/nvfp4_test.py", "tests/mxfp8_test.py",
               "tests/quartet_test.py", "benchmarks/__init__.py", "benchmarks/bench_mxfp4_sm100.py",
               "benchmarks/bench_nvfp4_sm100.py"]


def stage_suite() -> str:
    """Stage the reference's OWN tests/ and sm_100 benchmarks/, byte for byte, under oracle/_ref/ref_suite/ (git-ignored like the
    compiled library, travels to the GPU box like it) together with a sha256 manifest, so that tests/test_gpu_reference_suite.py
    can run them UNMODIFIED against the `qutlass` drop-in package and prove they were not edited.  /root/reference itself
    does not exist on the GPU box.  Nothing is copied into the tracked tree."""
    import hashlib
    import json
    import shutil
    if not os.path.isdir(os.path.join(REF, "tests")):
        raise RuntimeError(f"{REF}/tests not found: the reference checkout is only present in the build container")
    manifest = {}
    for rel in SUITE_FILES:
        src, dst = os.path.join(REF, rel), os.path.join(SUITE, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(src, "rb").read()).hexdigest()
    with open(os.path.join(SUITE, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF, "sha256": manifest}, f, indent=1, sort_keys=True)
    # The reference's own Python package around the compiled library (qutlass/__init__.py does `import qutlass._CUDA`;
    # bindings.cpp:538 REGISTER_EXTENSION(_CUDA) -> PyInit__CUDA is exported by qutlass_ref_C.so): with
    # PYTHONPATH=oracle/_ref/ref_pkg the SAME unmodified tests / benchmarks run against the REAL reference on the GPU box,
    # which is how tools/run_ref_benchmarks.sh produces the reference's curve next to ours.
    pkg = os.path.join(OUT, "ref_pkg", "qutlass")
    os.makedirs(pkg, exist_ok=True)
    for name in ("__init__.py", "utils.py"):
        shutil.copyfile(os.path.join(REF, "qutlass", name), os.path.join(pkg, name))
    if os.path.exists(LIB):
        shutil.copyfile(LIB, os.path.join(pkg, "_CUDA.so"))
    return SUITE


if __name__ == "__main__":
    jobs = int(sys.argv[sys.argv.index("--jobs") + 1]) if "--jobs" in sys.argv else 4
    if "--suite-only" not in sys.argv:
        print(build(jobs=jobs, force="--force" in sys.argv))
    print(stage_suite())
