"""Build libb200q.so (hand-written sm_100a CUDA behind the C-ABI in include/b200q.h).

In-tree build: the .so lands in qutlass_b200/lib/ so it travels with the repo snapshot to
the GPU box.  nvcc cross-compiles without a GPU.  Usage: python qutlass_b200/build.py [--force] [--profiling]
(run it as a script: `python -m qutlass_b200.build` imports the package first, which needs an up-to-date library)

--profiling builds a SECOND library, lib/libb200q_prof.so, with -DB200Q_PROFILING: the only build in which
B200Q_GEMM_DEBUG_FLAGS (timing-only switches that skip loads / copies / stores -> wrong results) is honoured.
The probe tools under tools/ select it with B200Q_LIB=prof; the product library ignores the variable.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libb200q.so")
OBJDIR = os.path.join(HERE, "build")

SOURCES = ["common.cu", "quantize.cu", "gemm_fp4.cu", "gemm_decode.cu", "linear_host.cu", "backward.cu", "quantize_tc.cu", "backward_tc.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-I" + os.path.join(ROOT, "include"),
    "-I" + CSRC,
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; set NVCC or add /usr/local/cuda/bin to PATH")
    return cand


def _digest(extra: str = "") -> str:
    h = hashlib.sha256()
    h.update(extra.encode())
    for name in sorted(os.listdir(CSRC)) + ["../../include/b200q.h"]:
        p = os.path.join(CSRC, name)
        if os.path.isfile(p) and name != "torch_ops.cpp":
            h.update(name.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, profiling: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    lib = os.path.join(LIBDIR, "libb200q_prof.so") if profiling else LIB
    objdir = OBJDIR + ("_prof" if profiling else "")
    flags = NVCC_FLAGS + (["-DB200Q_PROFILING"] if profiling else [])
    os.makedirs(objdir, exist_ok=True)
    stamp = lib.replace(".so", ".sha256")
    digest = _digest("prof" if profiling else "")
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        if not profiling:
            build_torch_ops(verbose=verbose)
        return lib
    nvcc = _nvcc()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *flags, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=len(srcs)) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    if not profiling:
        build_torch_ops(force=True, verbose=verbose)
    return lib


OPS_LIB = os.path.join(LIBDIR, "b200q_torch_ops.so")


def build_torch_ops(force: bool = False, verbose: bool = False) -> str:
    """The compiled torch op layer (csrc/torch_ops.cpp: torch stable ABI, no CUDA code) -> lib/b200q_torch_ops.so, linked
    against libb200q.so next to it (rpath $ORIGIN).  g++ only; needs torch's headers."""
    src = os.path.join(CSRC, "torch_ops.cpp")
    stamp = OPS_LIB.replace(".so", ".sha256")
    h = hashlib.sha256(open(src, "rb").read() + open(os.path.join(ROOT, "include", "b200q.h"), "rb").read())
    import torch
    h.update(torch.__version__.encode())
    digest = h.hexdigest()
    if not force and os.path.exists(OPS_LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return OPS_LIB
    tdir = os.path.dirname(torch.__file__)
    inc, tlib = os.path.join(tdir, "include"), os.path.join(tdir, "lib")
    major, minor = (int(x) for x in torch.__version__.split("+")[0].split(".")[:2])
    # The library throws C++ exceptions into libtorch (STD_TORCH_CHECK), so it MUST share libtorch's C++ runtime: a g++
    # whose libstdc++.so is missing links libstdc++.a instead (this image's $CXX = /opt/gcc/bin/g++ does: its own copy of
    # __cxa_throw inside the .so, and a failed argument check then SEGFAULTS on the GPU box instead of raising
    # RuntimeError -- observed, profiles/r02_notes.md).  Candidates are tried in order and the result is verified with nm.
    cands = [c for c in ("/usr/bin/g++", shutil.which("g++"), os.environ.get("CXX")) if c and os.path.exists(c)]
    errors = []
    for cxx in dict.fromkeys(cands):
        cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-DUSE_CUDA", "-Wno-attributes",
               "-DTORCH_TARGET_VERSION=0x%016XULL" % ((major << 56) | (minor << 48)),
               "-I" + inc, "-I" + os.path.join(inc, "torch", "csrc", "api", "include"), "-I" + os.path.join(ROOT, "include"),
               src, "-o", OPS_LIB, "-L" + LIBDIR, "-lb200q", "-L" + tlib, "-ltorch_cpu", "-ltorch_cuda", "-lc10",
               "-Wl,-rpath,$ORIGIN"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            errors.append(f"{cxx}: {r.stderr[-400:]}")
            continue
        syms = subprocess.run(["nm", "-D", OPS_LIB], capture_output=True, text=True).stdout
        if any(l.split()[-2:] == ["T", "__cxa_throw"] for l in syms.splitlines() if "__cxa_throw" in l and len(l.split()) >= 2):
            errors.append(f"{cxx}: linked libstdc++ statically (defines __cxa_throw)")
            os.remove(OPS_LIB)
            continue
        break
    else:
        raise RuntimeError("could not build b200q_torch_ops.so with a shared C++ runtime:\n" + "\n".join(errors))
    with open(stamp, "w") as f:
        f.write(digest)
    return OPS_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True, profiling="--profiling" in sys.argv))
