"""The reference's OWN test files, run UNMODIFIED against the drop-in (north_star: "drops in under tests/ and benchmarks/").

oracle/build_ref.py::stage_suite() copies /root/reference/tests/*.py and benchmarks/bench_{mxfp4,nvfp4}_sm100.py byte for byte
into the git-ignored oracle/_ref/ref_suite/ (it travels to the GPU box like the compiled reference library; /root/reference
itself does not) together with a sha256 manifest.  Each file runs in a child process whose PYTHONPATH puts the repo root
first -- so `import qutlass` resolves to the alias package qutlass/ -> qutlass_b200 -- plus oracle/ref_suite_shims/ (hides
the image's flashinfer, i.e. the reference's BACKENDS = ["cutlass"] configuration, and stubs the missing matplotlib).
Logs land in gpurun_out/ref_suite/ (copied to profiles/ by hand).  A missing staging directory is a skip; a failing
reference test is a FAILURE of this suite and its name is in the assertion message.

The benchmarks (minutes of GPU time each) are run by tools/run_ref_benchmarks.sh, not by pytest.
"""
import hashlib
import json
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SUITE = os.path.join(ROOT, "oracle", "_ref", "ref_suite")
SHIMS = os.path.join(ROOT, "oracle", "ref_suite_shims")
LOGDIR = os.path.join(ROOT, "gpurun_out", "ref_suite")


def _need_suite():
    if not os.path.exists(os.path.join(SUITE, "MANIFEST.json")):
        pytest.skip("oracle/_ref/ref_suite not staged (python oracle/build_ref.py --suite-only in the build container)")
    return json.load(open(os.path.join(SUITE, "MANIFEST.json")))["sha256"]


def test_staged_reference_files_are_byte_identical_to_the_reference():
    """CPU: the staged files match their manifest, and -- where the reference checkout exists (build container) -- the
    checkout itself."""
    man = _need_suite()
    assert {"tests/mxfp4_test.py", "tests/nvfp4_test.py", "tests/quartet_test.py", "benchmarks/bench_mxfp4_sm100.py"} <= set(man)
    for rel, digest in man.items():
        assert hashlib.sha256(open(os.path.join(SUITE, rel), "rb").read()).hexdigest() == digest, rel
        ref = os.path.join("/root/reference", rel)
        if os.path.exists(ref):
            assert hashlib.sha256(open(ref, "rb").read()).hexdigest() == digest, rel


def _child_env():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([ROOT, SHIMS] + ([env["PYTHONPATH"]] if env.get("PYTHONPATH") else []))
    env.pop("B200Q_LIB", None)
    return env


def _run(cmd, log_name, timeout):
    os.makedirs(LOGDIR, exist_ok=True)
    try:
        r = subprocess.run(cmd, cwd=SUITE, env=_child_env(), capture_output=True, text=True, timeout=timeout)
    except subprocess.TimeoutExpired as e:
        with open(os.path.join(LOGDIR, log_name), "w") as f:
            f.write(f"TIMEOUT after {timeout}s\n{(e.stdout or b'').decode(errors='replace') if isinstance(e.stdout, bytes) else (e.stdout or '')}")
        pytest.fail(f"{' '.join(cmd)} timed out after {timeout}s")
    with open(os.path.join(LOGDIR, log_name), "w") as f:
        f.write("$ " + " ".join(cmd) + f"\n[exit code {r.returncode}]\n" + r.stdout + "\n--- stderr ---\n" + r.stderr)
    return r


def _pytest_file(rel, log_name, timeout=1500):
    cmd = [sys.executable, "-m", "pytest", rel, "-q", "-x", "--no-header", "-p", "no:cacheprovider", "--rootdir", SUITE,
           "-c", os.devnull]
    r = _run(cmd, log_name, timeout)
    tail = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ""
    m = re.search(r"(\d+) passed", tail)
    failed = [l for l in r.stdout.splitlines() if l.startswith(("FAILED", "ERROR"))]
    assert r.returncode == 0 and m and not failed, (tail, failed[:10], r.stderr[-600:])
    return int(m.group(1))


@pytest.mark.gpu
def test_reference_mxfp4_tests_pass_unmodified():
    """tests/mxfp4_test.py: 3 + 3 fused-quantise cases (dequantised mismatch <= 1e-4 vs the reference's fp64 oracle, GEMM
    bit-exact) and the Llama 7B-70B shapes at batch 1 / 16 x Hadamard 32 / 64 / 128 (bit-exact)."""
    _need_suite()
    n = _pytest_file("tests/mxfp4_test.py", "mxfp4_test.log")
    assert n == 3 + 3 + 4 * 4 * 2 * 3, n


@pytest.mark.gpu
def test_reference_nvfp4_tests_pass_unmodified():
    """tests/nvfp4_test.py: 4 fused-quantise cases (<= 1e-1, the reference's bar; GEMM bit-exact) + Llama shapes x 4 sizes."""
    _need_suite()
    n = _pytest_file("tests/nvfp4_test.py", "nvfp4_test.log")
    assert n == 4 + 4 * 4 * 2 * 4, n


@pytest.mark.gpu
def test_reference_mxfp8_tests_pass_unmodified():
    """tests/mxfp8_test.py (unittest, collected by pytest): MXFP8 tn / nn GEMMs against torch on pseudo-quantised operands."""
    _need_suite()
    n = _pytest_file("tests/mxfp8_test.py", "mxfp8_test.log")
    assert n >= 1, n


@pytest.mark.gpu
def test_reference_quartet_test_passes_unmodified():
    """tests/quartet_test.py is a script (python quartet_test.py): forward quantisers incl. the clip mask bit-exact against its
    torch restatement, the four backward re-quantisers, MXFP4 / MXFP8 GEMMs.  It prints "Passed!" and the FP8 lines."""
    _need_suite()
    r = _run([sys.executable, "tests/quartet_test.py"], "quartet_test.log", 900)
    assert r.returncode == 0 and "Passed!" in r.stdout, (r.stdout[-800:], r.stderr[-800:])
