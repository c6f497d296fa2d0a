#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== decode kernel tests first (new tcgen05 kernel: own timeout)"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "decode" > gpurun_out/r02_decode_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_decode_tests.log; tail -15 gpurun_out/r02_decode_tests.log
echo "== pytest -m gpu (all)"; timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_gpu.log; tail -30 gpurun_out/r02_pytest_gpu.log
echo "== ref msweep"; timeout 500 python tools/ref_msweep.py > gpurun_out/r02_ref_msweep.jsonl 2> gpurun_out/r02_ref_msweep.err; tail -3 gpurun_out/r02_ref_msweep.err; cat gpurun_out/ref_msweep.md
echo "== host overhead"; PYTHONPATH=$PWD timeout 200 python tools/host_overhead.py ours > gpurun_out/r02_host_overhead.jsonl 2> gpurun_out/r02_host_overhead.err
B200Q_NO_COMPILED_OPS=1 PYTHONPATH=$PWD timeout 200 python tools/host_overhead.py ours-python-ops >> gpurun_out/r02_host_overhead.jsonl 2>> gpurun_out/r02_host_overhead.err
PYTHONPATH=$PWD/oracle/_ref/ref_pkg:$PWD/oracle/ref_suite_shims timeout 200 python tools/host_overhead.py reference >> gpurun_out/r02_host_overhead.jsonl 2>> gpurun_out/r02_host_overhead.err
cat gpurun_out/r02_host_overhead.jsonl; tail -5 gpurun_out/r02_host_overhead.err
echo "== fp4 peak probe"; B200Q_LIB=prof timeout 500 python tools/fp4_peak_probe.py > gpurun_out/r02_fp4_peak.jsonl 2> gpurun_out/r02_fp4_peak.err; cat gpurun_out/r02_fp4_peak.jsonl; tail -3 gpurun_out/r02_fp4_peak.err
