#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== decode tests first"; timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "decode" > gpurun_out/r02_decode_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_decode_tests.log; tail -8 gpurun_out/r02_decode_tests.log
echo "== pytest -m gpu (all)"; timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_gpu.log; tail -30 gpurun_out/r02_pytest_gpu.log
echo "== decode probe 2"; B200Q_LIB=prof timeout 300 python tools/decode_probe2.py > gpurun_out/r02_decode_probe2.jsonl 2> gpurun_out/r02_decode_probe2.err; cat gpurun_out/r02_decode_probe2.jsonl; tail -3 gpurun_out/r02_decode_probe2.err
echo "== host overhead"; PYTHONPATH=$PWD timeout 200 python tools/host_overhead.py ours > gpurun_out/r02_host_overhead.jsonl 2> gpurun_out/r02_host_overhead.err
B200Q_NO_COMPILED_OPS=1 PYTHONPATH=$PWD timeout 200 python tools/host_overhead.py ours-python-ops >> gpurun_out/r02_host_overhead.jsonl 2>> gpurun_out/r02_host_overhead.err
PYTHONPATH=$PWD/oracle/_ref/ref_pkg:$PWD/oracle/ref_suite_shims timeout 400 python tools/host_overhead.py reference >> gpurun_out/r02_host_overhead.jsonl 2>> gpurun_out/r02_host_overhead.err; echo "reference rc=$?"
cat gpurun_out/r02_host_overhead.jsonl; tail -5 gpurun_out/r02_host_overhead.err
