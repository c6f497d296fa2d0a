#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== pytest -m gpu (no -x)"; timeout 2400 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_pytest_gpu.log; tail -40 gpurun_out/r02_pytest_gpu.log
echo "== decode probe"; B200Q_LIB=prof timeout 300 python tools/decode_probe.py > gpurun_out/r02_decode_probe.jsonl 2> gpurun_out/r02_decode_probe.err; cat gpurun_out/r02_decode_probe.jsonl; tail -3 gpurun_out/r02_decode_probe.err
echo "== fp4 peak probe"; B200Q_LIB=prof timeout 400 python tools/fp4_peak_probe.py > gpurun_out/r02_fp4_peak.jsonl 2> gpurun_out/r02_fp4_peak.err; cat gpurun_out/r02_fp4_peak.jsonl; tail -3 gpurun_out/r02_fp4_peak.err
