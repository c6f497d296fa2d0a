#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== decode tests (cluster-of-8 one-launch step)"; timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "decode" > gpurun_out/r02_decode_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_decode_tests.log; tail -8 gpurun_out/r02_decode_tests.log
echo "== decode probe 2"; B200Q_LIB=prof timeout 300 python tools/decode_probe2.py > gpurun_out/r02_decode_probe2.jsonl 2> gpurun_out/r02_decode_probe2.err; grep "step" gpurun_out/r02_decode_probe2.jsonl; tail -3 gpurun_out/r02_decode_probe2.err
echo "== quant sweep (configs[3])"; timeout 400 python tools/quant_sweep.py > gpurun_out/quant_sweep.jsonl 2> gpurun_out/quant_sweep.err; cat gpurun_out/quant_sweep.md; tail -3 gpurun_out/quant_sweep.err
