#!/bin/bash
# Session-3 call K: backward_qt tensor-core kernel with the conflict-free decode mapping: parity, timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== backward tests (tensorcore form)"; timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -q -x -k "tensorcore" > gpurun_out/r02_s3_call_k_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_s3_call_k_tests.log; tail -3 gpurun_out/r02_s3_call_k_tests.log | cut -c1-200
B200Q_BWD_QT_TC=1 timeout 200 python tools/bwd_bench.py --shapes 4096x4096,16384x4096,4096x14336 > gpurun_out/r02_s3_bwd_bench_k.jsonl 2> gpurun_out/r02_s3_bwd_k.err
grep "backward_qt" gpurun_out/r02_s3_bwd_bench_k.jsonl | cut -c1-170; tail -2 gpurun_out/r02_s3_bwd_k.err
