#!/bin/bash
# Session-3 call E: backward_t_bf16 on the tensor cores (backward_tc.cu): parity in all forms, then timed against the CUDA-core kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== backward tests (three forms)"; timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q -x > gpurun_out/r02_s3_bwd_tc_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_s3_bwd_tc_tests.log; tail -25 gpurun_out/r02_s3_bwd_tc_tests.log | cut -c1-220
for tc in 0 1; do
  echo "== bwd bench BWD_T_TC=$tc"
  B200Q_BWD_T_TC=$tc timeout 200 python tools/bwd_bench.py --shapes 1024x1024,4096x4096,16384x4096,4096x14336 > gpurun_out/r02_s3_bwd_bench_tc$tc.jsonl 2> gpurun_out/r02_s3_bwd_tc$tc.err
  grep "backward_t_bf16\|comparison" gpurun_out/r02_s3_bwd_bench_tc$tc.jsonl | cut -c1-170; tail -2 gpurun_out/r02_s3_bwd_tc$tc.err
done
