#!/bin/bash
# Session-3 call F: tile order of the tensor-core backward_t kernel (m-tiles fastest vs n-tiles fastest), parity, 2M / 4M / 8M sizes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== backward tests"; timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q -x -k "tensorcore" > gpurun_out/r02_s3_bwd_tc_tests2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_s3_bwd_tc_tests2.log; tail -4 gpurun_out/r02_s3_bwd_tc_tests2.log | cut -c1-220
for order in m n; do
  echo "== bwd bench TC, ${order}-tiles fastest"
  if [ $order = n ]; then export B200Q_BWD_T_NFAST=1; else unset B200Q_BWD_T_NFAST; fi
  B200Q_BWD_T_TC=1 timeout 200 python tools/bwd_bench.py --shapes 2048x1024,2048x2048,4096x2048,4096x4096,16384x4096,4096x14336,14336x4096 > gpurun_out/r02_s3_bwd_bench_tc_${order}fast.jsonl 2> gpurun_out/r02_s3_bwd_tc_${order}.err
  grep "backward_t_bf16\"" gpurun_out/r02_s3_bwd_bench_tc_${order}fast.jsonl | cut -c1-170; tail -2 gpurun_out/r02_s3_bwd_tc_${order}.err
done
unset B200Q_BWD_T_NFAST
echo "== CUDA-core kernel at the small sizes"; B200Q_BWD_T_TC=0 timeout 200 python tools/bwd_bench.py --shapes 2048x1024,2048x2048,4096x2048,14336x4096 > gpurun_out/r02_s3_bwd_bench_tc0_small.jsonl 2>/dev/null; grep "backward_t_bf16\"" gpurun_out/r02_s3_bwd_bench_tc0_small.jsonl | cut -c1-170
