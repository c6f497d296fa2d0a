#!/bin/bash
# Session-3 call H: 12 epilogue warps (14 warps: 128 registers per thread, no spills, three tiles in flight) vs 16 (96 registers)
# for the tcgen05 quantiser and the tensor-core backward_t kernel: parity of the variant, then both timed.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export B200Q_NO_COMPILED_OPS=1
echo "== parity of the e12 variant (tcgen05 quantiser tests + backward tensorcore form)"
B200Q_LIB=e12 timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py -m gpu -q -x -k "tcgen05 or tensorcore" > gpurun_out/r02_s3_e12_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_s3_e12_tests.log; tail -4 gpurun_out/r02_s3_e12_tests.log | cut -c1-220
for lib in main e12; do
  if [ $lib = e12 ]; then export B200Q_LIB=e12; else unset B200Q_LIB; fi
  echo "== quant sweep lib=$lib"
  QUANT_SWEEP_M=4096,16384 QUANT_SWEEP_OUT=r02_s3_quant_sweep_$lib timeout 300 python tools/quant_sweep.py > gpurun_out/r02_s3_quant_sweep_$lib.jsonl 2> gpurun_out/r02_s3_quant_sweep_$lib.err
  cut -c1-120 gpurun_out/r02_s3_quant_sweep_$lib.jsonl; tail -2 gpurun_out/r02_s3_quant_sweep_$lib.err
  echo "== bwd bench lib=$lib"
  B200Q_BWD_T_TC=1 timeout 200 python tools/bwd_bench.py --shapes 4096x4096,16384x4096,4096x14336 > gpurun_out/r02_s3_bwd_bench_$lib.jsonl 2> gpurun_out/r02_s3_bwd_$lib.err
  grep "backward_t_bf16\"\|comparison" gpurun_out/r02_s3_bwd_bench_$lib.jsonl | cut -c1-170; tail -2 gpurun_out/r02_s3_bwd_$lib.err
done
