#!/bin/bash
# Session-3 call I: 12-epilogue-warp kernels as the library default (tcgen05 quantiser, tensor-core backward_t), backward_qt on the
# tensor cores with 14 warps / 128 registers: parity (quantiser + backward, all forms), then timed.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tests"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py -m gpu -q -x -k "quantize or quantise or tcgen05 or backward or square or mxfp4_transpose or fp8_requant" > gpurun_out/r02_s3_call_i_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_s3_call_i_tests.log; tail -4 gpurun_out/r02_s3_call_i_tests.log | cut -c1-220
for tc in 0 1; do
  echo "== bwd bench BWD_QT_TC=$tc"
  B200Q_BWD_QT_TC=$tc timeout 200 python tools/bwd_bench.py --shapes 4096x4096,16384x4096,4096x14336 > gpurun_out/r02_s3_bwd_bench_i_qt$tc.jsonl 2> gpurun_out/r02_s3_bwd_i_qt$tc.err
  grep -v "generic\|square" gpurun_out/r02_s3_bwd_bench_i_qt$tc.jsonl | cut -c1-170; tail -2 gpurun_out/r02_s3_bwd_i_qt$tc.err
done
