#!/bin/bash
# Session-3 call G: backward_qt_bf16 on the tensor cores + the division-free exact QT scale: parity of ALL backward forms, then timed
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== backward tests (all forms)"; timeout 400 python -m pytest tests/test_gpu_backward.py -m gpu -q -x > gpurun_out/r02_s3_bwd_qt_tc_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_s3_bwd_qt_tc_tests.log; tail -30 gpurun_out/r02_s3_bwd_qt_tc_tests.log | cut -c1-250
for tc in 0 1; do
  echo "== bwd bench BWD_QT_TC=$tc"
  B200Q_BWD_QT_TC=$tc timeout 200 python tools/bwd_bench.py --shapes 2048x2048,4096x2048,4096x4096,16384x4096,4096x14336 > gpurun_out/r02_s3_bwd_bench_qt_tc$tc.jsonl 2> gpurun_out/r02_s3_bwd_qt_tc$tc.err
  grep "backward_qt" gpurun_out/r02_s3_bwd_bench_qt_tc$tc.jsonl | cut -c1-170; tail -2 gpurun_out/r02_s3_bwd_qt_tc$tc.err
done
