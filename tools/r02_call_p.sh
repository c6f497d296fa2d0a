#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -q -k "library-rule" > gpurun_out/r02_s3_bwd_rule_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_s3_bwd_rule_tests.log; tail -3 gpurun_out/r02_s3_bwd_rule_tests.log
