#!/bin/bash
# Session-3 call B: persistent double-buffered backward kernels (parity in both forms, then timed), paced weight streaming of the
# decode kernel (parity, then the pace sweep), MXFP8 at M = 1024 against the compiled reference in one process.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== backward tests (both forms)"; timeout 600 python -m pytest tests/test_gpu_backward.py -m gpu -q > gpurun_out/r02_s3_bwd_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_s3_bwd_tests.log; tail -4 gpurun_out/r02_s3_bwd_tests.log
echo "== decode tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "decode" > gpurun_out/r02_s3_decode_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_s3_decode_tests.log; tail -4 gpurun_out/r02_s3_decode_tests.log
for form in 0 1; do
  echo "== bwd bench BWD_PIPE=$form"
  B200Q_BWD_PIPE=$form timeout 200 python tools/bwd_bench.py > gpurun_out/r02_s3_bwd_bench_pipe$form.jsonl 2> gpurun_out/r02_s3_bwd_pipe$form.err
  grep -v generic gpurun_out/r02_s3_bwd_bench_pipe$form.jsonl | cut -c1-160; tail -2 gpurun_out/r02_s3_bwd_pipe$form.err
done
echo "== decode pace probe (mx)"; timeout 300 python tools/decode_pace_probe.py > gpurun_out/r02_s3_decode_pace_mx.jsonl 2> gpurun_out/r02_s3_decode_pace_mx.err; cat gpurun_out/r02_s3_decode_pace_mx.jsonl; tail -3 gpurun_out/r02_s3_decode_pace_mx.err
echo "== decode pace probe (nv)"; PROBE_KIND=1 PROBE_PACES=0,460,540,620 timeout 300 python tools/decode_pace_probe.py > gpurun_out/r02_s3_decode_pace_nv.jsonl 2> gpurun_out/r02_s3_decode_pace_nv.err; cat gpurun_out/r02_s3_decode_pace_nv.jsonl; tail -3 gpurun_out/r02_s3_decode_pace_nv.err
echo "== f8 probe vs reference"; F8_PROBE_M=1024 timeout 200 python tools/f8_probe.py > gpurun_out/r02_s3_f8_probe_ref.jsonl 2> gpurun_out/r02_s3_f8_probe_ref.err; cat gpurun_out/r02_s3_f8_probe_ref.jsonl; tail -3 gpurun_out/r02_s3_f8_probe_ref.err
