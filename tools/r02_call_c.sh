#!/bin/bash
# Session-3 call C: decode kernel with the activation-scale copies interleaved into the k loop (parity, then timing against
# call B's numbers of the same harness), library pacing rule vs off, refreshed reference-vs-ours M sweep (the committed one
# predates the compile-out of the profiling flags, which handicapped M = 33..128).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== decode tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "decode" > gpurun_out/r02_s3_decode_tests2.log 2>&1; echo "rc=$?" >> gpurun_out/r02_s3_decode_tests2.log; tail -4 gpurun_out/r02_s3_decode_tests2.log
echo "== mxf8 / planner tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "f8 or mxf8 or plan" > gpurun_out/r02_s3_f8_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_s3_f8_tests.log; tail -3 gpurun_out/r02_s3_f8_tests.log
echo "== decode pace probe (mx): off vs library rule"; PROBE_PACES=0,-1 timeout 300 python tools/decode_pace_probe.py > gpurun_out/r02_s3_decode_interleave_mx.jsonl 2> gpurun_out/r02_s3_decode_interleave_mx.err; cat gpurun_out/r02_s3_decode_interleave_mx.jsonl; tail -3 gpurun_out/r02_s3_decode_interleave_mx.err
echo "== decode pace probe (nv)"; PROBE_KIND=1 PROBE_PACES=0,-1 timeout 300 python tools/decode_pace_probe.py > gpurun_out/r02_s3_decode_interleave_nv.jsonl 2> gpurun_out/r02_s3_decode_interleave_nv.err; cat gpurun_out/r02_s3_decode_interleave_nv.jsonl; tail -3 gpurun_out/r02_s3_decode_interleave_nv.err
echo "== ref msweep"; timeout 900 python tools/ref_msweep.py > gpurun_out/r02_s3_ref_msweep.jsonl 2> gpurun_out/r02_s3_ref_msweep.err; tail -3 gpurun_out/r02_s3_ref_msweep.err; cat gpurun_out/ref_msweep.md
