#!/bin/bash
# Session-3 call A: (1) the new chunk-per-thread epilogue of the tcgen05 quantiser: parity first, then both forms timed;
# (2) full GPU suite on the rebuilt library; (3) MXFP8 configuration sweep at M = 1024 (the one cell of
# profiles/r02_ref_msweep.md that loses to the reference); (4) backward-kernel baseline.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== tcgen05 quantiser tests"; timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tcgen05" > gpurun_out/r02_s3_tc_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_s3_tc_tests.log; tail -4 gpurun_out/r02_s3_tc_tests.log
for split in 0 1; do
  echo "== quant sweep split=$split"
  B200Q_QUANT_TC_SPLIT=$split QUANT_SWEEP_M=1024,4096,16384 QUANT_SWEEP_OUT=r02_s3_quant_sweep_split$split timeout 300 python tools/quant_sweep.py > gpurun_out/r02_s3_quant_sweep_split$split.jsonl 2> gpurun_out/r02_s3_quant_sweep_split$split.err
  cut -c1-160 gpurun_out/r02_s3_quant_sweep_split$split.jsonl; tail -2 gpurun_out/r02_s3_quant_sweep_split$split.err
done
echo "== pytest -m gpu (all)"; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_s3_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r02_s3_pytest_gpu.log; tail -4 gpurun_out/r02_s3_pytest_gpu.log
echo "== f8 probe"; F8_PROBE_M=1024,2048 timeout 200 python tools/f8_probe.py > gpurun_out/r02_s3_f8_probe.jsonl 2> gpurun_out/r02_s3_f8_probe.err; cat gpurun_out/r02_s3_f8_probe.jsonl; tail -3 gpurun_out/r02_s3_f8_probe.err
echo "== bwd bench"; timeout 200 python tools/bwd_bench.py > gpurun_out/r02_s3_bwd_bench.jsonl 2> gpurun_out/r02_s3_bwd.err; cat gpurun_out/r02_s3_bwd_bench.jsonl | cut -c1-200; tail -3 gpurun_out/r02_s3_bwd.err
