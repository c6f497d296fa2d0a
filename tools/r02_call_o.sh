#!/bin/bash
# ncu --set full of the decode kernel of the final library (paced weight stream)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_fp4_decode_kernel -s 3 -c 1 -f -o gpurun_out/r02_s3_prof_decode python tools/decode_one.py > gpurun_out/r02_s3_prof_decode.log 2>&1; ls -la gpurun_out/r02_s3_prof_decode.ncu-rep; tail -2 gpurun_out/r02_s3_prof_decode.log
