#!/bin/bash
# Session-3 call J: ncu --set full captures of the 14-warp tcgen05 quantiser (from the bench command) and of the two tensor-core
# backward kernels (tools/bwd_one.py)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== ncu full: quantiser"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:quantize_tc_kernel -s 6 -c 1 -f -o gpurun_out/r02_s3_prof_quant python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --no-c4 --sustain-s 0 > /dev/null 2>&1; ls -la gpurun_out/r02_s3_prof_quant.ncu-rep
echo "== ncu full: backward_t (tensor cores)"; B200Q_BWD_QT_TC=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:bwd_t_tc_kernel -s 3 -c 1 -f -o gpurun_out/r02_s3_prof_bwd_t python tools/bwd_one.py > gpurun_out/r02_s3_prof_bwd.log 2>&1; ls -la gpurun_out/r02_s3_prof_bwd_t.ncu-rep
echo "== ncu full: backward_qt (tensor cores)"; B200Q_BWD_QT_TC=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:bwd_qt_tc_kernel -s 3 -c 1 -f -o gpurun_out/r02_s3_prof_bwd_qt python tools/bwd_one.py >> gpurun_out/r02_s3_prof_bwd.log 2>&1; ls -la gpurun_out/r02_s3_prof_bwd_qt.ncu-rep
tail -3 gpurun_out/r02_s3_prof_bwd.log
