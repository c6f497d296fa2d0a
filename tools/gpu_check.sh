#!/bin/bash
# One GPU-box session: parity tests, smoke, config sweep, headline bench, ncu launch list.
# usage: tools/gpu_check.sh [tests] [sweep] [bench] [ncu] [ncufull]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
what="${*:-tests sweep bench ncu}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt
if [[ "$what" == *tests* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log
  timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
fi
if [[ "$what" == *sweep* ]]; then
  : > gpurun_out/sweep.jsonl
  for kind in mx nv; do
    for cfg in "1 128" "1 256" "2 128" "2 256"; do
      set -- $cfg
      timeout 300 python bench.py --kind $kind --cta-group $1 --block-n $2 --steps 30 --warmup 5 --no-cpu --no-e2e 2>>gpurun_out/sweep.err | tail -1 >> gpurun_out/sweep.jsonl
    done
  done
  python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    try: d=json.loads(l)
    except Exception: print('bad line', l[:200]); continue
    print(d['config']['workload'][:40], 'step %.1f us'%(d['ms_per_step']*1e3), 'value %.0f TF'%d['value'], 'gemm %.0f TF'%d['gemm_only_tflops_per_gpu'], 'quant %.1f us'%d['quantize_us'], d.get('clocks'))
PY
fi
if [[ "$what" == *bench* ]]; then
  timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; cat gpurun_out/bench_default.json
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_default.err; cat gpurun_out/bench_reference.json
fi
if [[ "$what" == *ncu* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
  tail -3 gpurun_out/ncu_launch.log
fi
if [[ "$what" == *ncufull* ]]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_fp4 -s 4 -c 2 -o gpurun_out/prof_gemm -f \
      python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_full.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:quantize_ -s 4 -c 1 -o gpurun_out/prof_quant -f \
      python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e >> gpurun_out/ncu_full.log 2>&1
  tail -3 gpurun_out/ncu_full.log
fi
