#!/bin/bash
# usage: gpurun --gpus N -- bash tools/r02_multi.sh N
N="${1:-2}"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== multi-GPU parity (row shards == full product), $N GPUs"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x > gpurun_out/r02_multi_tests_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/r02_multi_tests_n$N.log; tail -6 gpurun_out/r02_multi_tests_n$N.log
echo "== bench --gpus $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "rc=$?"; tail -c 1500 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/r02_bench_n$N.json") if l.startswith("{")][-1])
    print("value %.0f TF, step %.1f us, per-rank step ms %s" % (d["value"], d["ms_per_step"] * 1e3, [round(x, 4) for x in d["per_rank"]["step_ms"]]))
    print("sustained %.0f" % d.get("value_sustained", 0), "c4", d.get("c4", {}).get("value"), d.get("c4", {}).get("ms_per_rank"))
    print("e2e", d.get("e2e", {}).get("value"), d.get("e2e", {}).get("ms_per_step"), "ceiling ms", d.get("e2e", {}).get("host_copy_ceiling_ms"), "numa", d.get("e2e", {}).get("numa_node"))
    print("clocks per rank", [(c.get("sm_mhz"), c.get("reasons")) for c in d["per_rank"]["clocks"]])
except Exception as e:
    print("no line:", e)
PY
echo "== reference arm under torchrun"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
