#!/bin/bash
# The reference's OWN benchmark files (benchmarks/bench_{mxfp4,nvfp4}_sm100.py, staged unmodified by oracle/build_ref.py),
# run twice on the same box: against the `qutlass` drop-in (ours) and against the reference's own package around its compiled
# library (oracle/_ref/ref_pkg).  Output: gpurun_out/ref_bench/{ours,ref}/<name>.log + the CSVs triton.testing writes.
# usage (repo root, GPU box):  bash tools/run_ref_benchmarks.sh [mxfp4|nvfp4|both] [ours|ref|both]
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
WHICH="${1:-both}"; IMPL="${2:-both}"
SUITE="$ROOT/oracle/_ref/ref_suite"; SHIMS="$ROOT/oracle/ref_suite_shims"
[ -f "$SUITE/benchmarks/bench_mxfp4_sm100.py" ] || { echo "ref_suite not staged"; exit 0; }
run() {  # $1 = impl, $2 = bench name
  local out="$ROOT/gpurun_out/ref_bench/$1"; mkdir -p "$out"; cd "$out"
  local pp="$ROOT:$SHIMS"; [ "$1" = ref ] && pp="$ROOT/oracle/_ref/ref_pkg:$SHIMS"
  local t0=$(date +%s)
  PYTHONPATH="$pp" timeout "${BENCH_TIMEOUT:-900}" python "$SUITE/benchmarks/bench_$2_sm100.py" > "$2.log" 2>&1
  echo "[$1 $2] rc=$? $(( $(date +%s) - t0 )) s" | tee -a "$out/$2.log"
}
for b in mxfp4 nvfp4; do
  [ "$WHICH" = both ] || [ "$WHICH" = "$b" ] || continue
  for i in ours ref; do
    [ "$IMPL" = both ] || [ "$IMPL" = "$i" ] || continue
    run "$i" "$b"
  done
done
