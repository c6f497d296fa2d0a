#!/bin/bash
# Session-3 call N: final tables of the shipped library -- BASELINE configs[3] quantiser sweep and the reference-vs-ours M sweep
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== quant sweep (configs[3])"; QUANT_SWEEP_OUT=r02_final_quant_sweep timeout 400 python tools/quant_sweep.py > gpurun_out/r02_final_quant_sweep.jsonl 2> gpurun_out/r02_final_quant_sweep.err; tail -2 gpurun_out/r02_final_quant_sweep.err; wc -l gpurun_out/r02_final_quant_sweep.jsonl
echo "== ref msweep"; timeout 900 python tools/ref_msweep.py > gpurun_out/r02_final_ref_msweep.jsonl 2> gpurun_out/r02_final_ref_msweep.err; tail -3 gpurun_out/r02_final_ref_msweep.err; cp gpurun_out/ref_msweep.md gpurun_out/r02_final_ref_msweep.md; cat gpurun_out/ref_msweep.md
