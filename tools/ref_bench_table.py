#!/usr/bin/env python
"""Side-by-side table of the reference's OWN benchmark (benchmarks/bench_{mxfp4,nvfp4}_sm100.py, unmodified) run against the
drop-in (`ours`) and against the reference's package around its compiled library (`ref`) on the same box
(tools/run_ref_benchmarks.sh).  Reads gpurun_out/ref_bench/{ours,ref}/benchmarks_output/*/*.csv, copies them to
profiles/r02_ref_bench/ and writes profiles/r02_ref_bench/README.md."""
import csv, glob, os, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "gpurun_out", "ref_bench")
DST = os.path.join(ROOT, "profiles", "r02_ref_bench")
os.makedirs(DST, exist_ok=True)


def load(impl):
    out = {}
    for d in sorted(glob.glob(os.path.join(SRC, impl, "benchmarks_output", "*"))):
        for f in glob.glob(os.path.join(d, "*.csv")):
            rows = list(csv.reader(open(f)))
            out[os.path.basename(d)] = rows
            os.makedirs(os.path.join(DST, impl), exist_ok=True)
            shutil.copyfile(f, os.path.join(DST, impl, os.path.basename(d) + ".csv"))
    for f in glob.glob(os.path.join(SRC, impl, "*.log")):
        shutil.copyfile(f, os.path.join(DST, impl, os.path.basename(f)))
    return out


ours, ref = load("ours"), load("ref")
with open(os.path.join(DST, "README.md"), "w") as md:
    md.write("# The reference's own sm_100 benchmarks, unmodified, on one B200: drop-in (ours) vs the reference's package (ref)\n\n"
             "`tools/run_ref_benchmarks.sh` (files staged byte for byte by `oracle/build_ref.py`; flashinfer hidden, matplotlib stubbed: "
             "`oracle/ref_suite_shims/`).  Methodology is the benchmark's: `triton.testing.do_bench_cudagraph(rep=200)`, median; "
             "Llama-3.1-70B layer (K, N) = (8192, 57344); TFLOP/s = 2MNK / t.  `*-cutlass` = quantise + to_blocked + GEMM per iteration, "
             "`*-noquant` = GEMM only (activations pre-quantised).  CSVs and logs of both runs are next to this file.\n\n")
    for name in sorted(ours):
        o, r = ours[name], ref.get(name)
        hdr = o[0]
        cols = [i for i, h in enumerate(hdr) if i > 0 and "-min" not in h and "-max" not in h]
        md.write(f"## {name}\n\n| batch | " + " | ".join(f"{hdr[i].split(' (')[0]} ours / ref (x)" for i in cols) + " |\n|---|" + "---|" * len(cols) + "\n")
        rmap = {row[0]: row for row in (r[1:] if r else [])}
        for row in o[1:]:
            cells = []
            for i in cols:
                a = float(row[i])
                b = float(rmap[row[0]][i]) if row[0] in rmap else None
                cells.append(f"{a:.1f} / {b:.1f} ({a / b:.2f}x)" if b else f"{a:.1f} / -")
            md.write(f"| {int(float(row[0]))} | " + " | ".join(cells) + " |\n")
        md.write("\n")
print(os.path.join(DST, "README.md"))
