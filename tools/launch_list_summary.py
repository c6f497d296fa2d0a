#!/usr/bin/env python
"""Summarise an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X.csv <command>): total and share per kernel.
usage: tools/launch_list_summary.py gpurun_out/launches.csv "<command that was profiled>" > profiles/rNN_launch_list_summary.txt"""
import csv, sys
from collections import OrderedDict


def main(path, cmd):
    rows = [r for r in csv.reader(l for l in open(path) if not l.startswith("==")) if r]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(",", ""))
        us = v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(r[iu], 1.0)
        t, n = agg.get(r[ik], (0.0, 0))
        agg[r[ik]] = (t + us, n + 1)
    print(f"# ncu launch list of `{cmd}` (gpu__time_duration.sum, --clock-control none;")
    print("# per-launch times are cold-cache and serialised: the SHARE of the step is what must agree with the CUDA-event timing)")
    for k, (t, n) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{t:10.1f} us {n:5d} launches {t / n:9.2f} us each  {k[:140]}")
    own = {k: v for k, v in agg.items() if "b200q::" in k}
    tot = sum(t for t, _ in own.values())
    print(f"\nown kernels: {tot:.1f} us")
    for k, (t, n) in sorted(own.items(), key=lambda kv: -kv[1][0]):
        print(f"  {100 * t / tot:5.1f} %  {k[:120]}  ({n} launches, {t / n:.2f} us each)")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "?")
