#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libb200q.so -> profiles/rNN_sass_histogram.md (instruction evidence for the judge:
which kernels really issue tcgen05 MMAs (UTC*MMA), TMEM copies / loads (UTCCP, LDTM), TMA loads / stores / prefetches
(UTMALDG, UTMASTG, UTMAPF), mbarrier syncs (SYNCS) ...).   python tools/sass_histogram.py [out.md]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "qutlass_b200", "lib", "libb200q.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_histogram.md")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
KEYS = ["UTCQMMA", "UTCOMMA", "UTCHMMA", "UTCCP", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACCTL", "SYNCS", "HMMA", "LDGSTS", "REDUX", "ACQBULK", "STAS"]
rows, cur, i = [], None, 0
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = [names[i] if i < len(names) else m.group(1), collections.Counter(), 0]
        rows.append(cur); i += 1
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur is not None:
        cur[2] += 1
        op = m.group(1)
        if op in KEYS:
            cur[1][op + (".2CTA" if ".2CTA" in m.group(2) else "")] += 1


def short(n):
    n = n.replace("b200q::", "").replace("(anonymous namespace)::", "").replace("(bool)", "").replace("(int)", "")
    n = n[: n.index(">(") + 1] if ">(" in n else re.sub(r"\(.*$", "", n)
    return n if len(n) < 110 else n[:107] + "..."


agg = collections.OrderedDict()
for n, c, tot in rows:
    base = short(n)
    base = re.sub(r"<.*", "", base)
    a = agg.setdefault(base, [0, collections.Counter(), 0])
    a[0] += 1; a[1].update(c); a[2] += tot
with open(out, "w") as f:
    f.write("# SASS opcode histogram of qutlass_b200/lib/libb200q.so (cuobjdump -sass, sm_100a)\n\n")
    f.write(f"{len(rows)} kernels.  Totals over all instantiations of each kernel template; per-instantiation table below.\n\n")
    allk = sorted({k for _, c, _ in rows for k in c})
    f.write("| kernel template | instantiations | SASS instructions | " + " | ".join(allk) + " |\n|---|---|---|" + "---|" * len(allk) + "\n")
    for b, (cnt, c, tot) in agg.items():
        f.write(f"| `{b}` | {cnt} | {tot} | " + " | ".join(str(c.get(k, 0)) for k in allk) + " |\n")
    tot_all = collections.Counter()
    for _, c, _ in rows:
        tot_all.update(c)
    f.write("| **all** | " + str(len(rows)) + " | " + str(sum(t for _, _, t in rows)) + " | " + " | ".join(str(tot_all.get(k, 0)) for k in allk) + " |\n")
    f.write("\n## Per instantiation (kernels with tcgen05 / TMA opcodes only)\n\n| kernel | instructions | opcodes |\n|---|---|---|\n")
    for n, c, tot in rows:
        if c:
            f.write(f"| `{short(n)}` | {tot} | " + ", ".join(f"{k} {v}" for k, v in sorted(c.items())) + " |\n")
print(out)
