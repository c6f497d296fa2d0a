#!/usr/bin/env python
"""Which device kernels changed between two builds?  Extracts every kernel's instruction stream (cuobjdump -sass, addresses
and encodings stripped) from two objects / shared libraries and reports the kernels of OLD that have no instruction-identical
kernel in NEW (a renamed template instantiation with the same code counts as unchanged).  Used to prove that adding opt-in
kernels left the GPU-validated ones untouched.
usage: python tools/sass_diff.py OLD.{o,so} NEW.{o,so}      exit status 1 if a kernel of OLD changed or vanished"""
import collections, re, subprocess, sys


def funcs(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    out = {}
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        ins = [m.group(1).strip() for m in (re.search(r"/\*[0-9a-f]{4,5}\*/\s+(.*?);", l) for l in f.split("\n")[1:]) if m]
        out[f.split("\n")[0].strip()] = "\n".join(ins)
    return out


if __name__ == "__main__":
    a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
    have = collections.Counter(b.values())
    missing = [k for k in a if have[a[k]] == 0]
    old_bodies = set(a.values())
    new = [k for k in b if b[k] not in old_bodies]
    print(f"{len(a)} kernels in OLD, {len(b)} in NEW: {len(a) - len(missing)} of OLD instruction-identical in NEW, "
          f"{len(missing)} changed or removed, {len(new)} kernels with new code")
    for k in missing:
        print("changed/removed:", k)
    for k in new:
        print("new:", k)
    sys.exit(1 if missing else 0)
