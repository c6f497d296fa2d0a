#!/usr/bin/env python
"""Which device kernels changed between two builds?  Compares the SASS (cuobjdump -sass) of every kernel in two objects /
shared libraries by mangled name.  Used to prove that adding an opt-in kernel left the GPU-validated ones byte-identical.
usage: python tools/sass_diff.py OLD.{o,so} NEW.{o,so}"""
import re, subprocess, sys


def funcs(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    parts = re.split(r"\n\s*Function : ", txt)[1:]
    return {f.split("\n")[0].strip(): "\n".join(f.split("\n")[1:]) for f in parts}


if __name__ == "__main__":
    a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
    changed = [k for k in a if k in b and a[k] != b[k]]
    print(f"{len(a)} kernels in OLD, {len(b)} in NEW: {sum(1 for k in a if k in b and a[k] == b[k])} identical, "
          f"{len(changed)} changed, {len([k for k in a if k not in b])} removed, {len([k for k in b if k not in a])} new")
    for k in changed:
        print("changed:", k)
    for k in b:
        if k not in a:
            print("new:", k)
    sys.exit(1 if changed else 0)
