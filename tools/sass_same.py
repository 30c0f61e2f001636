#!/usr/bin/env python
"""Do two builds of librt_b200.so carry the same machine code for the kernels they share?  (cuobjdump -sass, per function,
whitespace-normalised, encodings included.)  Used to show that a source change is layout / host only.

    python tools/sass_same.py old.so new.so
"""
import re
import subprocess
import sys


def funcs(path):
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    out = {}
    for p in re.split(r"\n\s*Function : ", txt)[1:]:
        name, body = p.split("\n", 1)
        out[name.strip()] = [re.sub(r"\s+", " ", l).strip() for l in body.splitlines() if "/*" in l and not l.strip().startswith(".")]
    return out


a, b = funcs(sys.argv[1]), funcs(sys.argv[2])
shared = [k for k in a if k in b]
changed = [k for k in shared if a[k] != b[k]]
print(f"{len(shared)} shared kernels, {len(shared) - len(changed)} identical, {len(changed)} changed; only in old: {len(a) - len(shared)}, only in new: {len(b) - len(shared)}")
for k in changed:
    print("  changed:", k)
sys.exit(1 if changed else 0)
