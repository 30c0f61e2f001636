#!/usr/bin/env python
"""Do two builds of librt_b200.so carry the same machine code for the kernels they share?  (cuobjdump -sass, per function,
whitespace-normalised, encodings included.)  Used to show that a source change is layout / host only.

    python tools/sass_same.py old.so new.so [--lines]

--lines: for every kernel that changed, how many instruction lines differ, how many of those differ only in a constant-bank
offset (a kernel parameter that moved), and where the remaining ones sit (first / last differing line of the listing).
"""
import difflib
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


def instr(line):
    """The instruction text without its address and encoding."""
    return re.sub(r"/\*[0-9a-fx ]+\*/", "", line).strip()


def no_const_offset(text):
    return re.sub(r"c\[0x0\]\[0x[0-9a-f]+\]", "c[0x0][*]", text)


args = [a for a in sys.argv[1:] if not a.startswith("--")]
a, b = funcs(args[0]), funcs(args[1])
shared = [k for k in a if k in b]
changed = [k for k in shared if a[k] != b[k]]
print(f"{len(shared)} shared kernels, {len(shared) - len(changed)} identical, {len(changed)} changed; only in old: {len(a) - len(shared)}, only in new: {len(b) - len(shared)}")
for k in changed:
    if "--lines" not in sys.argv:
        print("  changed:", k)
        continue
    ia = [instr(l) for l in a[k] if instr(l)]
    ib = [instr(l) for l in b[k] if instr(l)]
    sm = difflib.SequenceMatcher(None, ia, ib, autojunk=False)
    diff_lines, only_offset, other_at = 0, 0, []
    for tag, i1, i2, j1, j2 in sm.get_opcodes():
        if tag == "equal":
            continue
        n = max(i2 - i1, j2 - j1)
        diff_lines += n
        if tag == "replace" and i2 - i1 == j2 - j1 and all(no_const_offset(x) == no_const_offset(y) for x, y in zip(ia[i1:i2], ib[j1:j2])):
            only_offset += n
        else:
            other_at.append(i1)
    where = f", others between listing lines {min(other_at)} and {max(other_at)}" if other_at else ""
    print(f"  changed: {k}: {diff_lines} of {len(ia)} instructions differ, {only_offset} only in a constant-bank offset{where}")
sys.exit(1 if changed else 0)
