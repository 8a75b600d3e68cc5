#!/usr/bin/env python
"""Compares the SASS of two objects / libraries kernel by kernel (instruction text only: addresses, encodings and
line tables ignored).  Runs without a GPU -- the check behind source clean-ups that must not change a kernel.
Usage: sass_diff.py before.o after.o"""
import re
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["cuobjdump", "-sass", path], check=True, capture_output=True, text=True).stdout
    table, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = table.setdefault(m.group(1), [])
            continue
        m = re.search(r"/\*[0-9a-f]{4}\*/\s+(.*?);", line)
        if m and cur is not None:
            cur.append(m.group(1))
    return table


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    same = [k for k in b if a.get(k) == b[k]]
    for k in a:
        if k not in b:
            print("removed  ", k)
    for k in b:
        if k not in a:
            print("added    ", k)
        elif a[k] != b[k]:
            print(f"CHANGED   {k}: {len(a[k])} -> {len(b[k])} instructions")
    print(f"{len(same)} of {len(b)} kernels identical")
    return 0 if len(same) == len(b) and len(a) == len(b) else 1


if __name__ == "__main__":
    sys.exit(main())
