#!/usr/bin/env python
"""Prints, for a kernel in the built library, the instruction mix of its longest backward-branch loop."""
import re, subprocess, sys, collections
lib, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", txt)
for f in funcs[1:]:
    name = f.split("\n", 1)[0]
    if not re.search(pat, name): continue
    ins = []
    for line in f.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
    addr2i = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt in addr2i and addr2i[tgt] < i:
                loops.append((i - addr2i[tgt] + 1, addr2i[tgt], i))
    print(name[:90], f"({len(ins)} instr)")
    for ln, s, e in sorted(loops, key=lambda x: x[1]):
        if ln < 40: continue
        mix = collections.Counter()
        for a, t in ins[s:e + 1]:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            mix[t.split()[0].split(".")[0]] += 1
        print(f"  loop @{ins[s][0]:05x} {ln} instr; " + ", ".join(f"{k}:{v}" for k, v in mix.most_common(14)))
