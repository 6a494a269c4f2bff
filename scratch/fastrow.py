#!/usr/bin/env python
"""usage: fastrow.py <obj> <mangled-kernel-substring> -> opcode histogram per thread-row of the kernel's FAST trip
(the three-row straight-line block that ends in the largest forward-skipping region: found as the longest branch-free-ish
span between the '@P BRA <generic>' that guards it and the loop-back)."""
import re, subprocess, sys, collections
obj, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur = None; kern = {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: cur = m.group(1); kern[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur: kern[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in kern.items():
    if pat not in name: continue
    # the FAST trip: between a forward branch that jumps over > 0x3000 bytes and its target
    best = None
    for a, t in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?0x([0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt - a > 0x3000 and (best is None or tgt - a < best[1] - best[0]) : best = (a, tgt)
    lo, hi = best
    c = collections.Counter()
    for a, t in ins:
        if lo < a < hi:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            c[t.split()[0].split(".")[0]] += 1
    tot = sum(c.values())
    print("==", name, hex(lo), hex(hi), "instructions per thread-row %.1f" % (tot / 3))
    print("  " + "  ".join(f"{k} {v/3:.1f}" for k, v in c.most_common(40)))
