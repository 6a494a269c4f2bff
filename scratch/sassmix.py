#!/usr/bin/env python
"""usage: sassmix.py <obj> <kernel-substring> [--loop]  -> static opcode histogram (optionally of the
largest backward-branch loop body only)."""
import re, subprocess, sys, collections
obj, pat = sys.argv[1], sys.argv[2]
loop = "--loop" in sys.argv
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur = None; ins = []
kern = {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); kern[cur] = []; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur: kern[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in kern.items():
    if pat not in name: continue
    lo, hi = 0, 1 << 60
    if loop:
        best = None
        for addr, t in ins:
            m = re.search(r"BRA\S*\s+(?:\S+,\s*)?`?\(?0x([0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt < addr and (best is None or addr - tgt > best[1] - best[0]): best = (tgt, addr)
        if best: lo, hi = best
    c = collections.Counter()
    for addr, t in ins:
        if lo <= addr <= hi:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            c[t.split()[0].split(".")[0]] += 1
    tot = sum(c.values())
    print("==", name, "range", hex(lo), hex(hi) if hi < 1 << 60 else "end", "total", tot)
    print("  " + "  ".join(f"{k} {v}" for k, v in c.most_common(40)))
