#!/usr/bin/env python
"""usage: sassdump.py <obj> <kernel-substring> [lo hi]  -> plain SASS listing of a kernel (address range)."""
import re, subprocess, sys
obj, pat = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4], 16) if len(sys.argv) > 4 else 1 << 60
txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
on = False
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: on = pat in m.group(1); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and on:
        a = int(m.group(1), 16)
        if lo <= a <= hi: print(f"{a:05x} {m.group(2).strip()}")
