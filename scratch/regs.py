#!/usr/bin/env python
"""registers / spills of the stage kernels from a ptxas log: scratch/regs.py [log] [filter]"""
import re, sys, subprocess
log = sys.argv[1] if len(sys.argv) > 1 else "volumefluxes.jl_b200/csrc/engine_fast.ptxas.log"
flt = sys.argv[2] if len(sys.argv) > 2 else "k_stage"
t = open(log).read()
for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n(?:.*\n)*?.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", t):
    n = m.group(1)
    if flt not in n:
        continue
    d = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    d = re.sub(r"swe::tu_\w+::", "", d)
    d = re.sub(r"\(.*", "", d).replace("void ", "")
    print(f"{d:90s} regs {m.group(5):>3s} stack {m.group(2):>4s} spill {m.group(3)}/{m.group(4)}")
