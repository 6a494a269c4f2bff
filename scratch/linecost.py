#!/usr/bin/env python
"""linecost.py <ncu sass-page csv> <obj> <kernel-mangled-substring> <kernel-demangled-substring> <cells>
Joins ncu's per-SASS-instruction 'Instructions Executed' with nvdisasm -gi line info and prints
thread-instructions per cell by (outermost source line in the kernel body, opcode class)."""
import csv, re, subprocess, sys, collections, tempfile, os, glob

ncu_csv, obj, mangled, demangled, cells = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5])
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = glob.glob(tmp + "/*.cubin")[0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout
on = False; cur = None; lines = {}
for l in dis.splitlines():
    if l.startswith("//---") and ".text." in l:
        on = mangled in l; continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        chain = [(m.group(1), int(m.group(2)))] + [(a, int(b)) for a, b in re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))]
        cur = chain; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: lines[int(m.group(1), 16)] = (cur, m.group(2).strip())
rows = list(csv.reader(open(ncu_csv)))
kern = None; hdr = None; per = collections.defaultdict(collections.Counter); base = None
for r in rows:
    if len(r) >= 2 and r[0] == "Kernel Name": kern = r[1]; base = None; continue
    if r and r[0] == "Address": hdr = r; continue
    if kern and demangled in kern and hdr and len(r) > 5 and r[0].startswith("0x"):
        a = int(r[0], 16)
        if base is None: base = a
        off = a - base
        n = int(r[hdr.index("Instructions Executed")])
        chain, txt = lines.get(off, (None, r[1].strip()))
        txt = re.sub(r"^@!?U?P\d+\s+", "", txt)
        op = txt.split()[0].split(".")[0]
        outer = chain[-1][1] if chain else -1
        inner = chain[0] if chain else ("?", -1)
        per[outer][op] += n
        per[outer]["_tot"] += n
        per[outer]["_inner_%s:%d" % (os.path.basename(inner[0]), inner[1])] += n
tot = sum(c["_tot"] for c in per.values())
print(f"total thread-inst/cell {tot * 32 / cells:.1f}")
for line in sorted(per):
    c = per[line]
    if c["_tot"] * 32 / cells < 0.5: continue
    ops = "  ".join(f"{k} {v * 32 / cells:.1f}" for k, v in c.most_common(12) if not k.startswith("_"))
    print(f"line {line:4d}: {c['_tot'] * 32 / cells:6.1f}   {ops}")
