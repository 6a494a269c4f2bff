#!/bin/bash
# SURVEY 8(f4) feasibility numbers (CPU only: nvcc + cuobjdump): registers, spills and static instructions per row of the
# K2(n)+K1(n+1) sweep (scratch/f4_prototype.cu) against K1 and K2 of the generic one-column kernels, same flags as engine_fast.cu
cd "$(dirname "$0")"
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xptxas -v -use_fast_math -c f4_prototype.cu -o /tmp/f4_prototype.o 2> /tmp/f4_prototype.ptxas.log || { cat /tmp/f4_prototype.ptxas.log; exit 1; }
python - <<'PY'
import re, subprocess
log = open('/tmp/f4_prototype.ptxas.log').read()
regs = {}
for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores.*?\n.*?Used (\d+) registers", log):
    regs[m.group(1)] = (int(m.group(4)), int(m.group(3)))
def loop(name):
    out = subprocess.run(['python', 'sassmix.py', '/tmp/f4_prototype.o', name, '--loop'], capture_output=True, text=True).stdout
    return int(re.search(r"total (\d+)", out).group(1)) / 2          # the march is unrolled by 2 (FUSED_UNROLL)
print(f"{'policy':8s} {'kernel':26s} {'registers':>9s} {'spill B':>8s} {'instr/row':>10s} {'valid cols/32':>14s} {'thread-instr/cell':>18s}")
for pol, tag in (('Float32', 'IffLb0E'), ('Float64', 'IddLb0E')):
    rows = []
    for label, pat, valid in (('K1 (stage 1)', f'k_stageINS0_6Policy{tag}EELi1ELb1ELi1E', 28), ('K2 (stage 2)', f'k_stageINS0_6Policy{tag}EELi1ELb1ELi2E', 28),
                              ('K2(n)+K1(n+1) fused', f'k_stage23INS0_6Policy{tag}EE', 24)):
        name = [k for k in regs if pat in k][0]
        n = loop(pat)
        rows.append((label, regs[name][0], regs[name][1], n, valid, n * 32 / valid))
    for r in rows: print(f"{pol:8s} {r[0]:26s} {r[1]:9d} {r[2]:8d} {r[3]:10.1f} {r[4]:14d} {r[5]:18.1f}")
    sep = rows[0][5] + rows[1][5]
    print(f"{pol:8s} separate K1 + K2: {sep:.1f} thread-instr/cell/step, fused: {rows[2][5]:.1f}  ({100 * (rows[2][5] / sep - 1):+.1f} %);  "
          f"DRAM words/cell/step 17 -> 13 ({100 * (13 / 17 - 1):+.1f} %)")
PY
