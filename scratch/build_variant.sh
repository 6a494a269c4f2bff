#!/bin/bash
# builds a variant of the library with extra -D flags: scratch/build_variant.sh NAME "-DX2_STCS=1 ..."
# result: gpurun_out/variants/libswe_b200_NAME.so  (use with SWE_B200_LIB=...)
set -e
name=$1; flags=$2
src=/root/repo/volumefluxes.jl_b200/csrc
out=/root/repo/scratch/variants; mkdir -p $out/$name
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH -Xptxas -v"
nvcc $COMMON -use_fast_math $flags -c $src/engine_fast.cu -o $out/$name/engine_fast.o 2> $out/$name/ptxas.log
nvcc -shared $ARCH -o $out/libswe_b200_$name.so $out/$name/engine_fast.o $src/engine_strict.o $src/capi.o -lcudart
python - <<PY
import re
t=open("$out/$name/ptxas.log").read()
for m in re.finditer(r"Compiling entry function '(\S*k_stage_x2\S*)'.*?\n.*?\n.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", t):
    n=m.group(1)
    if 'ILi1ELb1ELi1ELi0ELb0' in n or 'ILi1ELb1ELi2ELi1ELb0' in n:
        print("$name", n[-40:], "stack", m.group(2), "spill", m.group(3), m.group(4), "regs", m.group(5))
PY
