#!/bin/bash
# Offline (no GPU) look at what ptxas makes of the variable-base kernel: compiles a minimal translation unit
# (tools/sass/k_var_base_min.cu: the kernel of pa_kernels.cuh with plain loaders) for sm_100a and prints static SASS
# opcode counts per function (kernel body, fe_mul, fe_sqr, jac_dbl) and per loop (table build, doubling body,
# mixed-addition body, whole window loop).  The counts of the loop bodies x (130 doublings, ~62 additions) plus the
# callees x their call counts track the executed-instruction count ncu reports to within a few per cent, which is
# what the kernel time follows (profiles/r01f_ab_field_arithmetic.txt).
#   tools/sass/mix.sh [extra nvcc flags, e.g. -DPA_OPQ_MODE=3]
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=$HERE/../../privacy-auction_b200/csrc
OUT=${TMPDIR:-/tmp}/pa_sass_mix
mkdir -p "$OUT"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I"$SRC" -Xptxas -v "$@" -cubin -o "$OUT/k.cubin" "$HERE/k_var_base_min.cu" 2>&1 | grep -E "registers|spill|error" | head
cuobjdump -sass "$OUT/k.cubin" > "$OUT/k.sass"
python3 "$HERE/mix.py" "$OUT/k.sass"
