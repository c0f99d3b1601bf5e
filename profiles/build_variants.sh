#!/bin/bash
# Development aid: build tuning variants of the whole library for an A/B run on the GPU box.
# usage: profiles/build_variants.sh name:"extra nvcc flags" ...   -> profiles/variants/lib_<name>.so
#   e.g. profiles/build_variants.sh base: rows:"-DV5_PAIR_ROWS=0" split:"-DV5_SPLIT_BARRIER=1" narrow:"-DV5_TW_MAX=18 -DV5_MIN_CTAS=3"
# then, in ONE gpurun call (profiles/variants/ travels with the snapshot and is git-ignored):
#   for v in base rows base rows; do V5ELA_LIB=profiles/variants/lib_$v.so python profiles/variant_perf.py; done
# Compile-time knobs: V5_NT, V5_TW_MAX, V5_MIN_CTAS (CTA shape), V5_CONST_SEL, V5_NO_FAST_PATH, V5_RINGD, V5_PAIR_ROWS,
# V5_PAIR_UNROLL, V5_SPLIT_BARRIER, V5_CONVERT_UNROLL (fused kernel), V5J_HUFF_NT, V5J_SUB_BITS, V5J_HUFF_CTAS (Huffman decoder).
set -e
cd "$(dirname "$0")/.."
mkdir -p profiles/variants
C=fake-video-detection-engine_b200/csrc
for v in "$@"; do
  name=${v%%:*}; flags=${v#*:}
  out=profiles/variants/lib_${name}.so
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       -Iinclude -I$C $flags -Xptxas -v -shared -o $out $C/v5ela.cu $C/v5ela_mma.cu $C/v5jpeg.cu -lcudart -lpthread 2>&1 \
    | grep -A2 -E "Function properties for _ZN(2v5|3v5m)16ela_fused_kernelILb1ELb0ELb0" | grep -E "Function|registers|spill" | sed -E "s|.*_ZN(2v5\|3v5m)16ela.*|  build \1 <FAST>:|" | sed "s|^|$out |"
done
