#!/bin/bash
# Development aid: build tuning variants of libv5ela.so (threads per CTA, strip width in MCUs, resident CTAs per SM).
# usage: profiles/build_variants.sh "NT,TW,CTAS[,extra nvcc flags]" ...   -> profiles/variants/libv5ela_NT_TW_CTAS.so
set -e
cd "$(dirname "$0")/.."
mkdir -p profiles/variants
for v in "$@"; do
  IFS=, read -r nt tw ctas extra <<< "$v"
  out=profiles/variants/libv5ela_${nt}_${tw}_${ctas}.so
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
       -Iinclude -Ifake-video-detection-engine_b200/csrc -DV5_NT=$nt -DV5_TW_MAX=$tw -DV5_MIN_CTAS=$ctas $extra \
       -Xptxas -v -shared -o $out fake-video-detection-engine_b200/csrc/v5ela.cu -lcudart 2>&1 | grep -A1 "ela_fused" | grep -E "registers|spill" | sed "s|^|$out: |"
done
