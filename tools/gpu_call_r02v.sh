#!/bin/bash
# 1 GPU: unit-level expansion of big list records (FB_KEXPAND / FB_EXPAND_LEAVES variants of the trace kernel)
set -u
OUT=gpurun_out
mkdir -p $OUT
for v in default x12 x12b x20; do
  so=fluxpy_b200/libfluxb200_$v.so
  [ $v = default ] && so=fluxpy_b200/libfluxb200.so
  echo "== $v"
  FLUXB200_SO=$PWD/$so PROF_ONE_REPS=4 python tools/prof_one.py 4096 317 2>&1 | grep "^rep\|nnz" | tail -3 | cut -c1-400
done | tee $OUT/r02v_variants.log
