#!/bin/bash
# 1 GPU: the N3 kernels (block extraction, thin products) on the device-resident 50k-face matrix: times + ncu
set -u
OUT=gpurun_out
mkdir -p $OUT
python tools/prof_n3.py 2>$OUT/r02u_n3.err | tail -1 | tee $OUT/r02u_n3.json
ncu --set full --clock-control none --import-source on -k regex:"extract|csr_matmat|csr_rmatmat" -s 6 -c 6 -o $OUT/r02u_n3 \
    python tools/prof_n3.py > $OUT/r02u_n3_under_ncu.log 2>&1
ls -la $OUT | tail -4
