#!/bin/bash
set -x
mkdir -p gpurun_out
python tools/k3_variants.py v1 x2 x3 x4
S=32 ncu --set full --clock-control none --import-source on -k regex:orient_accumulate -s 1 -c 1 -f -o gpurun_out/k3_x2_full python tools/k3_variants.py x2 > gpurun_out/k3_x2_ncu.log 2>&1
S=32 ncu --set full --clock-control none --import-source on -k regex:orient_accumulate -s 1 -c 1 -f -o gpurun_out/k3_x4_full python tools/k3_variants.py x4 > gpurun_out/k3_x4_ncu.log 2>&1
