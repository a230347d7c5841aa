#!/bin/bash
# Runs on the GPU box (under gpurun): launch list of the bench command + one full ncu capture of K3 and of K2 (S=1).
# Outputs land in gpurun_out/ ; tools/summarize_profiles.py turns them into the committed summaries under profiles/.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --samples-per-rank 64 --no-cpu-baseline --no-hoi > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:orient_accumulate -s 1 -c 1 -f -o gpurun_out/k3_full \
    python bench.py --steps 1 --warmup 1 --samples-per-rank 32 --no-cpu-baseline --no-hoi > gpurun_out/k3_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pair_accumulate -s 12 -c 1 -f -o gpurun_out/k2_full \
    python bench.py --steps 1 --warmup 1 --samples-per-rank 32 --no-cpu-baseline --no-hoi > gpurun_out/k2_ncu.log 2>&1
ls -la gpurun_out
