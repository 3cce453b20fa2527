#!/bin/bash
# Run under gpurun (1 GPU).  Produces gpurun_out/launches_<tag>.csv (every launch with its device time, cold-cache
# and serialised: compare SHARES) and gpurun_out/prof_<tag>.ncu-rep (--set full on the dominant kernel).
TAG=${1:-r1}
PREC=${2:-f16x3}
mkdir -p gpurun_out
# one CEM iteration of c2 = 14 cell steps ~ 1000 launches; skip the first (cold) plan
ncu --metrics gpu__time_duration.sum --clock-control none -s 3200 -c 1100 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --precision ${PREC} --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_conv_mma -s 20 -c 3 -o gpurun_out/prof_${TAG} \
    python bench.py --steps 1 --warmup 1 --precision ${PREC} --no-cpu-baseline >> gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ls -la gpurun_out
