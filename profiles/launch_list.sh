#!/bin/bash
TAG=${1:-x}; PREC=${2:-f16x3}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,launch__shared_mem_per_block_dynamic,launch__grid_size --clock-control none -s 3200 -c 1100 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --precision ${PREC} --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
