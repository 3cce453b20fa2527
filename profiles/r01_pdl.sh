#!/bin/bash
# PDL on/off comparison + warm-cache launch list
TAG=${1:-pdl}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.log
tail -5 gpurun_out/pytest_${TAG}.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_${TAG}_on.json 2> gpurun_out/bench_${TAG}_on.err; cut -c1-330 gpurun_out/bench_${TAG}_on.json | tail -1
VF_PDL=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_${TAG}_off.json 2> gpurun_out/bench_${TAG}_off.err; cut -c1-330 gpurun_out/bench_${TAG}_off.json | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -s 3200 -c 1100 --csv --log-file gpurun_out/launches_${TAG}_warm.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
