#!/bin/bash
# DRAM traffic / throughput of every kernel of ~6 cell steps in steady state (320 launches: the DRAM counters need several
# replay passes per kernel, ~0.6 s per launch under ncu — 900 launches did not fit a 220 s call) (the HBM-bound pointwise, CDNA, cost and CEM
# kernels): time, DRAM bytes read + written, DRAM and SM throughput as % of peak
TAG=${1:-hbm}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -s 3200 -c 320 --csv --log-file gpurun_out/hbm_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_hbm_${TAG}.log 2>&1
ls -la gpurun_out/hbm_${TAG}.csv
