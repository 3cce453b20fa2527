#!/bin/bash
# round 2, call 29: optional bench lines of BASELINE configs c3 and c5 (one 1/8 shard) with the end-of-round code
mkdir -p gpurun_out
timeout 600 python bench.py --config c3 --steps 3 --warmup 2 > gpurun_out/bench_r2z_c3.json 2> gpurun_out/bench_r2z_c3.err; tail -c 700 gpurun_out/bench_r2z_c3.json; tail -2 gpurun_out/bench_r2z_c3.err
timeout 600 python bench.py --config c5 --steps 2 --warmup 2 > gpurun_out/bench_r2z_c5.json 2> gpurun_out/bench_r2z_c5.err; tail -c 700 gpurun_out/bench_r2z_c5.json; tail -2 gpurun_out/bench_r2z_c5.err
