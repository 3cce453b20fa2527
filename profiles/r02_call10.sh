#!/bin/bash
# round 2, call 10: warp-uniform / 32-bit MMA issue in the thin paths (parity + bench), optional c3 and c5 bench lines
mkdir -p gpurun_out
bash profiles/r01_ab.sh r2j
timeout 900 python bench.py --config c3 --steps 3 --warmup 2 > gpurun_out/bench_r2j_c3.json 2> gpurun_out/bench_r2j_c3.err; tail -c 1800 gpurun_out/bench_r2j_c3.json; tail -3 gpurun_out/bench_r2j_c3.err
timeout 900 python bench.py --config c5 --steps 2 --warmup 2 > gpurun_out/bench_r2j_c5.json 2> gpurun_out/bench_r2j_c5.err; tail -c 1800 gpurun_out/bench_r2j_c5.json; tail -3 gpurun_out/bench_r2j_c5.err
