#!/bin/bash
# full GPU parity suite, then experiment: SWIZZLE_128B operand layout (64-channel chunks, half as many weight stages) vs SWIZZLE_64B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_full.log 2>&1; tail -3 gpurun_out/pytest_full.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_layout1.json 2> gpurun_out/bench_layout1.err; cut -c1-330 gpurun_out/bench_layout1.json | tail -1
export VF_MMA_LAYOUT=2
timeout 600 python -m pytest tests -m gpu -x -q -k "conv_mma or tensor_core or full_horizon" > gpurun_out/pytest_layout2.log 2>&1; tail -3 gpurun_out/pytest_layout2.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_layout2.json 2> gpurun_out/bench_layout2.err; cut -c1-330 gpurun_out/bench_layout2.json | tail -1; tail -2 gpurun_out/bench_layout2.err
