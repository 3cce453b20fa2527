#!/bin/bash
# session-2 state check: GPU parity tests, bench line, ncu launch list, tensor-pipe activity of the conv kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_s2.json 2> gpurun_out/bench_s2.err; tail -1 gpurun_out/bench_s2.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_s2.json 2>&1; tail -1 gpurun_out/bench_ref_s2.json
timeout 600 bash profiles/launch_list.sh r1e
timeout 600 bash profiles/ncu_mma.sh r1e 640 30
