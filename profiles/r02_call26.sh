#!/bin/bash
# round 2, call 26 (2 GPUs): end-of-round code — 2-GPU parity tests (sharded plan / policy bit-identical) and the bench at N=2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py::test_two_gpu_sharded_plan_is_bit_identical -m gpu -q > gpurun_out/pytest_r2x_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2x_2gpu.log
tail -4 gpurun_out/pytest_r2x_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 \
   > gpurun_out/bench_r2x_n2.json 2> gpurun_out/bench_r2x_n2.err
tail -c 2500 gpurun_out/bench_r2x_n2.json; tail -3 gpurun_out/bench_r2x_n2.err
