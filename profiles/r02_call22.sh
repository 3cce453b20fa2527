#!/bin/bash
# round 2, call 22 (8 GPUs): end-of-round code — bit-identity of the sharded planner / policy on 8 ranks, bench at N=8 (weak value + strong legs)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 tests/multigpu_check.py > gpurun_out/multigpu_check_n8_r2t.log 2>&1; tail -12 gpurun_out/multigpu_check_n8_r2t.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --steps 5 --warmup 3 \
   > gpurun_out/bench_r2t_n8.json 2> gpurun_out/bench_r2t_n8.err
tail -c 3000 gpurun_out/bench_r2t_n8.json; tail -3 gpurun_out/bench_r2t_n8.err
