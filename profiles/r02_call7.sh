#!/bin/bash
# round 2, call 7 (2 GPUs): the engine's peer exchange across two devices, the policy surface on 2 ranks, bench at N=2 (peer and NCCL arms)
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_parity.py::test_two_gpu_sharded_plan_is_bit_identical tests/test_gpu_parity_r2.py -m gpu -q > gpurun_out/pytest_r2g_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2g_2gpu.log
tail -15 gpurun_out/pytest_r2g_2gpu.log
for COLL in peer nccl; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --collective $COLL \
     > gpurun_out/bench_r2g_n2_$COLL.json 2> gpurun_out/bench_r2g_n2_$COLL.err
  tail -c 2500 gpurun_out/bench_r2g_n2_$COLL.json; tail -3 gpurun_out/bench_r2g_n2_$COLL.err
done
