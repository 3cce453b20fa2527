#!/bin/bash
# 2-GPU check: sharded plan bit-identical to the single-GPU plan, pytest 2-GPU test, weak-scaling bench lines at N=1,2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/mg_smi.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/multigpu_check.py > gpurun_out/mg_check.log 2>&1; echo "rc=$?" >> gpurun_out/mg_check.log; tail -4 gpurun_out/mg_check.log
timeout 600 python -m pytest tests -m gpu -x -q -k "two_gpu" > gpurun_out/mg_pytest.log 2>&1; tail -3 gpurun_out/mg_pytest.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/mg_bench_n1.json 2> gpurun_out/mg_bench_n1.err; cut -c1-300 gpurun_out/mg_bench_n1.json | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/mg_bench_n2.json 2> gpurun_out/mg_bench_n2.err; tail -1 gpurun_out/mg_bench_n2.json | cut -c1-300
