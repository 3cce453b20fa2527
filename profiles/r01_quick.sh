#!/bin/bash
# quick GPU iteration: parity tests, bench line, launch list  (usage: bash profiles/r01_quick.sh TAG)
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.log
tail -25 gpurun_out/pytest_${TAG}.log
if grep -q "rc=0" gpurun_out/pytest_${TAG}.log; then
  timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err; tail -1 gpurun_out/bench_${TAG}.json; tail -3 gpurun_out/bench_${TAG}.err
  timeout 600 bash profiles/launch_list.sh ${TAG}
fi
