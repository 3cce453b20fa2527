#!/bin/bash
TAG=${1:-src}; SKIP=${2:-652}
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:k_conv_mma -s $SKIP -c 1 -o gpurun_out/src_${TAG} -f \
  python bench.py --steps 1 --warmup 1 --precision f16x3 --no-cpu-baseline > gpurun_out/src_${TAG}.log 2>&1
ls -la gpurun_out/src_${TAG}*
