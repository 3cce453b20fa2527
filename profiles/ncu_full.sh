#!/bin/bash
# --set full capture (with source) of COUNT consecutive k_conv_mma launches in steady state: one cell step = 15 launches
# in the order scratch0, scratch1, masks0, masks1, enc0, lstm0, enc1, lstm1, enc2, lstm2, dec0, lstm3, dec1, lstm4, dec2 (skip 640)
TAG=${1:-full}; SKIP=${2:-640}; CNT=${3:-15}
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:k_conv_ -s $SKIP -c $CNT -o gpurun_out/full_${TAG} -f \
  python bench.py --steps 1 --warmup 1 --precision f16x3 --no-cpu-baseline > gpurun_out/full_${TAG}.log 2>&1
ls -la gpurun_out/full_${TAG}*
