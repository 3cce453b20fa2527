#!/bin/bash
# --set full + source capture of the HBM / latency-bound kernels of one steady-state cell step (everything except k_conv_mma)
TAG=${1:-pw}; SKIP=${2:-920}; CNT=${3:-48}
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none \
  -k regex:'k_cdna|k_composite|k_norm_act|k_upsample|k_lstm|k_plane|k_stats|k_sabias|k_distrib|k_pack|k_build' -s $SKIP -c $CNT \
  -o gpurun_out/pw_${TAG} -f python bench.py --steps 1 --warmup 1 --precision f16x3 --no-cpu-baseline --no-strong > gpurun_out/pw_${TAG}.log 2>&1
ls -la gpurun_out/pw_${TAG}*
