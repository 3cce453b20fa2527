#!/bin/bash
# round 2, call 2: pool-fused encoder convs under test, then a --set full + source capture of one cell step's 14 convolutions
mkdir -p gpurun_out
VF_ENC_S2D=1 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2b_s2d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2b_s2d.log
tail -5 gpurun_out/pytest_r2b_s2d.log
timeout 1500 bash profiles/ncu_full.sh r2b 624 14
