#!/bin/bash
# round 2, call 6: device sampler options (correlated noise, discrete_ind, append_action) under test; 64-channel layers on the wide tiling A/B
mkdir -p gpurun_out
bash profiles/r01_ab.sh r2f "VF_THIN_MAX_COUT=32" "VF_HOIST_SA=0"
VF_THIN_MAX_COUT=32 timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2f_wide64.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2f_wide64.log
tail -5 gpurun_out/pytest_r2f_wide64.log
