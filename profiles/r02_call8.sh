#!/bin/bash
# round 2, call 8: parity, norm_act8 A/B, then the evidence captures of the current code: launch list, --set full of one cell step's
# 14 convolutions and of its pointwise kernels
mkdir -p gpurun_out
bash profiles/r01_ab.sh r2h "VF_NORM_ACT8=1"
timeout 600 bash profiles/launch_list.sh r2h
python profiles/summarize_launches.py gpurun_out/launches_r2h.csv 2>/dev/null | head -30
timeout 900 bash profiles/ncu_full.sh r2h 624 14
timeout 900 bash profiles/ncu_pointwise.sh r2h 920 48
