#!/bin/bash
# round 2, call 1: parity of the new epilogue statistics / merged head conv, A/B of the switches, the two-stream experiment
mkdir -p gpurun_out
bash profiles/r01_ab.sh r2a "VF_EPI_STATS=0" "VF_MERGE_HEADS=0" "VF_STATS_FIN=1" "VF_EPI_STATS=0 VF_MERGE_HEADS=0 VF_STATS_FIN=1" "VF_ENC_S2D=1"
timeout 600 python profiles/r02_twostream.py > gpurun_out/twostream_r2a.json 2> gpurun_out/twostream_r2a.err; tail -1 gpurun_out/twostream_r2a.json; tail -3 gpurun_out/twostream_r2a.err
timeout 600 bash profiles/launch_list.sh r2a
python profiles/summarize_launches.py gpurun_out/launches_r2a.csv 2>/dev/null | head -40
