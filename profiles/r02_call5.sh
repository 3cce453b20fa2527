#!/bin/bash
# round 2, call 5: deferred statistics arrival, hoisted sabias, fused conv-LSTM pointwise cluster kernel, dense head buffers
mkdir -p gpurun_out
bash profiles/r01_ab.sh r2e "VF_LSTM_FUSED=0" "VF_HOIST_SA=0" "VF_FUSE_FIN=0" "VF_LSTM_FUSED=0 VF_HOIST_SA=0 VF_FUSE_FIN=0"
timeout 600 bash profiles/launch_list.sh r2e
python profiles/summarize_launches.py gpurun_out/launches_r2e.csv 2>/dev/null | head -30
