#!/bin/bash
# round 2, call 17: --set full of one cell step's 14 convolutions INCLUDING the pair kernel (regex k_conv_), reduced to CSV on the box
mkdir -p gpurun_out
timeout 900 bash profiles/ncu_full.sh r2o 624 14
ncu -i gpurun_out/full_r2o.ncu-rep --page raw --csv > gpurun_out/raw_conv_r2o.csv 2>/dev/null
ncu -i gpurun_out/full_r2o.ncu-rep --page source --csv --print-source cuda,sass --launch-skip 3 --launch-count 1 > gpurun_out/src_conv3_r2o.csv 2>/dev/null
rm -f gpurun_out/full_r2o.ncu-rep
du -sh gpurun_out
