#!/bin/bash
mkdir -p gpurun_out
VF_CTA_PAIR=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_conv_pair -s 4 -c 1 -o gpurun_out/pair_prof -f python profiles/r02_pair_check.py > gpurun_out/pair_prof.log 2>&1
ncu -i gpurun_out/pair_prof.ncu-rep --page raw --csv > gpurun_out/pair_raw.csv 2>/dev/null
ncu -i gpurun_out/pair_prof.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/pair_src.csv 2>/dev/null
rm -f gpurun_out/pair_prof.ncu-rep
tail -3 gpurun_out/pair_prof.log; ls -la gpurun_out/pair_*
