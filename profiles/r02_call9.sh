#!/bin/bash
# round 2, call 9: CDNA head on a side stream (A/B), then the evidence captures with the reports reduced to CSV ON THE BOX
# (gpurun returns at most 64 MiB)
mkdir -p gpurun_out
bash profiles/r01_ab.sh r2i "VF_SIDE_CDNA=0"
timeout 600 bash profiles/launch_list.sh r2i
python profiles/summarize_launches.py gpurun_out/launches_r2i.csv 2>/dev/null | head -30
timeout 900 bash profiles/ncu_full.sh r2i 624 14
ncu -i gpurun_out/full_r2i.ncu-rep --page raw --csv > gpurun_out/raw_conv_r2i.csv 2>/dev/null
for i in 1 10 11 12 13; do   # lstm0, dec2, heads0, scratch1, masks1
  ncu -i gpurun_out/full_r2i.ncu-rep --page source --csv --print-source cuda,sass --launch-skip $i --launch-count 1 > gpurun_out/src_conv${i}_r2i.csv 2>/dev/null
done
rm -f gpurun_out/full_r2i.ncu-rep
timeout 900 bash profiles/ncu_pointwise.sh r2i 920 48
ncu -i gpurun_out/pw_r2i.ncu-rep --page raw --csv > gpurun_out/raw_pw_r2i.csv 2>/dev/null
for K in k_cdna_apply4s k_composite4s k_norm_act k_upsample2x k_lstm_gates k_lstm_out k_cdna_partial k_plane_stats; do
  ncu -i gpurun_out/pw_r2i.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:$K --launch-count 1 > gpurun_out/src_${K}_r2i.csv 2>/dev/null
done
rm -f gpurun_out/pw_r2i.ncu-rep
du -sh gpurun_out
