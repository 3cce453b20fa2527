#!/bin/bash
# round 2, call 16: final evidence of the end-of-round code (pair kernel on): driver-style bench lines, launch list, --set full of one
# cell step's convolutions (reports reduced to CSV on the box), smoke
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2n.log 2>&1; tail -2 gpurun_out/smoke_r2n.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_r2n.json 2> gpurun_out/bench_r2n.err; tail -c 600 gpurun_out/bench_r2n.json; tail -2 gpurun_out/bench_r2n.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref_r2n.json 2> gpurun_out/bench_ref_r2n.err; tail -c 900 gpurun_out/bench_ref_r2n.json
timeout 600 bash profiles/launch_list.sh r2n
python profiles/summarize_launches.py gpurun_out/launches_r2n.csv 2>/dev/null | head -12
timeout 900 bash profiles/ncu_full.sh r2n 624 14
ncu -i gpurun_out/full_r2n.ncu-rep --page raw --csv > gpurun_out/raw_conv_r2n.csv 2>/dev/null
ncu -i gpurun_out/full_r2n.ncu-rep --page source --csv --print-source cuda,sass --launch-skip 3 --launch-count 1 > gpurun_out/src_conv3_r2n.csv 2>/dev/null
rm -f gpurun_out/full_r2n.ncu-rep
du -sh gpurun_out
