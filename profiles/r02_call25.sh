#!/bin/bash
# round 2, call 25: end-of-round code (pair kernel, column strips for the 32-wide gate convs and enc2, stacked weight halves for masks1): full parity suite, smoke,
# driver-style bench line + reference arm, launch list, --set full of one cell step's convolutions reduced to CSV on the box
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_r2w.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2w.log
tail -4 gpurun_out/pytest_r2w.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2w.log 2>&1; tail -2 gpurun_out/smoke_r2w.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_r2w.json 2> gpurun_out/bench_r2w.err; tail -c 600 gpurun_out/bench_r2w.json; tail -2 gpurun_out/bench_r2w.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/bench_ref_r2w.json 2> gpurun_out/bench_ref_r2w.err; tail -c 900 gpurun_out/bench_ref_r2w.json
timeout 600 bash profiles/launch_list.sh r2w
python profiles/summarize_launches.py gpurun_out/launches_r2w.csv 2>/dev/null | head -14
timeout 900 bash profiles/ncu_full.sh r2w 624 14
ncu -i gpurun_out/full_r2w.ncu-rep --page raw --csv > gpurun_out/raw_conv_r2w.csv 2>/dev/null
rm -f gpurun_out/full_r2w.ncu-rep
du -sh gpurun_out
