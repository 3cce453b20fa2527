#!/bin/bash
# round 2, call 28: end-of-round code incl. the split last round: smoke, driver-style bench line, launch list
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r2z.log 2>&1; tail -1 gpurun_out/smoke_r2z.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_r2z.json 2> gpurun_out/bench_r2z.err; tail -c 400 gpurun_out/bench_r2z.json; tail -2 gpurun_out/bench_r2z.err
timeout 600 bash profiles/launch_list.sh r2z
python profiles/summarize_launches.py gpurun_out/launches_r2z.csv 2>/dev/null | head -8
