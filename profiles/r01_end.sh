#!/bin/bash
# end-of-round evidence (tag r1i): smoke(), launch list, --set full capture of 16 consecutive conv launches (one cell step),
# bench (ours, with the cpu_baseline leg) and the reference arm
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1i.log 2>&1; tail -n 1 gpurun_out/smoke_r1i.log
timeout 600 bash profiles/launch_list.sh r1i
timeout 900 bash profiles/ncu_full.sh r1i 640 16
timeout 600 python bench.py > gpurun_out/bench_r1i.json 2> gpurun_out/bench_r1i.err; cut -c1-300 gpurun_out/bench_r1i.json | tail -n 1
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1i.json 2>&1; cut -c1-300 gpurun_out/bench_ref_r1i.json | tail -n 1
