#!/bin/bash
# end-of-round evidence: smoke(), launch list, --set full capture of 10 conv launches, bench (ours + reference arm), layout experiment
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r1h.log 2>&1; tail -1 gpurun_out/smoke_r1h.log
timeout 600 bash profiles/launch_list.sh r1h
timeout 900 bash profiles/ncu_full.sh r1h 640 10
timeout 600 python bench.py > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; cut -c1-300 gpurun_out/bench_r1h.json | tail -1
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1h.json 2>&1; cut -c1-300 gpurun_out/bench_ref_r1h.json | tail -1
VF_MMA_LAYOUT=2 timeout 300 python -m pytest tests -m gpu -x -q -k "conv_mma or tensor_core or full_horizon" 2>&1 | tail -2
VF_MMA_LAYOUT=2 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_layout2.json 2> gpurun_out/bench_layout2.err; cut -c1-300 gpurun_out/bench_layout2.json | tail -1; tail -2 gpurun_out/bench_layout2.err
