#!/bin/bash
# round 2, call 3: full GPU suite incl. the round-2 parity tests, bench with the strong-scaling leg, per-tap thin convs A/B
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=8 > gpurun_out/pytest_r2c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2c.log
tail -40 gpurun_out/pytest_r2c.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -c 3000 gpurun_out/bench_r2c.json; tail -5 gpurun_out/bench_r2c.err
VF_THIN_PERTAP=1 timeout 600 python bench.py --no-cpu-baseline --no-strong > gpurun_out/bench_r2c_pertap.json 2> gpurun_out/bench_r2c_pertap.err; tail -c 1500 gpurun_out/bench_r2c_pertap.json; tail -3 gpurun_out/bench_r2c_pertap.err
timeout 600 bash profiles/launch_list.sh r2c
python profiles/summarize_launches.py gpurun_out/launches_r2c.csv 2>/dev/null | head -30
