#!/bin/bash
# per-launch timing + tensor-pipe activity of the tcgen05 conv kernel in steady state (M=200)
TAG=${1:-mma}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,launch__grid_size,launch__shared_mem_per_block_dynamic,lts__t_bytes.sum,smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct,l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum \
  --clock-control none -k regex:k_conv_mma -s 120 -c 12 --csv --log-file gpurun_out/mma_${TAG}.csv \
  python bench.py --steps 1 --warmup 1 --precision f16x3 --no-cpu-baseline > gpurun_out/mma_${TAG}.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/mma_${TAG}.csv') if not l.startswith('=='))]
hdr=rows[0]; i_id=hdr.index('ID'); i_m=hdr.index('Metric Name'); i_v=hdr.index('Metric Value')
d={}
for r in rows[1:]:
    d.setdefault(r[i_id],{})[r[i_m]]=r[i_v]
for k,v in d.items():
    print(k, {a.split('.')[0][-28:]:b for a,b in v.items()})
PY
