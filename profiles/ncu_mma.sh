#!/bin/bash
# per-launch timing + tensor-pipe activity of the tcgen05 conv kernel in steady state (M=200): one full cell step
TAG=${1:-mma}; SKIP=${2:-640}; CNT=${3:-15}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,launch__shared_mem_per_block_dynamic,lts__t_bytes.sum,sm__inst_executed.sum \
  --clock-control none -k regex:k_conv_mma -s $SKIP -c $CNT --csv --log-file gpurun_out/mma_${TAG}.csv \
  python bench.py --steps 1 --warmup 1 --precision f16x3 --no-cpu-baseline > gpurun_out/mma_${TAG}.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/mma_${TAG}.csv') if not l.startswith('=='))]
hdr=rows[0]; i_id=hdr.index('ID'); i_m=hdr.index('Metric Name'); i_v=hdr.index('Metric Value')
d={}
for r in rows[1:]:
    d.setdefault(r[i_id],{})[r[i_m]]=float(r[i_v].replace(',',''))
for k,v in d.items():
    print('%3s  %8.1f us  tensor %5.1f%%  smem %4dK  L2 %7.1f MB  inst %6.1fM' % (k, v['gpu__time_duration.sum']/1e3, v['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'], v['launch__shared_mem_per_block_dynamic']/1024, v['lts__t_bytes.sum']/1e6, v['sm__inst_executed.sum']/1e6))
PY
