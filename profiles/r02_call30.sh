#!/bin/bash
# round 2, call 30: staged split-K kernel of the CDNA kernel head (VF_CDNA_STAGED): parity + A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_r2aa.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2aa.log
tail -4 gpurun_out/pytest_r2aa.log
i=0
for E in "VF_CDNA_STAGED=1" "VF_CDNA_STAGED=0" "VF_CDNA_STAGED=1" "VF_CDNA_STAGED=0"; do
  env $E timeout 300 python bench.py --no-cpu-baseline --no-strong --steps 10 > gpurun_out/bench_r2aa_$i.json 2> gpurun_out/bench_r2aa_$i.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_r2aa_$i.json').read().strip().splitlines()[-1]); r=d['roofline']
    print('$E', 'ms/plan %.2f e2e %.0f gate ms/launch %.4f other_conv_ms %.2f' % (d['ms_per_step'], d['e2e']['value'], r['ms_per_launch'], r['other_conv_ms']), d['clocks']['sm_mhz'])
except Exception as e:
    print('$E failed', e)
PY
  i=$((i+1))
done
