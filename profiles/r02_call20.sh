#!/bin/bash
# round 2, call 20: hi/lo weight halves of the 3x3 Cout <= 16 row-stacked convs as ONE B operand (VF_STACK_NP: 2 MMAs per K step
# instead of 3 for scratch1 / masks1): parity + bench A/B + warm per-layer times
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_r2r.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2r.log
tail -4 gpurun_out/pytest_r2r.log
i=0
for E in "VF_STACK_NP=16" "VF_STACK_NP=0" "VF_STACK_NP=16" "VF_STACK_NP=0"; do
  env $E timeout 600 python bench.py --no-cpu-baseline --no-strong --steps 10 > gpurun_out/bench_r2r_$i.json 2> gpurun_out/bench_r2r_$i.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_r2r_$i.json').read().strip().splitlines()[-1]); r=d['roofline']
    print('$E', 'ms/plan %.2f e2e %.0f gate ms/launch %.4f other_conv_ms %.2f' % (d['ms_per_step'], d['e2e']['value'], r['ms_per_launch'], r['other_conv_ms']), d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print('$E failed', e)
PY
  i=$((i+1))
done
for E in "VF_STACK_NP=16" "VF_STACK_NP=0"; do echo "--- $E"; env $E timeout 300 python profiles/conv_microbench.py 2>&1 | grep -E "scratch1|masks1|dec2"; done
