#!/bin/bash
# round 2, call 15: cta_group::2 pair kernel for the 16x16 gate convs: full parity suite + bench A/B
mkdir -p gpurun_out
VF_CTA_PAIR=1 timeout 1500 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_r2m_pair.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2m_pair.log
tail -6 gpurun_out/pytest_r2m_pair.log
for V in "VF_CTA_PAIR=1" "VF_CTA_PAIR=0"; do
  env $V timeout 600 python bench.py --no-cpu-baseline --no-strong > gpurun_out/bench_r2m_$V.json 2> gpurun_out/bench_r2m_$V.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2m_$V.json').read().strip().splitlines()[-1]); r=d['roofline']
print('$V', 'ms/plan %.2f frames/s %.0f gate ms/launch %.4f frac %.3f other_conv_ms %.2f' % (d['ms_per_step'], d['value'], r['ms_per_launch'], r['frac'], r['other_conv_ms']), d['clocks'])
PY
done
