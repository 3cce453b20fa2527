#!/bin/bash
# round 2, call 27: split last round of the column-strip tiling (VF_TAIL_SPLIT): parity + A/B + warm layer times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "conv_mma" > gpurun_out/pytest_r2y_conv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2y_conv.log
tail -3 gpurun_out/pytest_r2y_conv.log
for E in "VF_TAIL_SPLIT=1" "VF_TAIL_SPLIT=0"; do echo "--- $E"; env $E timeout 300 python profiles/conv_microbench.py 2>&1 | grep -E "lstm0|lstm4|enc2"; done
timeout 1500 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_r2y.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2y.log
tail -4 gpurun_out/pytest_r2y.log
i=0
for E in "VF_TAIL_SPLIT=1" "VF_TAIL_SPLIT=0" "VF_TAIL_SPLIT=1" "VF_TAIL_SPLIT=0"; do
  env $E timeout 600 python bench.py --no-cpu-baseline --no-strong --steps 10 > gpurun_out/bench_r2y_$i.json 2> gpurun_out/bench_r2y_$i.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_r2y_$i.json').read().strip().splitlines()[-1]); r=d['roofline']
    print('$E', 'ms/plan %.2f e2e %.0f gate ms/launch %.4f frac %.3f other_conv_ms %.2f' % (d['ms_per_step'], d['e2e']['value'], r['ms_per_launch'], r['frac'], r['other_conv_ms']), d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print('$E failed', e)
PY
  i=$((i+1))
done
