#!/bin/bash
# round 2, call 19: programmatic dependent launch with the early trigger ONLY in the one-wave convolution kernels (pdlconv) or
# nowhere (pdllate) vs plain stream order; new pair-kernel shapes of the conv test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "conv_mma" > gpurun_out/pytest_r2q_conv.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2q_conv.log
tail -3 gpurun_out/pytest_r2q_conv.log
V=visual_foresight_b200/_variants
i=0
for E in "VF_PDL=0" "VF_PDL=1 VF_ENGINE_LIB=$V/libvfengine_pdlconv.so" "VF_PDL=1 VF_ENGINE_LIB=$V/libvfengine_pdllate.so" "VF_PDL=1" "VF_PDL=0"; do
  env $E timeout 600 python bench.py --no-cpu-baseline --no-strong --steps 10 > gpurun_out/bench_r2q_$i.json 2> gpurun_out/bench_r2q_$i.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_r2q_$i.json').read().strip().splitlines()[-1]); r=d['roofline']
    print('$E', 'ms/plan %.2f e2e %.0f gate ms/launch %.4f other_conv_ms %.2f' % (d['ms_per_step'], d['e2e']['value'], r['ms_per_launch'], r['other_conv_ms']), d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print('$E failed', e)
PY
  i=$((i+1))
done
