#!/bin/bash
# round 2, call 18: 16-byte conv-LSTM pointwise kernels (VF_LSTM_PW bits), consumer-side finalisation for planes with few
# partial slots (VF_FIN_OTF_SLOTS), programmatic dependent launch re-test (VF_PDL), per-layer warm conv times, row groups /
# CTA pairs on the 12x16 maps of 48x64 inputs (VF_RG2_ANYH: c3 bench A/B)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_r2p.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2p.log
tail -4 gpurun_out/pytest_r2p.log
if ! grep -q "pytest rc=0" gpurun_out/pytest_r2p.log; then
  VF_RG2_ANYH=0 timeout 1500 python -m pytest tests -m gpu -q --maxfail=5 > gpurun_out/pytest_r2p_norg.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2p_norg.log
  echo "--- with VF_RG2_ANYH=0:"; tail -4 gpurun_out/pytest_r2p_norg.log
fi
i=0
for V in "VF_LSTM_PW=3" "VF_LSTM_PW=0" "VF_LSTM_PW=1" "VF_LSTM_PW=2" "VF_FIN_OTF_SLOTS=1" "VF_FIN_OTF_SLOTS=2" "VF_FIN_OTF_SLOTS=4" "VF_FIN_OTF_SLOTS=8" "VF_PDL=1" "VF_LSTM_PW=3"; do
  env $V timeout 600 python bench.py --no-cpu-baseline --no-strong --steps 10 > gpurun_out/bench_r2p_$i.json 2> gpurun_out/bench_r2p_$i.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_r2p_$i.json').read().strip().splitlines()[-1]); r=d['roofline']
    print('$V', 'ms/plan %.2f e2e %.0f gate ms/launch %.4f other_conv_ms %.2f' % (d['ms_per_step'], d['e2e']['value'], r['ms_per_launch'], r['other_conv_ms']), d['clocks']['sm_mhz'], d['clocks']['reasons'])
except Exception as e:
    print('$V failed', e)
PY
  i=$((i+1))
done
timeout 300 python profiles/conv_microbench.py > gpurun_out/conv_microbench_r2p.txt 2>&1; cat gpurun_out/conv_microbench_r2p.txt
for V in "VF_RG2_ANYH=1" "VF_RG2_ANYH=0"; do
  env $V timeout 600 python bench.py --config c3 --steps 3 --warmup 2 > gpurun_out/bench_r2p_c3_$V.json 2> gpurun_out/bench_r2p_c3_$V.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_r2p_c3_$V.json').read().strip().splitlines()[-1])
    print('c3 $V', 'ms/plan %.2f value %.0f' % (d['ms_per_step'], d['value']))
except Exception as e:
    print('c3 $V failed', e)
PY
done
