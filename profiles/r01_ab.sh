#!/bin/bash
# A/B of one change set: parity tests, then the bench line per environment variant  (usage: bash profiles/r01_ab.sh TAG "VAR=val ..." ...)
TAG=${1:-ab}; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${TAG}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${TAG}.log
tail -n 25 gpurun_out/pytest_${TAG}.log
i=0
for V in "" "$@"; do
  env $V timeout 600 python bench.py --no-cpu-baseline --no-strong > gpurun_out/bench_${TAG}_$i.json 2> gpurun_out/bench_${TAG}_$i.err
  echo "variant $i [$V]: $(python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_${TAG}_$i.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('ms/plan %.2f  frames/s %.0f  gate ms/launch %.4f  other_conv_ms %.2f  clocks %s' % (d['ms_per_step'], d['value'], r['ms_per_launch'], r['other_conv_ms'], d['clocks']))
except Exception as e:
    print('FAILED', e)
PY
)"
  tail -n 2 gpurun_out/bench_${TAG}_$i.err
  i=$((i+1))
done
