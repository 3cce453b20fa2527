#!/bin/bash
# hardware experiment: which shifted-descriptor operand layouts are (a) correct, (b) fast
for mode in "1 0" "1 1" "2 0" "2 1" "0 0"; do
  set -- $mode
  export VF_MMA_LAYOUT=$1 VF_MMA_BO=$2
  echo "=== layout=$1 bo=$2"
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "conv_mma and not impl2" 2>&1 | grep -E "^E  .*err|passed|failed" | head -5
  if timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "conv_mma" >/dev/null 2>&1; then
    timeout 300 python bench.py --steps 2 --warmup 1 --precision f16x3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('   ms_per_step %.1f  lstm_conv TF/s %.1f  lstm_ms %.1f other_conv_ms %.1f' % (d['ms_per_step'], r['achieved'], r['ms_per_launch']*r['launches'], r['other_conv_ms']))"
  fi
done
