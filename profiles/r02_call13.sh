#!/bin/bash
mkdir -p gpurun_out
VF_CTA_PAIR=0 timeout 120 python profiles/r02_pair_check.py > gpurun_out/pair_off.log 2>&1; tail -6 gpurun_out/pair_off.log
VF_CTA_PAIR=1 timeout 120 python profiles/r02_pair_check.py > gpurun_out/pair_on.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pair_on.log
