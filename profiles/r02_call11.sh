#!/bin/bash
# round 2, call 11: two issuer warps for the row-stacked thin convs (parity + A/B)
mkdir -p gpurun_out
bash profiles/r01_ab.sh r2k "VF_DUAL_ISSUE=0"
