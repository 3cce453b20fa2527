#!/bin/bash
# round 2, call 4: fused statistics finalisation + 2x2-block upsample: parity, then A/B
mkdir -p gpurun_out
bash profiles/r01_ab.sh r2d "VF_FUSE_FIN=0" "VF_UPSAMPLE_BLK=0" "VF_FUSE_FIN=0 VF_UPSAMPLE_BLK=0"
