#!/bin/bash
# experiments: the TPS sweep under different staging-box shapes (DVSG_TILE_BOXES = bw0,bh0,bh1,bw2,bh2)
mkdir -p gpurun_out
for boxes in "$@"; do
  echo "=== boxes $boxes"
  DVSG_TILE_BOXES=$boxes timeout 600 python tools/sweep.py tps 2>&1 | grep -E "amp=|tps1080|5x5|\+xy|\+mask|Error|error"
done
