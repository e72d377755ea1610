#!/bin/bash
# A/B: the TPS sweep under alternative builds of the same sources (build_ab/lib_<tag>.so)
for tag in "$@"; do
  echo "=== $tag"
  DVSG_LIB=$PWD/build_ab/lib_$tag.so timeout 600 python tools/sweep.py tps 2>&1 | grep -E "amp=0.20 tile|amp=0.00|tps1080|5x5|\+xy|Error|error"
done
