#!/bin/bash
# k-means assignment kernel: per-warp tile size (boxes) x ring depth, 50 M boxes, k = 9.
# 256 x 2 and 128 x 4 use the same shared memory (2 CTAs/SM); larger rings drop to 1 CTA/SM.
for cfg in "256 2" "128 2" "128 3" "128 4" "64 4" "512 2"; do
  set -- $cfg
  echo "tile $1 stages $2"
  YB_KM_TILE=$1 YB_KM_STAGES=$2 python benchmarks/bench_configs.py kmeans 2>&1 | grep kmeans_50M | cut -c1-200
done
