#!/bin/bash
# A/B of library variants on the network path: numerics (tests/test_gpu_net.py) + forward / NN-search throughput
# usage: tools/ab_net.sh variantA variantB ...   (build/variants/librvs_<name>.so; the product library is restored)
set -u
LIB=alphazero-reversi_b200/librvs_b200.so
cp $LIB /tmp/librvs_keep.so
for v in "$@"; do
  cp build/variants/librvs_$v.so $LIB
  echo "== $v numerics"; timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -x -q 2>&1 | tail -3
  echo "== $v forward 5x128"; timeout 300 python tools/probe_net.py 5 128 4096 predict | tail -1
  echo "== $v forward 2x64"; timeout 300 python tools/probe_net.py 2 64 4096 predict | tail -1
  echo "== $v forward 20x256"; timeout 300 python tools/probe_net.py 20 256 4096 predict | tail -1
  echo "== $v NN search"; timeout 300 python tools/probe_nn_search.py 4096 2 2>&1 | head -1
done
cp /tmp/librvs_keep.so $LIB
