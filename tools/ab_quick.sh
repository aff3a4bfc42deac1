#!/bin/bash
# same-box A/B of library variants (build/variants/librvs_<name>.so), interleaved twice: boxes differ by a few percent
# usage: tools/ab_quick.sh variantA variantB ...
set -u
LIB=alphazero-reversi_b200/librvs_b200.so
cp $LIB /tmp/librvs_keep.so
for rep in 1 2; do
for v in "$@"; do
  cp build/variants/librvs_$v.so $LIB
  echo "== $v: $(timeout 300 python tools/probe_net.py 5 128 4096 predict | tail -1) | $(timeout 300 python tools/probe_nn_search.py 4096 4 2>&1 | head -1)"
done
done
cp /tmp/librvs_keep.so $LIB
