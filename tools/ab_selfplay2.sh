#!/bin/bash
# interleaved same-box A/B of library variants on the wave-1 self-play probe (4096 games, 8 lanes; 16384 games):
#   tools/ab_selfplay2.sh variantA variantB ...     (build/variants/librvs_<name>.so; product library restored at the end)
set -u
LIB=alphazero-reversi_b200/librvs_b200.so
cp $LIB /tmp/librvs_keep.so
for rep in 1 2 3; do
for v in "$@"; do
  cp build/variants/librvs_$v.so $LIB
  echo "== $v: $(python tools/probe_selfplay.py 4096 8 40 | tail -1) | $(python tools/probe_selfplay.py 16384 0 10 | tail -1)"
done
done
cp /tmp/librvs_keep.so $LIB
