#!/bin/bash
# same as ab_selfplay2.sh for the many-games regimes: 16384 games (4 lanes per game) and 65536 games (2 lanes)
set -u
LIB=alphazero-reversi_b200/librvs_b200.so
cp $LIB /tmp/librvs_keep.so
for rep in 1 2 3; do
for v in "$@"; do
  cp build/variants/librvs_$v.so $LIB
  echo "== $v: $(python tools/probe_selfplay.py 16384 0 10 | tail -1) | $(python tools/probe_selfplay.py 65536 0 4 | tail -1) | $(python tools/probe_selfplay.py 4096 4 20 | tail -1)"
done
done
cp /tmp/librvs_keep.so $LIB
