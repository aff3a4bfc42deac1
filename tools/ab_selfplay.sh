#!/bin/bash
# A/B of library variants on the wave-1 self-play probe: tools/ab_selfplay.sh variantA variantB ...
# (variants are build/variants/librvs_<name>.so; the product library is restored at the end)
set -u
LIB=alphazero-reversi_b200/librvs_b200.so
cp $LIB /tmp/librvs_keep.so
for v in "$@"; do
  cp build/variants/librvs_$v.so $LIB
  for rep in 1 2; do
    echo "== $v 4096 lpg8 (rep $rep)"; python tools/probe_selfplay.py 4096 8 40 | tail -1
  done
  echo "== $v 4096 auto"; python tools/probe_selfplay.py 4096 0 40 | tail -1
  echo "== $v 16384 auto"; python tools/probe_selfplay.py 16384 0 10 | tail -1
done
cp /tmp/librvs_keep.so $LIB
