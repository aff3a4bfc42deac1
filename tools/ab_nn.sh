#!/bin/bash
# same-box A/B of library variants on the default wave-1 NN search (tools/probe_nn_search.py: lockstep and pipelined lines)
set -u
LIB=alphazero-reversi_b200/librvs_b200.so
cp $LIB /tmp/librvs_keep.so
for rep in 1 2; do for v in "$@"; do cp build/variants/librvs_$v.so $LIB; echo "== $v: $(timeout 300 python tools/probe_nn_search.py 4096 4 2>&1 | head -2 | cut -c1-75 | tr '\n' '|')"; done; done
cp /tmp/librvs_keep.so $LIB
