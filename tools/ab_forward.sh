LIB=alphazero-reversi_b200/librvs_b200.so
cp $LIB /tmp/librvs_keep.so
for rep in 1 2; do for v in "$@"; do cp build/variants/librvs_$v.so $LIB; echo "== $v: $(timeout 120 python tools/probe_net.py 5 128 4096 predict | tail -1)"; done; done
cp /tmp/librvs_keep.so $LIB
