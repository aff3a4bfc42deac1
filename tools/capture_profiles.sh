#!/bin/bash
# Round-2 evidence run (one B200): plain bench lines first (numbers are never taken under a profiler), then the ncu
# launch list of the same command and one `--set full` capture per dominant kernel.  The reports are exported to
# CSV / text on the box (gpurun brings back at most 64 MiB) into gpurun_out/cap_r2/;
# `python profiles/make_capture.py` turns them into profiles/*_r2.*
#   usage: gpurun --timeout 2400 -- tools/capture_profiles.sh [bench|ncu|k1g|all]
set -u
WHAT=${1:-all}
O=gpurun_out/cap_r2
mkdir -p $O
NCU="ncu --clock-control none"
cap() {  # cap <name> <ncu filter args...> -- <command...>
  local name=$1; shift
  local filt=()
  while [ "$1" != "--" ]; do filt+=("$1"); shift; done
  shift
  $NCU --set full --import-source on "${filt[@]}" -f -o /tmp/$name "$@" > $O/$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/$name.raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip > $O/$name.src.csv.gz
  ncu -i /tmp/$name.ncu-rep --page details > $O/$name.details.txt 2>/dev/null
  rm -f /tmp/$name.ncu-rep
}
if [ "$WHAT" = bench ] || [ "$WHAT" = all ]; then
  python bench.py --steps 20 --warmup 5 > $O/bench_r2_1gpu.json 2> $O/bench_r2_1gpu.err
  python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r2_reference_arm.json 2> $O/bench_r2_reference_arm.err
fi
if [ "$WHAT" = k1g ]; then  # only the wave-1 group kernel changed: refresh its captures (the others stay in $O)
  $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/launches_bench_r2.csv \
      python bench.py --steps 2 --warmup 1 --min-seconds 0 --no-cpu --no-big > $O/launches_bench_r2.log 2>&1
  cap k1g -k regex:selfplay_k1g -s 3 -c 1 -- python bench.py --steps 2 --warmup 3 --steps-per-launch 2 --min-seconds 0 --no-cpu --no-big --no-nn
  cap k1g_16k -k regex:selfplay_k1g -s 1 -c 1 -- python tools/probe_selfplay.py 16384 0 2
fi
if [ "$WHAT" = ncu ] || [ "$WHAT" = all ]; then
  # launch list of the bench command (short: ncu serialises and replays)
  $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/launches_bench_r2.csv \
      python bench.py --steps 2 --warmup 1 --min-seconds 0 --no-cpu --no-big > $O/launches_bench_r2.log 2>&1
  # dominant kernel of the headline: one launch of 2 steps = 8192 game-plies x 100 simulations
  cap k1g -k regex:selfplay_k1g -s 3 -c 1 -- python bench.py --steps 2 --warmup 3 --steps-per-launch 2 --min-seconds 0 --no-cpu --no-big --no-nn
  # the many-games regime (16384 games, automatic lanes per game)
  cap k1g_16k -k regex:selfplay_k1g -s 1 -c 1 -- python tools/probe_selfplay.py 16384 0 2
  # 128 filters: the whole-network kernel (first layer + tower + head planes, one launch per forward) ...
  cap tower128 -k regex:conv_tower_kernel -s 3 -c 1 -- python tools/probe_net.py 5 128 4096 predict
  # ... and the per-layer kernel it replaces (RVS_OPT_NET_TOWER = 0: plain + residual layer); 256 filters
  cap conv128 -k regex:conv3x3_tc2_kernel -s 34 -c 2 -- env RVS_TOWER=0 python tools/probe_net.py 5 128 4096 predict
  cap conv256 -k regex:conv3x3_tc2s -s 122 -c 2 -- python tools/probe_net.py 20 256 4096 predict
  # the small kernels of an NN wave (tree step incl. input tiles, first layer, heads)
  cap nnaux -k "regex:heads_kernel|nn_step_kernel" --launch-skip 300 -c 4 -- python tools/probe_nn_wave.py
  $NCU --metrics gpu__time_duration.sum --launch-skip 800 -c 24 --csv --log-file $O/launches_nnwave_r2.csv \
      python tools/probe_nn_wave.py > /dev/null 2>&1
  # board kernels
  cap board -k "regex:legal_masks|apply_moves" -c 2 -- python tools/probe_board.py 4194304
fi
du -sh $O; ls $O | head -40
