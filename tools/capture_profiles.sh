#!/bin/bash
# Round-2 evidence run (one B200): plain bench lines first (numbers are never taken under a profiler), then the ncu
# launch list of the same command and one `--set full` capture per dominant kernel.  Everything lands in gpurun_out/;
# `python profiles/make_capture.py` (run where ncu can read the reports) turns it into profiles/*_r2.*
#   usage: gpurun --timeout 2400 -- tools/capture_profiles.sh
set -u
O=gpurun_out
NCU="ncu --clock-control none"
python bench.py --steps 20 --warmup 5 > $O/bench_r2_1gpu.json 2> $O/bench_r2_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r2_reference_arm.json 2> $O/bench_r2_reference_arm.err
# launch list of the bench command (short: ncu serialises and replays)
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/launches_bench_r2.csv \
    python bench.py --steps 2 --warmup 1 --min-seconds 0 --no-cpu --no-big > $O/launches_bench_r2.log 2>&1
# dominant kernel of the headline: one launch of 2 steps = 8192 game-plies x 100 simulations
$NCU --set full --import-source on -k regex:selfplay_k1g -s 3 -c 1 -f -o $O/prof_k1g_r2 \
    python bench.py --steps 2 --warmup 3 --steps-per-launch 2 --min-seconds 0 --no-cpu --no-big --no-nn > $O/prof_k1g_r2.log 2>&1
# the many-games regime (16384 games, automatic lanes per game)
$NCU --set full --import-source on -k regex:selfplay_k1g -s 1 -c 1 -f -o $O/prof_k1g_16k_r2 \
    python tools/probe_selfplay.py 16384 0 2 > $O/prof_k1g_16k_r2.log 2>&1
# tower layers: 128 filters (plain + residual layer) and 256 filters
$NCU --set full --import-source on -k regex:conv3x3_tc2_kernel -s 34 -c 2 -f -o $O/prof_conv128_r2 \
    python tools/probe_net.py 5 128 4096 predict > $O/prof_conv128_r2.log 2>&1
$NCU --set full --import-source on -k regex:conv3x3_tc2s -s 122 -c 2 -f -o $O/prof_conv256_r2 \
    python tools/probe_net.py 20 256 4096 predict > $O/prof_conv256_r2.log 2>&1
# the small kernels of an NN wave (tree step, input planes, first layer, heads)
$NCU --set full --import-source on -k regex:"heads_kernel|nn_step_kernel|planes_tiles|conv3x3_tc2_kernel<128, 64" --launch-skip 300 -c 4 -f -o $O/prof_nnaux_r2 \
    python tools/probe_nn_wave.py > $O/prof_nnaux_r2.log 2>&1
$NCU --metrics gpu__time_duration.sum --launch-skip 1700 -c 28 --csv --log-file $O/launches_nnwave_r2.csv \
    python tools/probe_nn_wave.py > /dev/null 2>&1
# board kernels
$NCU --set full --import-source on -k regex:"legal_masks|apply_moves" -c 2 -f -o $O/prof_board_r2 \
    python tools/probe_board.py 4194304 > $O/prof_board_r2.log 2>&1
tail -c 600 $O/bench_r2_1gpu.err
