#!/usr/bin/env python3
"""bench.py -- headline benchmark of the self-play hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--wave 1]

Workload (config.workload): BASELINE.json configs[1] -- batched pure MCTS (uniform prior, one
uniform-random rollout per leaf for the value), 100 simulations per move, 4096 concurrent games
per GPU, reference rules.  A *step* is one self-play ply of all games on the GPU: a full
100-simulation search of every game (one fused kernel) followed by move sampling, sample
recording, make_move and recycling of finished games (play + finalize kernels).

  value   MCTS simulations per second, whole job (all ranks), inputs resident in HBM
  e2e     the same searches through the C ABI with HOST buffers: every step uploads 4096
          positions (pinned host memory) and downloads the [4096,65] visit-count policies
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md "Measurement"

`--impl reference` times the CPU restatement of the reference path (oracle/, kind "port": the
reference itself is pure Python and /root/reference does not exist on the GPU box) with all host
threads on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_GAMES = 4096
N_SIMS = 100
METRIC = "mcts_sims_per_sec"
UNIT = "sims/s"
# algorithmic HBM bytes per simulation of the fused search kernel (DESIGN.md "K2 roofline"):
#   32 B per child row scanned on the way down + 32 B per path node at backup (hot row read +
#   write) + 32 B per child row created (when a traverse first needs it: lazy child rows; 32 B
#   for the legal mask an expansion records) + 24 B leaf record; the rollout itself is register
#   resident (0 B).  Measured per run from the engine's counters, see algorithmic_bytes().


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "of measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "of fallback (B200_PROFILING.md)"


# sources that define the dominant kernel: a committed ncu figure is only valid for the code it was captured on
KERNEL_SOURCES = ["rvs_treeg.cuh", "rvs_tree.cuh", "rvs_board.cuh", "rvs_engine.cu", "rvs_engine.cuh", "rvs_noise.cuh"]
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "traffic_r2.json")


def kernel_source_hash():
    import hashlib
    h = hashlib.sha256()
    for f in KERNEL_SOURCES:
        h.update(open(os.path.join(ROOT, "alphazero-reversi_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def committed_capture():
    """profiles/traffic_r2.json: ncu figures of the dominant kernel (DRAM bytes per step, warp instructions per
    simulation, issue-active) stamped with the commit and the hash of the kernel sources they were captured on"""
    try:
        d = json.load(open(TRAFFIC_FILE))
    except Exception:
        return None, True
    return d, d.get("kernel_source_hash") != kernel_source_hash()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def position_pool(az, n, seed):
    """realistic root positions for the e2e / reference legs: uniform-random games cut at a random
    ply, generated on the GPU (rvs_random_playouts stops are final; use rvs_apply_moves loop)"""
    import numpy as np
    rng = np.random.default_rng(seed)
    bl = np.full(n, 0x0000000810000000, dtype=np.uint64)
    wh = np.full(n, 0x0000001008000000, dtype=np.uint64)
    sd = np.ones(n, dtype=np.uint8)
    fl = np.zeros(n, dtype=np.uint8)
    cut = rng.integers(0, 56, n)
    for ply in range(56):
        lm = az.board_ops.legal_masks(bl, wh, sd)
        mv = np.full(n, 255, dtype=np.uint8)
        live = (cut > ply) & (lm != 0) & ((fl & 1) == 0)
        idx = np.nonzero(live)[0]
        r = rng.integers(0, 64, len(idx))
        for j, i in enumerate(idx):  # r-th set bit (mod popcount)
            m = int(lm[i])
            bits = [b for b in range(64) if (m >> b) & 1]
            mv[i] = bits[r[j] % len(bits)]
        az.board_ops.apply_moves(bl, wh, sd, fl, mv, want_legal=False)
    return bl, wh, sd


def generation_leg(az, dist, dev, rank, world, local, games, barrier):
    """BASELINE configs[4] (config 5) as ONE timed generation per rank: NCCL broadcast of the trainer rank's weights
    -> `games` games of ResNet 5x128 self-play (100 sims/move, one slot per game) played to completion -> replay
    samples gathered on rank 0.  The drain is asynchronous (rvs_engine_drain_packed_async) and the gather runs on a
    SIDE stream while the next generation's first ply is already searching on the main stream."""
    import torch
    from alphazero_reversi_b200 import dist as azd
    L = az._lib
    torch.manual_seed(42)
    rn = az.RvsNetwork.from_module(az.AlphaZeroNetwork(8, 5, 128).eval())
    eng = az.Engine(games, N_SIMS, 1, evaluator=az.EVAL_NN, c_puct=1.0, seed=azd.rank_seed(7, rank), device=local,
                    net_blocks=5, net_filters=128, sample_capacity=64 * games)
    eng.set_option(L.OPT_GAME_LIMIT, games)
    main, side = torch.cuda.current_stream(), torch.cuda.Stream(device=dev)
    E = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    b0, b1, g1, c0, c1, drained = E(), E(), E(), E(), E(), E()
    cnt = torch.zeros(1, dtype=torch.int64).pin_memory()
    barrier()
    t0 = time.perf_counter()
    b0.record()
    flat = rn.flat.to(dev) if rank == 0 else torch.empty_like(rn.flat, device=dev)
    azd.broadcast_weights(flat, src=0)
    b1.record()
    eng.load_weights(flat)
    rounds = 0
    s0 = eng.stats()
    while True:
        eng.selfplay(N_SIMS, plies=games, temperature=1.0, recycle=True)  # one lockstep ply of every live game
        rounds += 1
        if rounds >= 58:
            st = eng.stats()
            if st["games_finished"] >= games or rounds > 70:
                break
    g1.record()
    pk, cnt = eng.drain_packed_async(64 * games, dev, count_out=cnt)
    drained.record()
    # the NEXT generation starts searching at once (main stream); its first ply overlaps the gather below
    eng.reset()
    eng.selfplay(N_SIMS, plies=games, temperature=1.0, recycle=True)
    drained.synchronize()  # waits for this generation's drain only, not for the work enqueued after it
    k = int(cnt[0])
    info = {}
    with torch.cuda.stream(side):
        side.wait_event(drained)
        c0.record(side)
        mine = az.PackedSamples(pk.black[:k], pk.white[:k], pk.side[:k], pk.z[:k], pk.pi[:k])
        res = azd.gather_packed(mine, dst=0, info=info)
        c1.record(side)
    side.synchronize()
    wall = time.perf_counter() - t0
    torch.cuda.synchronize()
    s1 = eng.stats()
    if s1["overflow"] or s1["samples_dropped"] or s1["stalled"]:
        raise SystemExit(f"generation leg: engine error counters non-zero: {s1}")
    # the same gather once more with the GPU otherwise idle (untimed for `wall`): what NVLink does for this message when
    # the send / receive kernels do not have to wait for SMs behind the next generation's persistent network kernel
    alone_ms = 0.0
    if dist is not None:
        barrier()
        a0, a1 = E(), E()
        a0.record()
        azd.gather_packed(mine, dst=0)
        a1.record()
        torch.cuda.synchronize()
        alone_ms = a0.elapsed_time(a1)
    eng.close()
    vals = torch.tensor([wall * 1e3, b0.elapsed_time(b1), b1.elapsed_time(g1), c0.elapsed_time(c1), alone_ms], dtype=torch.float64, device=dev)
    sums = torch.tensor([float(k), float(st["sims"] - s0["sims"]), float(st["games_finished"])], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    wall_ms, bc_ms, sp_ms, ga_ms, alone_ms = vals.tolist()
    samples, sims, finished = sums.tolist()
    if rank != 0:
        return None
    got = len(res) if res is not None else 0
    assert got == int(samples), (got, samples)
    nbytes = info.get("bytes_received", 0)
    return {"config": f"configs[4]: {world} x {games} games of ResNet 5x128 self-play to completion, 100 sims/move, wave 1; "
                      "NCCL weight broadcast before, packed-sample gather (280 B/sample) to rank 0 after",
            "games": int(finished), "samples_gathered_rank0": got, "rounds": rounds, "wall_s": wall_ms * 1e-3,
            "selfplay_s": sp_ms * 1e-3, "sims_per_sec": sims / (wall_ms * 1e-3),
            "broadcast_ms": bc_ms, "broadcast_bytes": int(flat.numel() * 4),
            "gather_ms": ga_ms, "gather_bytes": int(nbytes), "gather_gbs": (nbytes / (ga_ms * 1e-3) / 1e9) if ga_ms > 0 else None,
            "gather_alone_ms": alone_ms, "gather_alone_gbs": (nbytes / (alone_ms * 1e-3) / 1e9) if alone_ms > 0 else None,
            "collective_share": (bc_ms + ga_ms) / wall_ms,
            "collective_exposed_ms": max(0.0, wall_ms - sp_ms), "collective_exposed_share": max(0.0, wall_ms - sp_ms) / wall_ms,
            "overlap": "gather on a side stream, concurrent with the first ply of the next generation on the main stream; "
                       "times are device events (max over ranks), wall is the host clock from before the broadcast to the end of the gather; "
                       "gather_ms is the overlapped gather on the receiving rank: it starts when rank 0 has finished ITS games and ends when "
                       "the slowest rank has sent its samples, so it mostly measures the skew between the ranks' self-play times "
                       "(tested: an NCCL high-priority stream does not shorten it); gather_alone_* is the same message repeated on "
                       "idle GPUs; collective_exposed_* = wall minus the slowest rank's self-play (the weight broadcast plus what "
                       "remains of drain + gather after the last rank has finished playing)"}


def run_ours(args):
    import numpy as np
    import torch
    import alphazero_reversi_b200 as az

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    lib = az._lib.lib()
    stream = torch.cuda.current_stream().cuda_stream
    wave = args.wave
    # sample ring: the timed region may be repeated up to the duration floor (about 1 ms per step at the fastest)
    ring_steps = args.steps + int(args.min_seconds * 1000) + args.steps_per_launch + 100
    eng = az.Engine(N_GAMES, N_SIMS, max(wave, 64), evaluator=az.EVAL_ROLLOUT, c_puct=1.0, seed=1000 + rank, device=local,
                    sample_capacity=ring_steps * N_GAMES)

    persistent = wave == 1 and not args.lockstep
    ppl = max(1, args.steps_per_launch)

    def step(e):
        if persistent:  # one work-conserving launch = one step = N_GAMES game-plies
            e.selfplay(N_SIMS, plies=N_GAMES, temperature=1.0, recycle=True, stream=stream)
        else:
            e.search(N_SIMS, wave, stream=stream)
            e.play(1.0, recycle=True, stream=stream)

    # steady state of a self-play farm: games are spread uniformly over all phases.  Start every
    # slot from a random-playout position cut at a random ply (the slots then recycle naturally);
    # with all games at the same ply the rollout length -- and the step time -- would be a
    # function of the ply instead of the workload.
    pb0, pw0, ps0 = position_pool(az, N_GAMES, 99 + rank)
    eng.set_positions(pb0, pw0, ps0, stream=stream)
    # untimed: `presteps` plies that spread the games over all phases, then W >= 3 warm-up steps
    warmup_run = max(args.warmup, 3)
    for _ in range(args.presteps + warmup_run):
        step(eng)
    torch.cuda.synchronize()
    # duration floor: a K-step region is ~1.4 ms x K (29 ms at the driver's K = 20, shorter than one clock sample), so
    # the K-step region is repeated back to back `reps` times inside ONE event pair until >= --min-seconds are timed
    # (steps stays K; ms_per_step = total / (K x reps))
    reps = 1
    if persistent and args.min_seconds > 0:
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        eng.selfplay(N_SIMS, plies=N_GAMES * min(ppl, args.steps), temperature=1.0, recycle=True, stream=stream)
        c1.record()
        torch.cuda.synchronize()
        est = c0.elapsed_time(c1) * 1e-3 * args.steps / min(ppl, args.steps)
        reps = max(1, int(args.min_seconds / max(est, 1e-6) + 0.999))
    if dist is not None:  # every rank times the same number of steps
        rt = torch.tensor([reps], dtype=torch.int64, device=dev)
        dist.all_reduce(rt, op=dist.ReduceOp.MAX)
        reps = int(rt.item())

    eng.drain_packed(device=dev)  # samples of the untimed steps are not part of the measured generation
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    s0 = eng.stats()
    l0 = lib.rvs_launch_count()
    ks = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps * reps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    if persistent:
        # K x reps steps = K x reps x N_GAMES game-plies, issued as persistent launches of `ppl` steps each: the launch size
        # is the engine's business (a self-play farm calls rvs_engine_selfplay with a large ply budget), the timed region is
        # exactly K x reps steps.  A launch ends with a tail of about one early-game ply, so short launches cost throughput:
        # 20 / 50 / 100 / 400 steps per launch give 2.75 / 2.84 / 2.88 / 2.91e8 sims/s.
        used, launch_steps = [], []
        done, total = 0, args.steps * reps
        while done < total:
            n = min(ppl, total - done)
            ev = ks[len(used)]
            ev[0].record()
            eng.selfplay(N_SIMS, plies=N_GAMES * n, temperature=1.0, recycle=True, stream=stream)
            ev[1].record()
            used.append(ev)
            launch_steps.append(n)
            done += n
        ks = used
    else:
        for i in range(args.steps * reps):
            ks[i][0].record()
            eng.search(N_SIMS, wave, stream=stream)
            ks[i][1].record()
            eng.play(1.0, recycle=True, stream=stream)
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    s1 = eng.stats()
    launches = lib.rvs_launch_count() - l0
    kernel_ms = sum(a.elapsed_time(b) for a, b in ks) / len(ks)  # average launch duration of the dominant kernel
    total_steps = args.steps * reps
    d = {k: s1[k] - s0[k] for k in s1}
    d["n_search_launches"] = len(ks)
    if s1["overflow"] or s1["samples_dropped"] or s1["stalled"]:
        raise SystemExit(f"engine error counters non-zero: {s1}")

    # ---- e2e: C ABI with host buffers (pinned), H2D + D2H inside the timed region ----------
    pool_steps = 4
    pb, pw, ps = position_pool(az, N_GAMES * pool_steps, 7 + rank)
    hb = torch.empty(N_GAMES * pool_steps, dtype=torch.int64).pin_memory()
    hw = torch.empty_like(hb).pin_memory()
    hs = torch.empty(N_GAMES * pool_steps, dtype=torch.uint8).pin_memory()
    hb.numpy()[:] = pb.view(np.int64)
    hw.numpy()[:] = pw.view(np.int64)
    hs.numpy()[:] = ps
    # The lockstep search call ends with a tail (the longest, early-game searches), so the headline e2e leg
    # pipelines `depth` engine handles round-robin on their own CUDA streams through the asynchronous
    # host mode of the C ABI (RVS_MEM_HOST_ASYNC): while one batch drains its tail the next fills the SMs.
    # The same loop at depth 1 (one handle, one stream, 4096 games in flight = the concurrency of `value`)
    # is reported as e2e_depth1.
    MH = az._lib.MEM_HOST_ASYNC

    def measure_e2e(depth, lanes):
        streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]
        hvs = [torch.empty((N_GAMES, 65), dtype=torch.int32).pin_memory() for _ in range(depth)]
        engs = [az.Engine(N_GAMES, N_SIMS, max(wave, 64), evaluator=az.EVAL_ROLLOUT, c_puct=1.0, seed=2000 + rank, device=local)
                for _ in range(depth)]
        for e in engs:  # depth x 4096 games are in flight: the many-games setting of the wave-1 kernels
            e.set_lanes_per_game(lanes)

        def e2e_step(i):
            o = (i % pool_steps) * N_GAMES
            j = i % depth
            st, e = streams[j], engs[j]
            st.synchronize()  # results of this handle's previous step are complete and consumed
            az._lib.check(lib.rvs_engine_set_positions(e._h, hb[o:].data_ptr(), hw[o:].data_ptr(), hs[o:].data_ptr(),
                                                       N_GAMES, MH, st.cuda_stream))
            az._lib.check(lib.rvs_engine_search(e._h, N_SIMS, wave, st.cuda_stream))
            az._lib.check(lib.rvs_engine_root_visits(e._h, hvs[j].data_ptr(), N_GAMES, MH, st.cuda_stream))

        for i in range(3 * depth):
            e2e_step(i)
        torch.cuda.synchronize()
        steps_ = max(4 * depth, args.steps // 2, 24)
        barrier()
        t0 = time.perf_counter()
        for i in range(steps_):
            e2e_step(i)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3  # host clock around H2D + search + D2H of every step, all streams drained
        barrier()
        for hv in hvs:
            assert int(hv.numpy().sum()) > 0
        for e in engs:
            e.close()
        return steps_, wall

    depth = max(1, args.e2e_depth)
    e2e_steps, e2e_ms = measure_e2e(depth, args.e2e_lanes if args.e2e_lanes else (2 if depth >= 12 else (4 if depth >= 3 else 0)))
    e2e1_steps, e2e1_ms = measure_e2e(1, 0) if depth != 1 else (e2e_steps, e2e_ms)

    # ---- config-1 side metric: register-resident uniform-random playouts ------------------
    n_po = 1 << 20
    az.board_ops.random_playouts(n_po, 1, outputs=False, stream=stream)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    *_, po_steps = az.board_ops.random_playouts(n_po, 2 + rank, outputs=False, stream=stream)
    g1.record()
    torch.cuda.synchronize()
    po_ms = g0.elapsed_time(g1)

    # ---- config-0 side metric: perft(8) from the start position under the reference's rules (391 210 leaves)
    az.board_ops.perft(6)
    tp0 = time.perf_counter()
    perft_nodes = az.board_ops.perft(8)
    perft_ms = (time.perf_counter() - tp0) * 1e3

    # ---- side metric: the same workload with 16384 concurrent games (4 games' worth of latency hiding per
    # scheduler more than the headline): shows how far the 4096-game figure is from the issue-bound rate
    big = None
    if not args.no_big:
        GB = 16384
        engb = az.Engine(GB, N_SIMS, 1, evaluator=az.EVAL_ROLLOUT, c_puct=1.0, seed=4000 + rank, device=local,
                         sample_capacity=80 * GB)
        engb.set_positions(np.tile(pb0, GB // N_GAMES), np.tile(pw0, GB // N_GAMES), np.tile(ps0, GB // N_GAMES), stream=stream)
        engb.selfplay(N_SIMS, plies=GB * 2, temperature=1.0, recycle=True, stream=stream)
        torch.cuda.synchronize()
        b0s = engb.stats()
        h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h0.record()
        engb.selfplay(N_SIMS, plies=GB * 6, temperature=1.0, recycle=True, stream=stream)
        h1.record()
        torch.cuda.synchronize()
        b1s = engb.stats()
        bms = h0.elapsed_time(h1)
        big = {"games": GB, "sims_per_sec": (b1s["sims"] - b0s["sims"]) / (bms * 1e-3),
               "board_steps_per_sec": (b1s["board_steps"] - b0s["board_steps"]) / (bms * 1e-3), "ms": bms}
        engb.close()

    # ---- config-3 side metric: ResNet 5x128 (default_config.json) NN-evaluated self-play ------
    nn = None
    if not args.no_nn:
        torch.manual_seed(42)
        net = az.AlphaZeroNetwork(8, 5, 128).eval()  # random-init weights of the reference architecture
        rn = az.RvsNetwork.from_module(net)
        if dist is not None:  # config 5: the trainer rank broadcasts the packed weights over NCCL
            from alphazero_reversi_b200 import dist as azd
            wdev = rn.flat.to(dev) if rank == 0 else torch.zeros_like(rn.flat, device=dev)
            azd.broadcast_weights(wdev, src=0)
            rn = az.RvsNetwork(wdev, rn.net_blocks, rn.net_filters)
        eng3 = az.Engine(N_GAMES, N_SIMS, 1, evaluator=az.EVAL_NN, c_puct=1.0, seed=3000 + rank, device=local,
                         net_blocks=5, net_filters=128)
        rn.attach(eng3)
        eng3.set_positions(pb0, pw0, ps0, stream=stream)
        for _ in range(2):
            eng3.search(N_SIMS, 1, stream=stream); eng3.play(1.0, recycle=True, stream=stream)
        torch.cuda.synchronize()
        t0s = eng3.stats()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nn_steps = 3
        n0.record()
        for _ in range(nn_steps):
            eng3.search(N_SIMS, 1, stream=stream); eng3.play(1.0, recycle=True, stream=stream)
        n1.record()
        torch.cuda.synchronize()
        nn_ms = n0.elapsed_time(n1)
        t1s = eng3.stats()
        flops_per_eval = 2 * (64 * 27 * 128 + 2 * 5 * 64 * 9 * 128 * 128 + 64 * 128 * 2 + 128 * 65 + 64 * 128 + 64 * 256 + 256)
        batch_evals = t1s["nn_evals"] - t0s["nn_evals"]  # boards run through the network (compacted leaf batches)
        nn = {"sims_per_sec": (t1s["sims"] - t0s["sims"]) / (nn_ms * 1e-3),
              "consumed_evals_per_sec": (t1s["evals"] - t0s["evals"]) / (nn_ms * 1e-3),
              "network_evals_per_sec": batch_evals / (nn_ms * 1e-3), "ms_per_step": nn_ms / nn_steps,
              "tflops": batch_evals * flops_per_eval / (nn_ms * 1e-3) / 1e12, "flops_per_eval": flops_per_eval}
        # the reference's own default: MCTS(batch_size=64) -- most simulations of a wave share one leaf
        # (SURVEY.md 0.3), which the engine evaluates once per wave
        eng3.close()
        eng3 = az.Engine(N_GAMES, N_SIMS, 64, evaluator=az.EVAL_NN, c_puct=1.0, seed=3500 + rank, device=local,
                         net_blocks=5, net_filters=128)
        rn.attach(eng3)
        eng3.set_positions(pb0, pw0, ps0, stream=stream)
        eng3.search(N_SIMS, 64, stream=stream); eng3.play(1.0, recycle=True, stream=stream)
        torch.cuda.synchronize()
        u0 = eng3.stats()
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0.record()
        for _ in range(5):
            eng3.search(N_SIMS, 64, stream=stream); eng3.play(1.0, recycle=True, stream=stream)
        m1.record()
        torch.cuda.synchronize()
        u1 = eng3.stats()
        wms = m0.elapsed_time(m1)
        # FAST mode (engine feature, virtual-loss PUCT): the waves spread over distinct leaves, so wave > 1 buys search
        fastm = {}
        for K in (8, 16, 64):
            engf = az.Engine(N_GAMES, N_SIMS, K, evaluator=az.EVAL_NN, c_puct=1.0, seed=3600 + rank, device=local,
                             net_blocks=5, net_filters=128)
            engf.set_search_mode(az.MODE_FAST)
            rn.attach(engf)
            engf.set_positions(pb0, pw0, ps0, stream=stream)
            engf.search(N_SIMS, K, stream=stream); engf.play(1.0, recycle=True, stream=stream)
            torch.cuda.synchronize()
            f0 = engf.stats()
            y0, y1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            y0.record()
            for _ in range(2):
                engf.search(N_SIMS, K, stream=stream); engf.play(1.0, recycle=True, stream=stream)
            y1.record()
            torch.cuda.synchronize()
            f1 = engf.stats()
            fms = y0.elapsed_time(y1)
            fastm[f"wave{K}"] = {"sims_per_sec": (f1["sims"] - f0["sims"]) / (fms * 1e-3),
                                 "network_evals_per_sec": (f1["nn_evals"] - f0["nn_evals"]) / (fms * 1e-3),
                                 "unique_evals_per_sim": (f1["nn_evals"] - f0["nn_evals"]) / max(1, f1["sims"] - f0["sims"]),
                                 "tflops": (f1["nn_evals"] - f0["nn_evals"]) * flops_per_eval / (fms * 1e-3) / 1e12}
            engf.close()
        nn["fast_mode"] = dict(fastm, config="RVS_MODE_FAST (virtual-loss PUCT leaf batching, DESIGN.md): same network and games; "
                                             "reference-compatible waves evaluate ~1 unique leaf per wave (wave64 block)")
        nn["wave64"] = {"sims_per_sec": (u1["sims"] - u0["sims"]) / (wms * 1e-3),
                        "network_evals_per_sec": (u1["nn_evals"] - u0["nn_evals"]) / (wms * 1e-3),
                        "consumed_evals_per_sec": (u1["evals"] - u0["evals"]) / (wms * 1e-3), "ms_per_step": wms / 5,
                        "unique_evals_per_sim": (u1["nn_evals"] - u0["nn_evals"]) / max(1, u1["sims"] - u0["sims"]),
                        "config": "same network and games with the reference's default MCTS(batch_size=64) wave semantics"}
        eng3.close()

    # ---- config-4 side metric: ResNet 20x256, 16384 concurrent games, Dirichlet root noise; a bounded slice
    # (32 of the 800 simulations of one ply) -- a full ply is 13.1 M network evaluations (~27 s)
    nn4 = None
    if not args.no_nn and not args.no_big:
        torch.manual_seed(42)
        net4 = az.AlphaZeroNetwork(8, 20, 256).eval()
        rn4 = az.RvsNetwork.from_module(net4)
        G4, S4 = 16384, 32
        # node pools sized for the full 800 simulations per move (14 GB of tree rows), the timed slice runs 32 of them
        eng4 = az.Engine(G4, 800, 1, evaluator=az.EVAL_NN, c_puct=1.0, seed=5000 + rank, device=local, net_blocks=20, net_filters=256)
        rn4.attach(eng4)
        eng4.set_root_noise(0.03, 0.25)  # src/config.py:25-26
        eng4.set_positions(np.tile(pb0, G4 // N_GAMES), np.tile(pw0, G4 // N_GAMES), np.tile(ps0, G4 // N_GAMES), stream=stream)
        eng4.search(4, 1, stream=stream)
        torch.cuda.synchronize()
        q0s = eng4.stats()
        q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        q0.record()
        eng4.search(S4, 1, stream=stream)
        q1.record()
        torch.cuda.synchronize()
        q1s = eng4.stats()
        qms = q0.elapsed_time(q1)
        f4 = 2 * (64 * 27 * 256 + 2 * 20 * 64 * 9 * 256 * 256 + 64 * 256 * 2 + 128 * 65 + 64 * 256 + 64 * 256 + 256)
        e4 = q1s["nn_evals"] - q0s["nn_evals"]
        nn4 = {"sims_per_sec": (q1s["sims"] - q0s["sims"]) / (qms * 1e-3), "network_evals_per_sec": e4 / (qms * 1e-3),
               "tflops": e4 * f4 / (qms * 1e-3) / 1e12, "flops_per_eval": f4, "ms": qms,
               "config": "configs[3]: ResNet 20x256, 16384 concurrent games, Dirichlet(0.03, 0.25) root noise, bf16, wave 1; "
                         f"timed slice = {S4} of the 800 simulations of one ply"}
        eng4.close()
        del net4, rn4

    # ---- config 5: replay samples of the timed self-play gathered to rank 0 (outside the timing) --
    gathered = None
    pk_ = eng.drain_packed(device=dev)  # packed rows (280 B) travel over NVLink; planes are re-derived on rank 0
    if dist is not None:
        from alphazero_reversi_b200 import dist as azd
        cap = min(len(pk_), 65536)
        res = azd.gather_packed(az.PackedSamples(pk_.black[:cap], pk_.white[:cap], pk_.side[:cap], pk_.z[:cap], pk_.pi[:cap]), dst=0)
        gathered = None if res is None else int(res.states().shape[0])
    else:
        gathered = len(pk_)

    # ---- config 5 as a timed leg: broadcast -> N x 8192 NN self-play games to completion -> gather ----------
    generation = None
    if (dist is not None or args.generation) and not args.no_nn and not args.no_generation:
        generation = generation_leg(az, dist, dev, rank, world, local, args.gen_games, barrier)

    # ---- reductions over ranks --------------------------------------------------------------
    rank_sims_per_sec = d["sims"] / (ms * 1e-3)  # this rank's GPU (roofline of the dominant kernel)
    vals = torch.tensor([ms, e2e_ms, kernel_ms, e2e1_ms], dtype=torch.float64, device=dev)
    sums = torch.tensor([d["sims"], d["board_steps"], d["evals"], float(launches), e2e_steps * N_GAMES * N_SIMS,
                         po_steps / (po_ms * 1e-3), e2e1_steps * N_GAMES * N_SIMS,
                         0.0 if big is None else big["sims_per_sec"]], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    ms, e2e_ms, kernel_ms, e2e1_ms = vals.tolist()
    sims, bsteps, evals, launches_all, e2e_sims, po_rate, e2e1_sims, big_all = sums.tolist()

    if rank == 0:
        hbm, which = peaks()
        # algorithmic bytes of the fused search kernel per launch (this rank), from its counters
        tree_bytes = algorithmic_bytes(d)
        achieved = tree_bytes / (kernel_ms * 1e-3) / 1e9
        cap, stale = committed_capture()
        kcap = (cap or {}).get("selfplay_k1g_kernel", {}) if persistent else {}
        traffic = args.traffic  # DRAM bytes per launch of the dominant kernel, from the committed ncu capture
        if traffic is None and kcap.get("dram_bytes_per_step") is not None:
            traffic = kcap["dram_bytes_per_step"] * (total_steps / len(ks) if persistent else 1)
        # The binding roof of the dominant kernel is the ISSUE rate, not HBM (rollouts are register resident; DRAM
        # traffic is a few % of peak): achieved = warp instructions per simulation (ncu smsp__inst_executed.sum /
        # simulations of the committed capture, stamped with the hash of the kernel sources) x measured sims/s;
        # peak = 148 SMs x 4 schedulers x 1 warp instruction per cycle x the SM clock sampled under this load
        sm_mhz = (clocks or {}).get("sm_mhz") or (clocks or {}).get("sm_max_mhz") or 1965.0
        issue_peak = 148 * 4 * sm_mhz * 1e6
        wips = kcap.get("warp_inst_per_sim")
        issue_ach = wips * rank_sims_per_sec if wips else None
        kname = ("selfplay_k1g_kernel<REF,ROLLOUT,8>" if persistent else
                 ("search_k1g_kernel<REF,ROLLOUT,8>" if wave == 1 else "search_fused_kernel<REF,ROLLOUT>"))
        out = {
            "metric": METRIC, "value": sims / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / total_steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64+f32", "data": "synthetic",
            "timing": {"timed_repeats": reps, "timed_steps_total": total_steps, "timed_ms": ms, "warmup_steps_run": warmup_run,
                       "presteps": args.presteps,
                       "note": "the K-step region is repeated back to back inside one CUDA-event pair until >= min_seconds are "
                               "timed; presteps spread the games over all phases before the warm-up", "min_seconds": args.min_seconds},
            "config": {"workload": "configs[1]: pure MCTS, uniform prior + uniform-random rollout value, 100 sims/move, "
                                   "4096 concurrent games per GPU, REF rules, self-play with recycling",
                       "games_per_gpu": N_GAMES, "sims_per_move": N_SIMS, "wave": wave, "c_puct": 1.0, "temperature": 1.0,
                       "schedule": ("persistent work-conserving self-play launches" if persistent else "lockstep search+play per ply"),
                       "parallelism": f"games sharded x{world}, no data-path collective",
                       "cache": "node pools 669 MB per GPU (> 126 MB L2); every step rebuilds all 4096 trees from their roots, nothing a step "
                                "reads was produced by an earlier step except the 18-byte game positions (the rows a step touches, ~25 MB of "
                                "DRAM traffic under ncu, stay L2 resident by design: lazy child rows)"},
            "board_steps_per_sec": bsteps / (ms * 1e-3),
            "unique_evals_per_sec": evals / (ms * 1e-3),
            "playout_board_steps_per_sec": po_rate,
            "value_16384": (big_all if big is not None else None),
            "e2e": {"value": e2e_sims / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": N_GAMES * 17,
                    "d2h_bytes_per_step": N_GAMES * 65 * 4, "steps": e2e_steps, "pipeline_depth": depth, "games_in_flight": depth * N_GAMES,
                    "api": "rvs_engine_set_positions(host) -> rvs_engine_search -> rvs_engine_root_visits(host), "
                           "RVS_MEM_HOST_ASYNC on one stream per engine handle; timed with the host clock"},
            "e2e_depth1": {"value": e2e1_sims / (e2e1_ms * 1e-3), "unit": UNIT, "steps": e2e1_steps, "pipeline_depth": 1,
                           "games_in_flight": N_GAMES,
                           "note": "the same host-buffer loop with ONE engine handle: the concurrency `value` is measured at; each call "
                                   "ends in the tail of its longest searches, which the pipelined e2e hides"},
            "gpu_launches": int(launches_all),
            "samples_gathered_rank0": gathered,
            "roofline": {"bound": "issue", "achieved": (issue_ach / 1e9 if issue_ach else None), "peak": issue_peak / 1e9,
                         "unit": "Gwarp-inst/s", "frac": (issue_ach / issue_peak if issue_ach else None),
                         "traffic": traffic, "traffic_stale": bool(stale),
                         "kernel": kname, "warp_inst_per_sim": wips, "sm_mhz_used": sm_mhz,
                         "peak_source": "148 SMs x 4 schedulers x SM clock sampled by nvidia-smi during the timed region",
                         "capture": ({"file": os.path.relpath(TRAFFIC_FILE, ROOT), "commit": cap.get("commit"),
                                      "kernel_source_hash": cap.get("kernel_source_hash"), "current_hash": kernel_source_hash(),
                                      "issue_active_pct_under_ncu": kcap.get("issue_active_pct"),
                                      "alu_pipe_active_pct_under_ncu": kcap.get("alu_pipe_active_pct")} if cap else None),
                         "steps_per_launch": (total_steps / len(ks)) if persistent else 1,
                         "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms * len(ks) / ms,
                         "hbm": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                                 "peak_source": which, "algorithmic_bytes_per_sim": d["tree_bytes"] / max(1, d["sims"])},
                         "note": "rollouts are register resident (0 B), so HBM is not the roof (secondary `hbm` block); the kernel is "
                                 "bound by instruction issue: its 64-bit shift/logic floods run on the half-rate ALU pipe, which a "
                                 "single warp already saturates during a flood, and at 4096 games (1.7 warps per scheduler) the "
                                 "exchange / vote latencies between floods stay exposed (see DESIGN.md K2)"},
            "clocks": clocks,
        }
        if generation is not None:
            out["generation"] = generation
        out["perft8"] = {"leaves": int(perft_nodes), "ms": perft_ms, "expected": 391210}
        if big is not None:
            out["games_16384"] = big
        if nn is not None:
            try:
                pk = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"])
                pk_src = "of measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernels timed inside a long step)"
            except Exception:
                pk = 1590.0  # B200_PROFILING.md fallback (burst figure; the sustained one is about 15 % lower)
                pk_src = "of fallback (B200_PROFILING.md)"
            nn["roofline"] = {"bound": "tensor", "achieved": nn["tflops"], "peak": pk, "unit": "TFLOP/s",
                              "frac": nn["tflops"] / pk, "peak_source": pk_src, "note": "whole search step incl. tree kernels; rank 0"}
            nn["config"] = "configs[2]: ResNet 5x128 self-play, 100 sims/move, 4096 games, bf16, random-init weights, wave 1"
            out["nn"] = nn
            if nn4 is not None:
                nn4["roofline"] = {"bound": "tensor", "achieved": nn4["tflops"], "peak": pk, "unit": "TFLOP/s", "frac": nn4["tflops"] / pk,
                                   "peak_source": pk_src, "note": "whole search slice incl. tree kernels; rank 0"}
                out["nn_20x256"] = nn4
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(wave, threads=os.cpu_count() or 1, budget_s=12.0)
        _emit(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


def algorithmic_bytes(d):
    """algorithmic HBM bytes of ONE search launch (this rank): the engine counts 32 B for every
    node row the search must touch (children scanned, path rows read+written at backup, rows
    created) -- DESIGN.md 'K2 roofline'; rollouts are register resident and count 0 B."""
    return d["tree_bytes"] / max(1, d["n_search_launches"])


def cpu_baseline(wave, threads, budget_s):
    """oracle port timed on the host cores: searches of the same workload on a bounded sample"""
    import numpy as np
    import orc
    n = 256
    rng = np.random.default_rng(3)
    bl, wh, wi, pl = orc.random_playouts(1, 1)  # warm the library
    pos_b = np.full(n, orc.START[0], dtype=np.uint64)
    pos_w = np.full(n, orc.START[1], dtype=np.uint64)
    pos_s = np.ones(n, dtype=np.uint8)
    # mid-game roots: play k random plies with the oracle
    import ctypes as C
    L = orc.lib()
    for i in range(n):
        b = orc.make_board(*orc.START)
        st = L.orc_stream_seed(11, i, 0)
        for _ in range(int(rng.integers(0, 56))):
            if b.over:
                break
            lm = L.orc_board_legal(C.byref(b), 0)
            bits = [q for q in range(64) if (lm >> q) & 1]
            st = orc.mix64(st + 1)
            L.orc_apply(C.byref(b), bits[st % len(bits)], 0)
        pos_b[i], pos_w[i], pos_s[i] = b.black, b.white, b.side
    t0 = time.perf_counter()
    sims = steps = 0
    rounds = 0
    while time.perf_counter() - t0 < budget_s:
        v, ev, st = orc.search_batch(pos_b, pos_w, pos_s, N_SIMS, wave, evaluator=orc.EVAL_ROLLOUT, seed=rounds,
                                     threads=threads)
        sims += n * N_SIMS
        steps += st
        rounds += 1
    dt = time.perf_counter() - t0
    tq = time.perf_counter()
    p8 = orc.perft(8)
    p8_ms = (time.perf_counter() - tq) * 1e3
    port = {"value": sims / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "board_steps_per_sec": steps / dt, "perft8_ms_one_thread": p8_ms, "perft8_leaves": int(p8),
            "sample": f"{rounds} x {n} mid-game roots x {N_SIMS} sims (wave {wave}), C oracle (oracle/rvs_oracle.c), "
                      f"{threads} host threads, {dt:.1f} s"}
    out = dict(port)
    out["port"] = port
    out["python_reference"] = python_reference()
    return out


def python_reference():
    """The reference's OWN Python path.  It is pure Python, /root/reference does not exist on
    the GPU box and reference sources may not be copied into the repo, so it cannot run here: if it IS importable
    (a checkout at $RVS_REFERENCE or /root/reference) the committed script is run live in a CUDA_VISIBLE_DEVICES=""
    subprocess; otherwise the figures measured with that script in the build container are reported, labelled."""
    ref = os.environ.get("RVS_REFERENCE", "/root/reference")
    script = os.path.join(ROOT, "oracle", "time_python_reference.py")
    if os.path.isdir(os.path.join(ref, "src", "mcts")):
        try:
            tmp = os.path.join("/tmp", f"pyref_{os.getpid()}.json")
            env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RVS_REFERENCE=ref)
            subprocess.run([sys.executable, script, "--quick", "--out", tmp], env=env, check=True, capture_output=True, timeout=600)
            live = json.load(open(tmp))
            live["where"] = "live on this host (quick sample)"
            return live
        except Exception as ex:  # noqa: BLE001
            return {"unavailable": f"reference present but the timing script failed: {ex!r}"}
    path = os.path.join(ROOT, "profiles", "python_reference_cpu_r2.json")
    if os.path.exists(path):
        d = json.load(open(path))
        d["where"] = ("NOT this host: the reference is not importable here (no /root/reference on the GPU box); figures measured "
                      "in the build container by oracle/time_python_reference.py and committed under profiles/")
        return d
    return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    import numpy as np
    import orc
    import ctypes as C
    n = N_GAMES  # the full 4096-game step (the C port needs ~0.3 s per step on 16 threads)
    rng = np.random.default_rng(3)
    L = orc.lib()
    pos_b = np.zeros(n, dtype=np.uint64); pos_w = np.zeros(n, dtype=np.uint64); pos_s = np.ones(n, dtype=np.uint8)
    for i in range(n):
        b = orc.make_board(*orc.START)
        st = L.orc_stream_seed(11, i, 0)
        for _ in range(int(rng.integers(0, 56))):
            if b.over:
                break
            lm = L.orc_board_legal(C.byref(b), 0)
            bits = [q for q in range(64) if (lm >> q) & 1]
            st = orc.mix64(st + 1)
            L.orc_apply(C.byref(b), bits[st % len(bits)], 0)
        pos_b[i], pos_w[i], pos_s[i] = b.black, b.white, b.side
    for w in range(max(args.warmup, 1)):
        orc.search_batch(pos_b, pos_w, pos_s, N_SIMS, args.wave, seed=w, threads=threads)
    t0 = time.perf_counter()
    steps_total = 0
    for k in range(args.steps):
        _, _, st = orc.search_batch(pos_b, pos_w, pos_s, N_SIMS, args.wave, seed=100 + k, threads=threads)
        steps_total += st
    dt = time.perf_counter() - t0
    v = args.steps * n * N_SIMS / dt
    sample = (f"{args.steps} steps x {n} mid-game roots x {N_SIMS} sims (wave {args.wave}), C oracle port of the "
              f"reference path, {threads} host threads")
    _emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3 / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64+f32", "data": "synthetic",
        "config": {"workload": "configs[1]: pure MCTS, uniform prior + uniform-random rollout value, 100 sims/move, "
                               f"{n} concurrent games per step (searches of mid-game roots)", "sims_per_move": N_SIMS,
                   "wave": args.wave},
        "board_steps_per_sec": steps_total / dt,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "python_reference": python_reference()},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def _emit(line):
    """the ONE JSON line goes to the process's original stdout"""
    os.write(_REAL_STDOUT, (line + "\n").encode())


_REAL_STDOUT = 1


def main():
    # libraries write to stdout too (NCCL prints its version banner there): keep fd 1 for the JSON line only
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=600)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--presteps", type=int, default=5, help="untimed plies that spread games over all phases")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--wave", type=int, default=1, help="MCTS batch_size per game (reference default 64)")
    ap.add_argument("--steps-per-launch", type=int, default=100,
                    help="persistent self-play: steps (x4096 game-plies) per launch")
    ap.add_argument("--lockstep", action="store_true", help="wave 1 through search+play launches per ply")
    ap.add_argument("--traffic", type=float, default=None, help="ncu dram bytes per launch of the search kernel")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-depth", type=int, default=16, help="engine handles pipelined in the e2e leg")
    ap.add_argument("--e2e-lanes", type=int, default=0, help="lanes per game of the e2e engines (0: 2 from 12 handles, 4 from 3)")
    ap.add_argument("--no-big", action="store_true", help="skip the 16384-game side measurement")
    ap.add_argument("--no-nn", action="store_true", help="skip the config-3 (ResNet) side measurement")
    ap.add_argument("--min-seconds", type=float, default=0.5, help="floor of the timed region (the K steps are repeated)")
    ap.add_argument("--generation", action="store_true", help="run the config-5 generation leg on one GPU too")
    ap.add_argument("--no-generation", action="store_true")
    ap.add_argument("--gen-games", type=int, default=8192, help="games per rank of the generation leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
