"""probe: configs[2] NN self-play step (5x128, 4096 games, 100 sims, wave 1) under the engine options
usage: python tools/probe_nn_search.py [games] [steps] -- prints sims/s and TFLOP/s per option set"""
import sys, time, json
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import alphazero_reversi_b200 as az
from bench import position_pool

if "--timeline" in sys.argv:  # debug build with -DRVS_TIMELINE (build/tl/): phase boundaries of three waves on stderr
    sys.argv.remove("--timeline")
    az._lib.LIB_PATH = "build/tl/librvs_b200_tl.so"
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
L = az._lib
torch.manual_seed(42)
rn = az.RvsNetwork.from_module(az.AlphaZeroNetwork(8, 5, 128).eval())
flops = 2 * (64 * 27 * 128 + 2 * 5 * 64 * 9 * 128 * 128 + 64 * 128 * 2 + 128 * 65 + 64 * 128 + 64 * 256 + 256)
pb, pw, ps = position_pool(az, G, 99)
ref = None
for name, opts in (("lockstep (graph off, pipeline off)", {L.OPT_NET_PIPELINE: 0}),
                   ("pipelined halves", {L.OPT_NET_PIPELINE: 1}),
                   ("pipelined halves, tower on 140 CTAs", {L.OPT_NET_PIPELINE: 1, L.OPT_NET_MAX_CTAS: 140}),
                   ("pipelined halves, tower on 132 CTAs", {L.OPT_NET_PIPELINE: 1, L.OPT_NET_MAX_CTAS: 132})):
    eng = az.Engine(G, 100, 1, evaluator=az.EVAL_NN, seed=3000, net_blocks=5, net_filters=128)
    for k, v in opts.items():
        eng.set_option(k, v)
    rn.attach(eng)
    import os
    if os.environ.get('RVS_TOWER') is not None: eng.set_option(L.OPT_NET_TOWER, int(os.environ['RVS_TOWER']))
    eng.set_positions(pb, pw, ps)
    eng.search(100, 1)
    v = eng.root_visits()
    if ref is None:
        ref = v
    same = bool(np.array_equal(v, ref))
    eng.play(1.0, recycle=True)
    eng.search(100, 1); eng.play(1.0, recycle=True)
    torch.cuda.synchronize()
    s0 = eng.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        eng.search(100, 1); eng.play(1.0, recycle=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    s1 = eng.stats()
    ev = s1["nn_evals"] - s0["nn_evals"]
    print(f"{name:45s} {(s1['sims']-s0['sims'])/ms/1e3:7.3f} M sims/s  {ev*flops/ms/1e9:7.1f} TFLOP/s  {ms/steps:7.2f} ms/ply  first-search visits identical to lockstep: {same}")
    eng.close()
