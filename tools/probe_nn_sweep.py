"""probe: 5x128 wave-1 NN search step (4096 games, 100 sims) over the pipeline option and the cap of the tensor-core grid
usage: python tools/probe_nn_sweep.py [steps]"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import alphazero_reversi_b200 as az
from bench import position_pool
G, steps = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 4
L = az._lib
torch.manual_seed(42)
rn = az.RvsNetwork.from_module(az.AlphaZeroNetwork(8, 5, 128).eval())
flops = 2 * (64 * 27 * 128 + 2 * 5 * 64 * 9 * 128 * 128 + 64 * 128 * 2 + 128 * 65 + 64 * 128 + 64 * 256 + 256)
pb, pw, ps = position_pool(az, G, 99)
for pipe in (0, 1):
    for ctas in (0, 140, 136, 132, 128, 124, 116):
        eng = az.Engine(G, 100, 1, evaluator=az.EVAL_NN, seed=3000, net_blocks=5, net_filters=128)
        eng.set_option(L.OPT_NET_PIPELINE, pipe); eng.set_option(L.OPT_NET_MAX_CTAS, ctas)
        rn.attach(eng)
        eng.set_positions(pb, pw, ps)
        for _ in range(2): eng.search(100, 1); eng.play(1.0, recycle=True)
        torch.cuda.synchronize()
        s0 = eng.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps): eng.search(100, 1); eng.play(1.0, recycle=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1); s1 = eng.stats()
        print(f"pipeline {pipe} ctas {ctas or 148:3d}: {(s1['sims']-s0['sims'])/ms/1e3:6.3f} M sims/s {(s1['nn_evals']-s0['nn_evals'])*flops/ms/1e9:7.1f} TFLOP/s", flush=True)
        eng.close()
