"""probe: a few lockstep waves of the configs[2] NN search (5x128, 4096 games, wave 1) for an ncu launch list
usage: ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1600 -c 48 --csv python tools/probe_nn_wave.py [games]"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import alphazero_reversi_b200 as az
from bench import position_pool
G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
torch.manual_seed(42)
rn = az.RvsNetwork.from_module(az.AlphaZeroNetwork(8, 5, 128).eval())
pb, pw, ps = position_pool(az, G, 99)
eng = az.Engine(G, 100, 1, evaluator=az.EVAL_NN, seed=3000, net_blocks=5, net_filters=128)
rn.attach(eng)
eng.set_positions(pb, pw, ps)
eng.search(100, 1)
eng.play(1.0, recycle=True)
eng.search(100, 1)
torch.cuda.synchronize()
print("done", eng.stats()["nn_evals"])
