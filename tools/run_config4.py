"""BASELINE configs[3] at full size: ResNet 20x256, 16384 concurrent games, 800 simulations per move with
Dirichlet root noise -- one complete self-play ply (search + move sampling + sample record) on one B200.
usage: python tools/run_config4.py [plies]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import alphazero_reversi_b200 as az
plies = int(sys.argv[1]) if len(sys.argv) > 1 else 1
G, S = 16384, 800
torch.manual_seed(42)
rn = az.RvsNetwork.from_module(az.AlphaZeroNetwork(8, 20, 256).eval())
eng = az.Engine(G, S, 1, evaluator=az.EVAL_NN, c_puct=1.0, seed=1, net_blocks=20, net_filters=256)
rn.attach(eng)
eng.set_root_noise(0.03, 0.25)
eng.search(8, 1); torch.cuda.synchronize()          # warm-up (kernel setup), result discarded by the next search
s0 = eng.stats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(plies):
    eng.search(S, 1)
    eng.play(1.0, recycle=True)
e1.record(); torch.cuda.synchronize()
s1 = eng.stats()
ms = e0.elapsed_time(e1)
f = 3020931840
ev = s1["nn_evals"] - s0["nn_evals"]
out = {"config": "configs[3]: 20x256, 16384 games, 800 sims/move, Dirichlet(0.03, 0.25), wave 1, bf16", "plies": plies, "ms": ms,
       "sims_per_sec": (s1["sims"] - s0["sims"]) / ms * 1e3, "network_evals_per_sec": ev / ms * 1e3, "tflops": ev * f / ms / 1e9,
       "board_steps_played": G * plies, "overflow": s1["overflow"], "nodes_created": s1["nodes"] - s0["nodes"],
       "hbm_gb_allocated": torch.cuda.mem_get_info()[1] / 1e9 - torch.cuda.mem_get_info()[0] / 1e9}
print(json.dumps(out))
