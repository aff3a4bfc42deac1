"""probe: search throughput when every game is at the same phase (no tail imbalance)"""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import alphazero_reversi_b200 as az
import orc

G = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
def run(eng, tag):
    st = torch.cuda.current_stream().cuda_stream
    eng.search(100, 1); torch.cuda.synchronize()
    s0 = eng.stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.search(100, 1)
    e1.record(); torch.cuda.synchronize()
    s1 = eng.stats(); ms = e0.elapsed_time(e1) / 5
    steps = (s1["board_steps"] - s0["board_steps"]) / 5
    print(f"{tag:28s} G={eng.n_games} {ms:8.3f} ms/search  {eng.n_games*100/ms/1e3:8.1f} Msims/s  {steps/ms/1e6:7.2f} Gsteps/s  steps/sim={steps/eng.n_games/100:.1f}")

for G in (4096, 8192, 16384):
    eng = az.Engine(G, 100, 1, evaluator=az.EVAL_ROLLOUT, seed=1)
    run(eng, "all at start")
    # advance every game by p plies with random play (same phase for all)
    for p in (20, 40, 52):
        eng.reset()
        for _ in range(p):
            eng.search(2, 1); eng.play(1.0, recycle=False)
        run(eng, f"all at ply {p}")
    eng.close()
