"""probe: persistent pure-MCTS self-play at a given game count / lanes per game (for ncu captures)
usage: python tools/probe_selfplay.py [games] [lanes_per_game] [plies_per_game]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import alphazero_reversi_b200 as az
G = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
lpg = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ppg = int(sys.argv[3]) if len(sys.argv) > 3 else 4
eng = az.Engine(G, 100, 1, evaluator=az.EVAL_ROLLOUT, seed=1, sample_capacity=80 * G)
eng.set_lanes_per_game(lpg)
eng.selfplay(100, plies=G * 20, temperature=1.0, recycle=True)   # spread the games over the phases
torch.cuda.synchronize()
s0 = eng.stats()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); eng.selfplay(100, plies=G * ppg, temperature=1.0, recycle=True); e1.record(); torch.cuda.synchronize()
s1 = eng.stats()
ms = e0.elapsed_time(e1)
print(f"games {G} lpg {lpg}: {ms:.2f} ms  {(s1['sims'] - s0['sims']) / ms / 1e3:.1f} M sims/s  {(s1['board_steps'] - s0['board_steps']) / ms / 1e6:.2f} G board-steps/s")
