"""one-off large parity sweep: every game of a big batched wave-1 search against the oracle"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
import alphazero_reversi_b200 as az
from test_gpu_mcts import _random_roots
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
bl, wh, sd = _random_roots(n, 777)
bad = 0
for lpg, ev, S in ((8, 1, 100), (4, 1, 100), (2, 0, 100), (8, 0, 200)):
    e = az.Engine(n, S, 1, evaluator=ev, seed=1234 + lpg)
    e.set_lanes_per_game(lpg)
    e.set_positions(bl, wh, sd)
    e.search(S, 1)
    v = e.root_visits()
    t0 = time.time()
    for g in range(n):
        ov, *_ = orc.mcts_search((int(bl[g]), int(wh[g]), int(sd[g])), S, 1, evaluator=ev, seed=1234 + lpg, game_id=g)
        bad += int(not np.array_equal(v[g], ov))
    print(f"lpg {lpg} evaluator {ev} S {S}: {n} games compared, mismatches so far {bad} ({time.time() - t0:.1f} s oracle)")
    e.close()
assert bad == 0
print("parity sweep ok")
