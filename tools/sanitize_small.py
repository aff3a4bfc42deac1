"""small end-to-end run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import alphazero_reversi_b200 as az
which = sys.argv[1] if len(sys.argv) > 1 else "tree"
if which == "tree":
    for lpg in (8, 4, 2):
        e = az.Engine(13, 24, 8, evaluator=az.EVAL_ROLLOUT, seed=lpg)
        e.set_lanes_per_game(lpg)
        e.set_root_noise(0.3, 0.25)
        e.search(24, 1); e.root_visits()
        e.selfplay(24, plies=13 * 70, temperature=1.0, recycle=False)
        assert e.stats()["games_finished"] == 13 and e.stats()["overflow"] == 0
        e.drain_packed(); e.search(24, 8); e.play(1.0); e.close()
    az.board_ops.random_playouts(1000, 3); assert az.board_ops.perft(5) == 1396
    print("tree ok")
else:
    torch.manual_seed(1)
    for nb, nf in ((1, 64), (1, 128), (1, 256)):
        rn = az.RvsNetwork.from_module(az.AlphaZeroNetwork(8, nb, nf).eval())
        e = az.Engine(9, 12, 4, evaluator=az.EVAL_NN, net_blocks=nb, net_filters=nf)
        rn.attach(e)
        e.search(12, 1); e.search(12, 4); e.play(1.0)
        bl = np.full(5, 0x0000000810000000, np.uint64); wh = np.full(5, 0x0000001008000000, np.uint64)
        e.predict(bl, wh, np.ones(5, np.uint8)); e.close()
    print("net ok")
