#!/usr/bin/env python3
"""Times the UNMODIFIED Python reference (/root/reference) on the host CPU cores of the container it
runs in -- BASELINE.md section 3's legs -- and writes profiles/python_reference_cpu_r2.json.

Test/measurement infrastructure (oracle/): the reference is pure Python, cannot travel to the GPU
box (no /root/reference there) and must not be copied into the repo, so its throughput is measured
HERE, with this committed script, and bench.py reports the committed figures as
cpu_baseline.python_reference (with where/when they were taken).  Nothing of the product imports it.

    CUDA_VISIBLE_DEVICES="" python oracle/time_python_reference.py [--quick]

Legs (reference files exercised):
  config1  ReversiGame random playouts                 src/game/game.py:36-92, src/game/board.py:70-251
  config2  MCTS(stub uniform prior + Board rollout)    src/mcts/mcts.py:322-444, 544-640  (wave 64 and 1)
  config3  SelfPlay(AlphaZeroNetwork(8,5,128), S=100)  src/self_play/self_play.py:51-145
           (i) one process, default torch threads; (ii) one process per core, torch threads = 1, summed
"""
import argparse
import contextlib
import io
import json
import multiprocessing as mp
import os
import platform
import random
import sys
import tempfile
import time

os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
REF = os.environ.get("RVS_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _imports():
    sys.path.insert(0, REF)
    import numpy as np
    import torch
    from src.game.game import ReversiGame
    from src.mcts.mcts import MCTS
    from src.model.network import AlphaZeroNetwork
    from src.self_play.self_play import SelfPlay
    return np, torch, ReversiGame, MCTS, AlphaZeroNetwork, SelfPlay


def leg_config1(n_games):
    np, torch, ReversiGame, *_ = _imports()
    rng = random.Random(12345)
    steps = 0
    t0 = time.perf_counter()
    for _ in range(n_games):
        g = ReversiGame()
        while not g.is_game_over():
            vm = g.get_valid_moves()
            g.make_move(*vm[rng.randrange(len(vm))])
            steps += 1
    dt = time.perf_counter() - t0
    return {"board_steps_per_sec": steps / dt, "games": n_games, "board_steps": steps, "seconds": dt, "cores": 1}


class RolloutStub:
    """model duck type (mcts.py:211,235,501): zero logits, value = one uniform-random playout with the
    reference Board from planes 0/1"""

    def __init__(self):
        import torch
        self._p = torch.nn.Parameter(torch.zeros(1))
        self.rng = random.Random(7)
        self.rollout_steps = 0

    def parameters(self):
        return iter([self._p])

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def predict(self, x):
        import numpy as np
        import torch
        from src.game.board import Board
        xs = x.detach().cpu().numpy()
        B = xs.shape[0]
        vals = np.zeros(B, dtype=np.float32)
        w = (1 << np.arange(64, dtype=np.uint64))
        for b in range(B):
            own = int(((xs[b, 0].reshape(-1) > 0.5).astype(np.uint64) * w).sum())
            opp = int(((xs[b, 1].reshape(-1) > 0.5).astype(np.uint64) * w).sum())
            bd = Board()
            bd.black, bd.white, bd.current_player = own, opp, 1  # the side to move plays "black"
            bd._update_board_state()
            while not bd.game_over:
                vm = bd.get_valid_moves()
                if not vm:
                    break
                bd.make_move(*vm[self.rng.randrange(len(vm))])
                self.rollout_steps += 1
            wnr = bd.winner if bd.game_over else 0
            vals[b] = 0.0 if not wnr else (1.0 if wnr == 1 else -1.0)
        return torch.zeros((B, 65)), torch.from_numpy(vals)


def leg_config2(wave, n_searches):
    np, torch, ReversiGame, MCTS, *_ = _imports()
    rng = random.Random(99)
    stub = RolloutStub()
    m = MCTS(stub, c_puct=1.0, num_simulations=100, batch_size=wave)
    sims = 0
    t0 = time.perf_counter()
    for _ in range(n_searches):
        g = ReversiGame()
        for _ in range(rng.randrange(0, 50)):  # a mid-game root
            if g.is_game_over():
                break
            vm = g.get_valid_moves()
            g.make_move(*vm[rng.randrange(len(vm))])
        if g.is_game_over():
            continue
        with contextlib.redirect_stdout(io.StringIO()):
            m.search(g)
        sims += 100
    dt = time.perf_counter() - t0
    return {"sims_per_sec": sims / dt, "wave": wave, "searches": n_searches, "seconds": dt, "cores": 1,
            "rollout_board_steps": stub.rollout_steps}


def _selfplay_worker(args):
    n_games, threads, seed, blocks, filters, sims = args
    np, torch, ReversiGame, MCTS, AlphaZeroNetwork, SelfPlay = _imports()
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(42)
    np.random.seed(seed)
    net = AlphaZeroNetwork(8, blocks, filters)
    with tempfile.TemporaryDirectory() as td, contextlib.redirect_stdout(io.StringIO()):
        sp = SelfPlay(net, {"num_simulations": sims, "c_puct": 1.0, "temperature": 1.0, "save_dir": td})
        t0 = time.perf_counter()
        games = sp.generate_games(n_games)
        dt = time.perf_counter() - t0
    plies = sum(len(g["states"]) for g in games)
    return plies, dt, torch.get_num_threads()


def leg_config3(n_games, per_core, blocks=5, filters=128, sims=100):
    if not per_core:
        plies, dt, thr = _selfplay_worker((n_games, 0, 1, blocks, filters, sims))
        return {"sims_per_sec": plies * sims / dt, "plies_per_sec": plies / dt, "games_per_sec": n_games / dt,
                "games": n_games, "plies": plies, "seconds": dt, "processes": 1, "torch_threads": thr}
    n = os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(n) as pool:
        res = pool.map(_selfplay_worker, [(n_games, 1, 100 + i, blocks, filters, sims) for i in range(n)])
    wall = time.perf_counter() - t0
    plies = sum(r[0] for r in res)
    rate = sum(r[0] * sims / r[1] for r in res)  # summed per-process rates
    return {"sims_per_sec": rate, "plies_per_sec": rate / sims, "games_per_sec": sum(n_games / r[1] for r in res),
            "games": n_games * n, "plies": plies, "seconds_wall": wall, "processes": n, "torch_threads": 1}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="smaller samples (smoke run of this script)")
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "python_reference_cpu_r2.json"))
    a = ap.parse_args()
    import numpy as np
    import torch
    q = a.quick
    out = {
        "what": "the unmodified Python reference timed on the host CPU of the BUILD container (it cannot travel to the GPU box)",
        "script": "oracle/time_python_reference.py",
        "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
        "host": {"nproc": os.cpu_count(), "cpu": _cpu_model(), "python": platform.python_version(),
                 "torch": torch.__version__, "numpy": np.__version__},
    }
    out["config1_random_playouts"] = leg_config1(20 if q else 300)
    out["config2_mcts_rollout_wave64"] = leg_config2(64, 3 if q else 30)
    out["config2_mcts_rollout_wave1"] = leg_config2(1, 3 if q else 30)
    out["config3_selfplay_5x128_one_process"] = leg_config3(1 if q else 3, per_core=False)
    out["config3_selfplay_5x128_process_per_core"] = leg_config3(1 if q else 2, per_core=True)
    if not q:
        out["config4_selfplay_20x256_800sims_one_process"] = leg_config3(1, per_core=False, blocks=20, filters=256, sims=800)
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


def _cpu_model():
    try:
        for l in open("/proc/cpuinfo"):
            if l.startswith("model name"):
                return l.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor()


if __name__ == "__main__":
    main()
