"""Drop-in mirror of the reference's self-play generator (src/self_play/self_play.py:15-219).

`SelfPlay(model, args).generate_games(n)` returns the same list of dicts
(`states`, `action_probs`, `current_players`, `values`) and `generate_training_data(n)` the same
stacked arrays the trainer consumes (src/trainer/pipeline.py:163-169, 226-228).

Two execution modes, chosen by args['num_parallel_games'] (an argument the reference accepts and
ignores, self_play.py:30):
  * absent / 1  -> game-by-game through the MCTS mirror with numpy move sampling: reproduces the
                   reference bit for bit under the same np.random seed.
  * > 1         -> lockstep batched self-play on the device (Engine.search + Engine.play):
                   thousands of games advance one ply per step, samples stay in HBM until drained.
"""
import os
import time
from typing import Dict, List

import numpy as np

from . import _lib as L
from .engine import Engine
from .game import ReversiGame
from .mcts import MCTS
from .replay import to_reference_games


class SelfPlay:
    def __init__(self, model, args: dict):
        self.model = model
        self.args = args
        self._builtin = getattr(model, "evaluator", None)
        if self._builtin is None:
            self.device = next(model.parameters()).device
            self.model.to(self.device)
            self.model.eval()
        else:
            self.device = None
        self.cuda_device = args.get("device", None)  # CUDA device index of the engines (None: the current device)
        self.mcts = MCTS(model=model, c_puct=args.get("c_puct", 1.0),
                         num_simulations=args.get("num_simulations", 800),
                         batch_size=args.get("batch_size", 64), device=self.cuda_device,
                         search_mode=args.get("search_mode", L.MODE_REF))
        self.save_dir = args.get("save_dir", None)
        if self.save_dir:
            os.makedirs(self.save_dir, exist_ok=True)
        self.verbose = args.get("verbose", False)

    # ------------------------------------------------------------------ sequential (bit-exact)
    def _play_one(self) -> Dict:
        game = ReversiGame()
        gd = {"states": [], "action_probs": [], "current_players": [], "values": []}
        T = self.args.get("temperature", 1.0)
        while not game.is_game_over():  # self_play.py:80-101
            action, action_probs = self.mcts.get_action_probs(game, temperature=T)
            gd["states"].append(game.get_canonical_state())
            gd["current_players"].append(game.current_player)
            gd["action_probs"].append(action_probs)
            row, col = action
            if not game.make_move(row, col):
                # the reference would loop forever here; fail loudly instead
                raise RuntimeError(f"self-play selected an illegal move {action}; "
                                   "num_simulations must exceed the MCTS wave size")
            self.mcts.update_with_move(action)
        winner = game.get_winner()
        for player in gd["current_players"]:  # self_play.py:117-126
            gd["values"].append(0.0 if winner == 0 else (1.0 if player == winner else -1.0))
        return gd

    # ------------------------------------------------------------------ batched (device)
    def _play_batched(self, num_games: int, parallel: int) -> List[Dict]:
        if self._builtin is None:
            raise L.RvsError("batched self-play needs a built-in evaluator (UniformDiscDiff, UniformRollout or "
                             "RvsNetwork); wrap external torch models with MCTS/SelfPlay in sequential mode")
        S = self.args.get("num_simulations", 800)
        K = self.args.get("batch_size", 64)
        T = self.args.get("temperature", 1.0)
        slots = min(parallel, num_games)
        eng = Engine(slots, S, K, evaluator=self._builtin, c_puct=self.args.get("c_puct", 1.0),
                     seed=self.args.get("seed", getattr(self.model, "seed", 0)), device=self.cuda_device,
                     sample_capacity=64 * max(slots, 1) * 2, net_blocks=getattr(self.model, "net_blocks", 0),
                     net_filters=getattr(self.model, "net_filters", 0))
        if hasattr(self.model, "attach"):
            self.model.attach(eng)
        if self.args.get("apply_dirichlet_noise", False):
            # engine feature: the reference accepts dirichlet_alpha / dirichlet_epsilon and never applies
            # them (self_play.py:18-47), so the noise needs this explicit switch
            eng.set_root_noise(self.args.get("dirichlet_alpha", 0.3), self.args.get("dirichlet_epsilon", 0.25))
        # exactly `num_games` games are started (ids 0 .. num_games-1): a slot restarts only while its next game id
        # stays below the limit, so no surplus game is searched and the set returned does not depend on which
        # games happen to finish first (the reference plays its games one after another, self_play.py:66)
        eng.set_option(L.OPT_GAME_LIMIT, num_games)
        fast = self.args.get("search_mode", L.MODE_REF) == L.MODE_FAST
        if fast:
            eng.set_search_mode(L.MODE_FAST)
        persistent = not fast and K == 1 and self._builtin in (L.EVAL_E0, L.EVAL_ROLLOUT, L.EVAL_NN)
        games: List[Dict] = []
        collected = 0
        while len(games) < num_games:
            if persistent:  # one work-conserving launch plays 8 plies of every slot (rvs_engine_selfplay)
                eng.selfplay(S, plies=16 * slots, temperature=T, recycle=True)
            else:
                eng.search(S, K)
                eng.play(T, recycle=True)  # finished slots restart at once, up to the game limit
            st = eng.stats()
            if st["overflow"] or st["stalled"] or st["samples_dropped"]:
                raise L.RvsError(f"engine error counters non-zero: {st}")
            if st["games_finished"] > collected:
                collected = st["games_finished"]
                games.extend(to_reference_games(eng.drain_packed()))
        eng.close()
        return games[:num_games]

    # ------------------------------------------------------------------ reference API
    def generate_games(self, num_games: int) -> List[Dict]:
        parallel = int(self.args.get("num_parallel_games", 1) or 1)
        if parallel > 1:
            all_games = self._play_batched(num_games, parallel)
        else:
            all_games = []
            for game_idx in range(num_games):
                t0 = time.time()
                gd = self._play_one()
                all_games.append(gd)
                if self.verbose:
                    print(f"Game {game_idx + 1} completed in {time.time() - t0:.1f}s, {len(gd['states'])} states")
        if self.save_dir:
            import torch
            from datetime import datetime
            ts = datetime.now().strftime("%Y%m%d_%H%M%S")
            for i, gd in enumerate(all_games):  # self_play.py:129-131
                torch.save(gd, os.path.join(self.save_dir, f"game_{ts}_{i}.pt"))
        return all_games

    def generate_training_data(self, num_games: int):
        games = self.generate_games(num_games)
        all_states, all_probs, all_values = [], [], []
        for g in games:
            all_states.extend(g.get("states", []))
            all_probs.extend(g.get("action_probs", []))
            all_values.extend(g.get("values", []))
        if not all_states or not all_probs or not all_values:
            return None
        return {"states": np.array(all_states, dtype=np.float32),
                "action_probs": np.array(all_probs, dtype=np.float32),
                "values": np.array(all_values, dtype=np.float32).reshape(-1, 1)}
