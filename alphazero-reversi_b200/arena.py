"""Mirror of the reference's evaluation arena (src/arena/arena.py:19-409) on the GPU engine.

`ELORatingSystem`, `ELOPlayer`, `Arena` keep the reference's names, arguments, return values and rating
arithmetic (expected score 1/(1+10^((Rb-Ra)/400)), K = 32, ratings updated after every game in schedule
order, arena.py:46-97).  `Arena.play_game` plays one game through the MCTS mirror exactly like the
reference loop (arena.py:218-286: `get_action_probs(game, temperature=1.0)` for model players,
`random.choice` for the random player).

`Arena.run_tournament` keeps the reference's schedule (round robin, first mover alternating with
(i + j + round) % 2, arena.py:328-341) but, when every player has a built-in evaluator (RvsNetwork,
UniformRollout, UniformDiscDiff, or none = random mover), plays ALL games of the tournament concurrently:
the outcome of a game does not depend on the ratings, so the games run as batched searches (one engine per
player: every ply each engine searches the positions in which its player is to move) and the ELO updates
are applied afterwards in the reference's game order.  Move sampling draws from numpy's global generator
(one `random_sample` per move, inverse CDF over pi like np.random.choice), so a seeded tournament is
reproducible; it does not consume the generator in the reference's interleaving.
"""
import json
import os
import random
import time
from datetime import datetime
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import _lib as L
from . import board_ops
from .engine import Engine
from .game import ReversiGame
from .mcts import MCTS


class ELORatingSystem:
    def __init__(self, k: float = 32, initial_rating: float = 1500.0):
        self.k = k
        self.initial_rating = initial_rating
        self.ratings: Dict[str, float] = {}
        self.games_played: Dict[str, int] = {}
        self.history: List[Dict] = []

    def add_player(self, player_id: str, rating: Optional[float] = None):
        if player_id not in self.ratings:
            self.ratings[player_id] = rating if rating is not None else self.initial_rating
            self.games_played[player_id] = 0

    def get_rating(self, player_id: str) -> float:
        return self.ratings.get(player_id, self.initial_rating)

    def get_expected_score(self, rating_a: float, rating_b: float) -> float:
        return 1.0 / (1.0 + 10.0 ** ((rating_b - rating_a) / 400.0))  # arena.py:46-48

    def update_ratings(self, player_a: str, player_b: str, score_a: float):
        self.add_player(player_a)
        self.add_player(player_b)
        ra, rb = self.ratings[player_a], self.ratings[player_b]
        ea = self.get_expected_score(ra, rb)
        eb = 1.0 - ea
        na = ra + self.k * (score_a - ea)           # arena.py:71-72
        nb = rb + self.k * ((1 - score_a) - eb)
        self.ratings[player_a], self.ratings[player_b] = na, nb
        self.games_played[player_a] += 1
        self.games_played[player_b] += 1
        rec = {"timestamp": time.time(), "player_a": player_a, "player_b": player_b, "score_a": score_a,
               "score_b": 1.0 - score_a, "rating_a_before": ra, "rating_b_before": rb, "rating_a_after": na,
               "rating_b_after": nb}
        self.history.append(rec)
        return rec

    def get_leaderboard(self) -> List[Dict]:
        board = [{"player_id": p, "rating": r, "games_played": self.games_played[p]} for p, r in self.ratings.items()]
        board.sort(key=lambda x: x["rating"], reverse=True)
        return board

    def save_ratings(self, filepath: str):
        with open(filepath, "w") as f:
            json.dump({"k": self.k, "initial_rating": self.initial_rating, "ratings": self.ratings,
                       "games_played": self.games_played, "history": self.history,
                       "last_updated": datetime.now().isoformat()}, f, indent=2)

    @classmethod
    def load_ratings(cls, filepath: str) -> "ELORatingSystem":
        with open(filepath) as f:
            data = json.load(f)
        elo = cls(k=data["k"], initial_rating=data["initial_rating"])
        elo.ratings = {k: float(v) for k, v in data["ratings"].items()}
        elo.games_played = {k: int(v) for k, v in data["games_played"].items()}
        elo.history = data.get("history", [])
        return elo


def _cuda_index(device):
    """'cuda' / 'cuda:1' / torch.device / int -> CUDA device index for the engines (None = the current device)"""
    if device is None or isinstance(device, int):
        return device
    d = str(device)  # 'cuda', 'cuda:1', or str(torch.device)
    return int(d.split(":", 1)[1]) if ":" in d else None


class ELOPlayer:
    """player_id + model (None = uniform-random mover) + MCTS parameters (arena.py:137-197)"""

    def __init__(self, player_id: str, model=None, mcts_params: Optional[Dict] = None, device: str = "cuda"):
        self.player_id = player_id
        self.model = model
        self.device = device
        self.mcts = None
        self.mcts_params = dict(mcts_params) if mcts_params is not None else {"num_simulations": 800, "c_puct": 1.0, "temperature": 1.0}
        if model is not None:
            if getattr(model, "evaluator", None) is None:  # torch module: the reference's eval()/to(device)
                model.eval()
                model.to(device)
            self.mcts = MCTS(model=model, c_puct=self.mcts_params.get("c_puct", 1.0),
                             num_simulations=self.mcts_params.get("num_simulations", 800),
                             batch_size=self.mcts_params.get("batch_size", 64), device=_cuda_index(device))

    def get_move(self, game: ReversiGame) -> Tuple[int, int]:
        if self.model is None:
            valid = game.get_valid_moves()
            return random.choice(valid) if valid else (-1, -1)
        action, _ = self.mcts.get_action_probs(game, temperature=1.0)  # arena.py:183-186
        return action

    def reset(self):
        if self.mcts is not None:
            self.mcts.update_with_move(None)


class Arena:
    def __init__(self, elo_system: Optional[ELORatingSystem] = None):
        self.elo = elo_system if elo_system is not None else ELORatingSystem()
        self.players: Dict[str, ELOPlayer] = {}

    def add_player(self, player: ELOPlayer):
        self.players[player.player_id] = player
        self.elo.add_player(player.player_id)

    # ------------------------------------------------------------------ one game, reference loop
    def play_game(self, player1_id: str, player2_id: str, verbose: bool = False, print_games: bool = False) -> float:
        if player1_id not in self.players or player2_id not in self.players:
            raise ValueError(f"One or both players not found: {player1_id}, {player2_id}")
        p1, p2 = self.players[player1_id], self.players[player2_id]
        p1.reset()
        p2.reset()
        game = ReversiGame()
        while not game.is_game_over():  # arena.py:249-267
            cur = p1 if game.current_player == 1 else p2
            move = cur.get_move(game)
            if move == (-1, -1):
                raise RuntimeError("a pass was selected: the reference would call the missing game.pass_turn() here "
                                   "(arena.py:257); auto-pass makes this unreachable")
            game.make_move(*move)
            if verbose or print_games:
                print(f"{cur.player_id} plays at {move}")
        b, w = game.get_score()
        return 1.0 if b > w else (0.0 if w > b else 0.5)

    # ------------------------------------------------------------------ all games of a schedule at once
    def _batchable(self) -> bool:
        return all(p.model is None or getattr(p.model, "evaluator", None) is not None for p in self.players.values())

    def play_games_batched(self, schedule: List[Tuple[str, str]]) -> List[float]:
        """results (1 / 0.5 / 0 from the first player's view) of every (black, white) pairing, played concurrently"""
        n = len(schedule)
        if n == 0:
            return []
        ids = list(self.players.keys())
        black_p = np.array([ids.index(a) for a, _ in schedule])
        white_p = np.array([ids.index(b) for _, b in schedule])
        bl = np.full(n, 0x0000000810000000, dtype=np.uint64)
        wh = np.full(n, 0x0000001008000000, dtype=np.uint64)
        sd = np.ones(n, dtype=np.uint8)
        fl = np.zeros(n, dtype=np.uint8)
        engines: Dict[int, Engine] = {}
        try:
            for pi, pid in enumerate(ids):
                pl = self.players[pid]
                cap = int(((black_p == pi) | (white_p == pi)).sum())
                if pl.model is None or cap == 0:
                    continue
                S = pl.mcts_params.get("num_simulations", 800)
                K = max(1, pl.mcts_params.get("batch_size", 64))
                eng = Engine(cap, S, K, evaluator=pl.model.evaluator, c_puct=pl.mcts_params.get("c_puct", 1.0),
                             seed=getattr(pl.model, "seed", 0) + 7919 * pi, device=_cuda_index(pl.device),
                             net_blocks=getattr(pl.model, "net_blocks", 0),
                             net_filters=getattr(pl.model, "net_filters", 0))
                if hasattr(pl.model, "attach"):
                    pl.model.attach(eng)
                engines[pi] = eng
            for _ply in range(130):  # a game has at most 60 moves; the bound only guards against a logic error
                live = (fl & 1) == 0
                if not live.any():
                    break
                mover = np.where(sd == 1, black_p, white_p)
                moves = np.full(n, 255, dtype=np.uint8)
                for pi, pid in enumerate(ids):
                    idx = np.nonzero(live & (mover == pi))[0]
                    if len(idx) == 0:
                        continue
                    pl = self.players[pid]
                    if pl.model is None:  # uniform random legal move (arena.py:176-179)
                        lm = board_ops.legal_masks(bl[idx], wh[idx], sd[idx])
                        for j, g in enumerate(idx):
                            sq = [q for q in range(64) if (int(lm[j]) >> q) & 1]
                            moves[g] = sq[int(np.random.random_sample() * len(sq))]
                        continue
                    eng = engines[pi]
                    eng.set_positions(bl[idx], wh[idx], sd[idx])
                    eng.search(pl.mcts_params.get("num_simulations", 800), max(1, pl.mcts_params.get("batch_size", 64)))
                    v = eng.root_visits(len(idx)).astype(np.float64)
                    st = eng.stats()
                    if st["overflow"]:
                        raise L.RvsError(f"engine error counters non-zero: {st}")
                    for j, g in enumerate(idx):
                        tot = v[j].sum()
                        if tot <= 0:  # num_simulations <= wave: no child visited (SURVEY.md 8(a) A7 hazard)
                            raise L.RvsError("search returned no visits: num_simulations must exceed the MCTS wave size")
                        cdf = np.cumsum(v[j] / tot)   # np.random.choice: cumsum, normalise, searchsorted(right)
                        cdf /= cdf[-1]
                        moves[g] = min(int(np.searchsorted(cdf, np.random.random_sample(), side="right")), 63)
                idx = np.nonzero(moves != 255)[0]
                b2, w2, s2, f2 = bl[idx].copy(), wh[idx].copy(), sd[idx].copy(), fl[idx].copy()
                ok, _ = board_ops.apply_moves(b2, w2, s2, f2, moves[idx], want_legal=False)
                if not ok.all():
                    raise L.RvsError("arena: an illegal move was selected")
                bl[idx], wh[idx], sd[idx], fl[idx] = b2, w2, s2, f2
            else:
                raise L.RvsError("arena: games did not terminate")
        finally:
            for e in engines.values():
                e.close()
        nb = np.array([bin(int(x)).count("1") for x in bl])
        nw = np.array([bin(int(x)).count("1") for x in wh])
        return [1.0 if a > b else (0.0 if b > a else 0.5) for a, b in zip(nb, nw)]

    # ------------------------------------------------------------------ reference API
    def run_tournament(self, rounds: int = 100, verbose: bool = False, print_games: bool = False, batched: Optional[bool] = None) -> Dict:
        player_ids = list(self.players.keys())
        n = len(player_ids)
        if n < 2:
            raise ValueError("Need at least 2 players for a tournament")
        results = {"games_played": 0, "matchups": {}, "start_time": time.time(), "end_time": None, "rounds": []}
        for i in range(n):
            for j in range(i + 1, n):
                p1, p2 = player_ids[i], player_ids[j]
                results["matchups"][f"{p1}_vs_{p2}"] = {"player1": p1, "player2": p2, "games_played": 0, "wins1": 0, "wins2": 0, "draws": 0}
        schedule = []  # (round, first mover, second mover) in the reference's order (arena.py:328-341)
        for r in range(rounds):
            for i in range(n):
                for j in range(i + 1, n):
                    p1, p2 = player_ids[i], player_ids[j]
                    if (i + j + r) % 2 == 0:
                        p1, p2 = p2, p1
                    schedule.append((r, p1, p2))
        if batched is None:
            batched = self._batchable()
        outcomes = self.play_games_batched([(a, b) for _, a, b in schedule]) if batched else None
        cur_round = None
        for gi, (r, p1, p2) in enumerate(schedule):
            if cur_round is None or cur_round["round"] != r + 1:
                cur_round = {"round": r + 1, "games": []}
                results["rounds"].append(cur_round)
            result = outcomes[gi] if batched else self.play_game(p1, p2, verbose=verbose, print_games=print_games)
            rec = self.elo.update_ratings(p1, p2, result)
            key = f"{p1}_vs_{p2}" if f"{p1}_vs_{p2}" in results["matchups"] else f"{p2}_vs_{p1}"
            m = results["matchups"][key]
            m["games_played"] += 1
            results["games_played"] += 1
            # wins1 / wins2 count wins of the game's first / second mover, as the reference does (arena.py:353-358)
            if result == 1.0:
                m["wins1"] += 1
            elif result == 0.0:
                m["wins2"] += 1
            else:
                m["draws"] += 1
            cur_round["games"].append({"player1": p1, "player2": p2, "result": result,
                                       "elo1_before": rec["rating_a_before"], "elo2_before": rec["rating_b_before"],
                                       "elo1_after": rec["rating_a_after"], "elo2_after": rec["rating_b_after"]})
        results["end_time"] = time.time()
        results["duration"] = results["end_time"] - results["start_time"]
        results["leaderboard"] = self.elo.get_leaderboard()
        if verbose or print_games:
            self.print_leaderboard()
        return results

    def print_leaderboard(self):
        print("\nCurrent Leaderboard:")
        print("Rank  Player ID               Rating  Games Played")
        print("----  ---------------------  -------  ------------")
        for i, p in enumerate(self.elo.get_leaderboard(), 1):
            print(f"{i:4d}  {p['player_id']:22s}  {p['rating']:7.1f}  {p['games_played']:12d}")

    def save_results(self, filepath: str):
        self.elo.save_ratings(os.path.splitext(filepath)[0] + "_elo.json")
        with open(filepath, "w") as f:
            json.dump(self.elo.get_leaderboard(), f, indent=2)
