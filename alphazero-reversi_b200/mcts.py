"""Drop-in mirror of the reference's MCTS (src/mcts/mcts.py:191-719) on the GPU engine.

`MCTS(model, c_puct, num_simulations, batch_size, ...)` keeps the reference signature and
semantics -- including its wave behaviour (all sims of a wave share stale cached UCB scores,
unvisited children score +inf) -- because parity is defined as bit-exact visit
counts.  Selection, expansion and backup run in the CUDA tree kernels; `model` is only asked for
`predict(x)` on the leaf batch, exactly where the reference calls it (mcts.py:501).

Built-in evaluators avoid the model round trip entirely: pass `UniformDiscDiff()` (E0),
`UniformRollout()` or an `RvsNetwork` (K4) as the model.
"""
from typing import Dict, Optional, Tuple

import numpy as np

from . import _lib as L
from .engine import Engine


class UniformDiscDiff:
    """Deterministic evaluator E0: logits == 0, value = (own - opp) / 64."""
    evaluator = L.EVAL_E0


class UniformRollout:
    """BASELINE config 2: uniform prior, value = one uniform random playout from the leaf."""
    evaluator = L.EVAL_ROLLOUT

    def __init__(self, seed: int = 0):
        self.seed = seed


class _Root:
    """what callers read of mcts.root (visit_count / children visit counts)"""

    def __init__(self, visits):
        self.children = {divmod(i, 8): int(n) for i, n in enumerate(visits[:64]) if n}
        self.visit_count = int(sum(visits))


class MCTS:
    def __init__(self, model, c_puct: float = 1.0, num_simulations: int = 800, batch_size: int = 64,
                 num_threads: int = 1, use_transposition: bool = True, rules: int = L.RULES_REF,
                 device=None, search_mode: int = L.MODE_REF):
        self.model = model
        self.c_puct = c_puct
        self.num_simulations = num_simulations
        self.batch_size = batch_size
        self.root = None
        self.lock = None
        self.use_transposition = use_transposition  # dead code in the reference
        self.rules = rules
        # L.MODE_REF: the reference's wave semantics, bugs included (graded).  L.MODE_FAST (engine feature): virtual-loss
        # PUCT whose waves spread over distinct leaves -- the mode in which batch_size > 1 buys search quality
        self.search_mode = search_mode
        self._cuda_device = device
        self._engine = None
        self._engine_key = None
        self._builtin = getattr(model, "evaluator", None)
        if self._builtin is None:
            import torch  # noqa: F401  (external models are torch modules / duck types)
            self.device = next(model.parameters()).device  # mcts.py:211
            model.eval()                                   # mcts.py:235
        else:
            self.device = None

    # ------------------------------------------------------------------------------------
    def _get_engine(self) -> Engine:
        key = (self.num_simulations, self.batch_size, float(self.c_puct), self.search_mode)
        if self._engine is None or self._engine_key != key:
            if self._engine is not None:
                self._engine.close()
            ev = self._builtin if self._builtin is not None else L.EVAL_EXTERNAL
            self._engine = Engine(1, self.num_simulations, max(1, self.batch_size), evaluator=ev,
                                  c_puct=self.c_puct, rules=self.rules, seed=getattr(self.model, "seed", 0),
                                  device=self._cuda_device, net_blocks=getattr(self.model, "net_blocks", 0),
                                  net_filters=getattr(self.model, "net_filters", 0))
            if hasattr(self.model, "attach"):
                self.model.attach(self._engine)
            if self.search_mode != L.MODE_REF:
                self._engine.set_search_mode(self.search_mode)
            self._engine_key = key
        return self._engine

    def _predict(self, planes):
        """model.predict on the leaf batch + softmax, as mcts.py:591-597"""
        import torch
        import torch.nn.functional as F
        with torch.no_grad():
            x = planes if isinstance(planes, torch.Tensor) else torch.tensor(planes, dtype=torch.float32)
            logits, values = self.model.predict(x.to(self.device))
            probs = F.softmax(logits, dim=1)
        return probs, values

    def search(self, game) -> Dict[Tuple[int, int], int]:
        """Run MCTS from `game` (not mutated) and return {(row, col): visit_count} (mcts.py:322-407)."""
        eng = self._get_engine()
        eng.set_positions(np.array([game.board.black], dtype=np.uint64), np.array([game.board.white], dtype=np.uint64),
                          np.array([game.current_player], dtype=np.uint8))
        S, K = self.num_simulations, max(1, self.batch_size)
        if self._builtin is not None:
            eng.search(S, K)
        else:
            on_gpu = self.device is not None and self.device.type == "cuda"
            eng.begin_search()
            start = 0
            while start < S:  # mcts.py:348-349 (FAST mode: the first wave is one simulation, it expands the root)
                k = 1 if (self.search_mode == L.MODE_FAST and start == 0) else min(K, S - start)
                start += k
                eng.select(k)
                planes, valid = eng.leaf_planes(device=self.device if on_gpu else None)
                idx = valid.nonzero().flatten() if on_gpu else np.nonzero(valid)[0]
                if on_gpu:
                    import torch
                    probs = torch.zeros((k, 65), dtype=torch.float32, device=self.device)
                    values = torch.zeros(k, dtype=torch.float32, device=self.device)
                    if idx.numel():
                        p, v = self._predict(planes[idx])
                        probs[idx] = p.float()
                        values[idx] = v.float().reshape(-1)
                else:
                    probs = np.zeros((k, 65), dtype=np.float32)
                    values = np.zeros(k, dtype=np.float32)
                    if len(idx):
                        p, v = self._predict(planes[idx])
                        probs[idx] = p.cpu().numpy()
                        values[idx] = v.cpu().numpy().reshape(-1)
                eng.process(probs, values)
        visits = eng.root_visits(1)[0]
        st = eng.stats()
        if st["overflow"] or st["stalled"] or st["samples_dropped"]:
            raise L.RvsError("MCTS engine reported a node-pool overflow / stalled slot")
        self.root = _Root(visits)
        # children exist for every legal square once the root was expanded (mcts.py:406-407)
        if S >= 1:
            legal = game.get_valid_moves()
            return {mv: int(visits[mv[0] * 8 + mv[1]]) for mv in legal}
        return {}

    def get_action_probs(self, game, temperature: float = 1.0) -> Tuple[Tuple[int, int], np.ndarray]:
        """(action, pi[65]) exactly as mcts.py:642-694, including the global numpy RNG draw."""
        visit_counts = self.search(game)
        board_size = game.size
        action_probs = np.zeros(board_size * board_size + 1)
        total_visits = sum(visit_counts.values())
        if total_visits > 0:
            for (row, col), count in visit_counts.items():
                if (row, col) == (-1, -1):
                    action_probs[-1] = count / total_visits
                else:
                    action_probs[row * board_size + col] = count / total_visits
        if temperature > 0 and not np.all(action_probs == 0):
            temp_probs = action_probs ** (1.0 / temperature)
            action_probs = temp_probs / np.sum(temp_probs)
        if temperature == 0.0 or np.all(action_probs == 0):
            best_action_idx = np.argmax(action_probs)
        else:
            best_action_idx = np.random.choice(len(action_probs), p=action_probs)
        if best_action_idx == len(action_probs) - 1:
            best_action = (-1, -1)
        else:
            best_action = (best_action_idx // board_size, best_action_idx % board_size)
        return best_action, action_probs

    def update_with_move(self, move: Optional[Tuple[int, int]] = None):
        """The reference prunes the subtree, but search() rebuilds the root anyway (mcts.py:334)."""
        self.root = None
