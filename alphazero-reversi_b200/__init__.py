"""B200-native Reversi self-play hot path behind the reference's Python API.

    from alphazero_reversi_b200 import ReversiGame, MCTS, SelfPlay      # drop-in mirrors
    from alphazero_reversi_b200 import Engine, board_ops                 # batched device API

All computation happens in librvs_b200.so (hand-written sm_100a CUDA, C ABI in include/rvs_b200.h).
There is no CPU fallback: importing is cheap, the first call raises if the library or GPU is absent.
"""
from . import _lib, board_ops
from ._lib import (EVAL_E0, EVAL_EXTERNAL, EVAL_NN, EVAL_ROLLOUT, MODE_FAST, MODE_REF, RULES_REF, RULES_STRICT, RvsError)
from .engine import Engine
from .game import Board, ReversiGame
from .mcts import MCTS, UniformDiscDiff, UniformRollout
from .self_play import SelfPlay
from . import replay
from .arena import Arena, ELOPlayer, ELORatingSystem
from .replay import PackedSamples


def __getattr__(name):  # torch-dependent members are imported lazily
    if name in ("AlphaZeroNetwork", "RvsNetwork", "pack_state_dict", "network"):
        import importlib
        network = importlib.import_module(__name__ + ".network")
        return network if name == "network" else getattr(network, name)
    raise AttributeError(name)

__all__ = ["Board", "ReversiGame", "MCTS", "SelfPlay", "Engine", "board_ops", "UniformDiscDiff",
           "UniformRollout", "RvsError", "RULES_REF", "RULES_STRICT", "EVAL_E0", "EVAL_ROLLOUT",
           "EVAL_EXTERNAL", "EVAL_NN", "MODE_REF", "MODE_FAST", "replay", "PackedSamples", "Arena", "ELOPlayer", "ELORatingSystem"]
