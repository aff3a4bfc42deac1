"""CPU tier: packed replay samples (alphazero-reversi_b200/replay.py) -- file format round trip,
import of the reference's game dicts (golden games recorded from the live reference SelfPlay),
checkpoint loading.  Nothing here needs the GPU."""
import os

import numpy as np
import pytest
import torch

import alphazero_reversi_b200 as az
from alphazero_reversi_b200 import replay

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "selfplay.npz")


def _golden_games():
    g = np.load(GOLD)
    games = []
    for tag in ("e0_t1", "t1_t1", "t1_t05"):
        games.append({"states": list(g[tag + "_states"]), "action_probs": list(g[tag + "_pi"]),
                      "current_players": [int(x) for x in g[tag + "_players"]], "values": [float(x) for x in g[tag + "_z"]]})
    return games


def test_from_reference_games_decodes_the_reference_dicts():
    games = _golden_games()
    s = replay.from_reference_games(games)
    assert len(s) == sum(len(g["states"]) for g in games) == 180
    # first sample of every game is the start position with black to move (board.py:31-32)
    for k in (0, 60, 120):
        assert (int(s.black[k]), int(s.white[k]), int(s.side[k])) == (replay.START_BLACK, replay.START_WHITE, 1)
    # planes 0/1 are the side to move's / opponent's discs (game.py:131-162)
    w = (np.uint64(1) << np.arange(64, dtype=np.uint64))
    i = 37
    own = int(((games[0]["states"][i][0].reshape(-1) > 0.5).astype(np.uint64) * w).sum())
    assert own == int(s.black[i] if s.side[i] == 1 else s.white[i])
    assert set(np.unique(s.z)) <= {-1, 0, 1} and s.pi.dtype == np.float32
    assert np.array_equal(s.pi[i], games[0]["action_probs"][i].astype(np.float32))


def test_replay_file_round_trip_and_errors(tmp_path):
    s = replay.from_reference_games(_golden_games())
    p = str(tmp_path / "gen0.rvsr")
    replay.save_replay(p, s)
    assert os.path.getsize(p) == 16 + len(s) * (8 + 8 + 1 + 1 + 65 * 4)   # 278 B per sample
    t = replay.load_replay(p)
    for f in ("black", "white", "side", "z", "pi"):
        assert np.array_equal(getattr(s, f), getattr(t, f)) and getattr(s, f).dtype == getattr(t, f).dtype
    # empty set, truncated file, wrong magic
    e = replay.PackedSamples(*(getattr(s, f)[:0] for f in ("black", "white", "side", "z", "pi")))
    replay.save_replay(p, e)
    assert len(replay.load_replay(p)) == 0
    replay.save_replay(p, s)
    with open(p, "r+b") as f:
        f.truncate(16 + 100)
    with pytest.raises(ValueError):
        replay.load_replay(p)
    with open(p, "wb") as f:
        f.write(b"not a replay file at all")
    with pytest.raises(ValueError):
        replay.load_replay(p)
    both = replay.PackedSamples.concat([s, s])
    assert len(both) == 2 * len(s) and np.array_equal(both.black[len(s):], s.black)


def test_checkpoint_formats(tmp_path):
    """checkpoint_XXXX.pth (dict, pipeline.py:463-480), best_model.pth (bare state_dict, :482-485) and
    the 168-key form with `_script_module.` duplicates (mcts.py:459-479) all pack to the same blob"""
    torch.manual_seed(3)
    net = az.AlphaZeroNetwork(8, 2, 64)
    sd = net.state_dict()
    ref = az.RvsNetwork.from_module(net)
    a = str(tmp_path / "checkpoint_0001.pth")
    torch.save({"iteration": 1, "model_state_dict": sd, "optimizer_state_dict": {}, "best_elo": 0.0}, a)
    b = str(tmp_path / "best_model.pth")
    torch.save(sd, b)
    c = str(tmp_path / "scripted.pth")
    dup = dict(sd)
    dup.update({"_script_module." + k: v for k, v in sd.items()})
    torch.save(dup, c)
    for path in (a, b, c):
        rn = az.RvsNetwork.from_checkpoint(path)
        assert (rn.net_blocks, rn.net_filters) == (2, 64) and torch.equal(rn.flat, ref.flat)
    assert ref.flat.numel() == sum(v.numel() for k, v in sd.items() if "num_batches_tracked" not in k)


def test_checkpoint_loader_is_safe_by_default(tmp_path):
    import torch
    import alphazero_reversi_b200 as az
    torch.manual_seed(42)
    net = az.AlphaZeroNetwork(8, 2, 64).eval()
    p1 = tmp_path / "best_model.pth"
    torch.save(net.state_dict(), p1)                                            # pipeline.py:482-485
    p2 = tmp_path / "checkpoint_0001.pth"
    torch.save({"model_state_dict": net.state_dict(), "iteration": 1, "optimizer_state_dict": {}}, p2)  # pipeline.py:463-480
    a = az.RvsNetwork.from_checkpoint(str(p1))
    b = az.RvsNetwork.from_checkpoint(str(p2))
    assert torch.equal(a.flat, b.flat) and a.net_blocks == 2 and a.net_filters == 64
    p3 = tmp_path / "module.pth"
    torch.save(net, p3)                                                         # a pickled module = code
    with pytest.raises(Exception):
        az.RvsNetwork.from_checkpoint(str(p3))
    c = az.RvsNetwork.from_checkpoint(str(p3), trusted=True)
    assert torch.equal(c.flat, a.flat)
