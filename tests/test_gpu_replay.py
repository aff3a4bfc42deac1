"""GPU tier: packed samples against the trainer-format drain, the reference's game dicts (golden
games of the live reference SelfPlay) through the K3 encode kernel, and the trainer-ingest batches
(SURVEY.md 8(f) N1/N2)."""
import numpy as np
import pytest
import torch

from test_replay_cpu import _golden_games

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def az():
    import alphazero_reversi_b200 as m
    return m


def test_packed_drain_equals_trainer_format_drain(az):
    outs = []
    for packed in (False, True):
        eng = az.Engine(64, 30, 1, evaluator=az.EVAL_ROLLOUT, seed=11)
        eng.selfplay(30, plies=64 * 70, temperature=1.0, recycle=False)
        outs.append(eng.drain_packed() if packed else eng.drain_samples())
        assert eng.stats()["games_finished"] == 64
        eng.close()
    (st, pi, z), pk = outs
    assert len(pk) == len(z) > 64 * 50
    # the ring order depends on which game finishes first; compare as multisets keyed by the full row
    a = np.concatenate([st.reshape(len(z), -1), pi, z.reshape(-1, 1)], axis=1)
    b = np.concatenate([np.asarray(pk.states()).reshape(len(z), -1), pk.pi, pk.z.astype(np.float32).reshape(-1, 1)], axis=1)
    assert np.array_equal(a[np.lexsort(a.T[::-1])], b[np.lexsort(b.T[::-1])])
    assert set(np.unique(pk.side)) == {1, 2}


def test_reference_game_dicts_round_trip_through_k3(az):
    games = _golden_games()
    s = az.replay.from_reference_games(games)
    back = az.replay.to_reference_games(s)
    assert len(back) == len(games)
    for g, h in zip(games, back):
        assert h["current_players"] == g["current_players"] and h["values"] == g["values"]
        assert np.array_equal(np.array(h["states"]), np.array(g["states"]))   # planes incl. the legal-move plane
        assert np.array_equal(np.array(h["action_probs"]).astype(np.float32), np.array(g["action_probs"]).astype(np.float32))
    dev = s.to("cuda:0")
    td = az.replay.to_training_data(dev)
    assert td["states"].is_cuda and td["states"].shape == (180, 3, 8, 8) and td["value_targets"].shape == (180, 1)
    assert np.array_equal(td["states"].cpu().numpy(), np.array([x for g in games for x in g["states"]]))


def test_training_batches_feed_a_torch_step(az):
    s = az.replay.from_reference_games(_golden_games()).to("cuda:0")
    torch.manual_seed(0)
    net = az.AlphaZeroNetwork(8, 1, 64).cuda().train()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3)
    seen = 0
    g = torch.Generator(device="cuda:0").manual_seed(1)
    for states, labels, zt in az.replay.training_batches(s, 64, shuffle=True, generator=g):
        assert states.is_cuda and labels.dtype == torch.int64 and states.shape[1:] == (3, 8, 8)
        logits, v = net.predict(states)
        loss = torch.nn.functional.cross_entropy(logits, labels) + torch.nn.functional.mse_loss(v, zt)   # pipeline.py:306-327
        opt.zero_grad(); loss.backward(); opt.step()
        seen += states.shape[0]
    assert seen == 180 and torch.isfinite(loss)


def test_selfplay_mirror_reports_players_and_noise_switch(az):
    sp = az.SelfPlay(az.UniformRollout(seed=3), {"num_simulations": 20, "batch_size": 1, "temperature": 1.0, "num_parallel_games": 16,
                                                 "apply_dirichlet_noise": True, "dirichlet_alpha": 0.3, "dirichlet_epsilon": 0.25})
    games = sp.generate_games(20)
    assert len(games) == 20
    for gd in games:
        assert set(gd) == {"states", "action_probs", "current_players", "values"}
        assert gd["current_players"][0] == 1 and set(gd["current_players"]) <= {1, 2}
        assert len(gd["states"]) == len(gd["values"]) == len(gd["current_players"]) >= 50


def test_sharded_self_play_single_rank(az):
    """dist.sharded_self_play without a process group = one shard (the N > 1 path runs under torchrun:
    tools/run_sharded_selfplay.py, verified on 2 x B200 with NCCL broadcast + packed gather)"""
    from alphazero_reversi_b200 import dist as azd
    res = azd.sharded_self_play(az.UniformRollout(seed=2), {"num_simulations": 20, "batch_size": 1, "temperature": 1.0, "seed": 5}, 70,
                                slots_per_rank=32)
    starts = int(((res.black == 0x0000000810000000) & (res.white == 0x0000001008000000) & (res.side == 1)).sum())
    assert starts >= 70 and len(res) >= 70 * 50 and res.pi.is_cuda
    games = az.replay.to_reference_games(res)
    assert len(games) == starts and all(len(g["states"]) >= 50 for g in games)


def test_selfplay_train_reload_cycle(az):
    """one iteration of the reference pipeline (pipeline.py:114-150) with the engine on the self-play side:
    NN self-play -> packed samples on the device -> the trainer's batches -> optimiser steps -> the updated
    weights back into the SAME engine handle -> its predictions follow the trained torch module"""
    torch.manual_seed(7)
    net = az.AlphaZeroNetwork(8, 2, 64).cuda()
    eng = az.Engine(64, 16, 1, evaluator=az.EVAL_NN, seed=3, net_blocks=2, net_filters=64)
    az.RvsNetwork.from_module(net.eval()).attach(eng)
    eng.selfplay(16, plies=64 * 64, temperature=1.0, recycle=False)
    assert eng.stats()["games_finished"] == 64 and eng.stats()["nn_evals"] > 0
    samples = eng.drain_packed(device="cuda:0")
    assert len(samples) >= 64 * 50
    bl = np.array([0x0000000810000000], dtype=np.uint64); wh = np.array([0x0000001008000000], dtype=np.uint64); sd = np.ones(1, np.uint8)
    lg0, v0 = eng.predict(bl, wh, sd)
    net.train()
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=1e-4)       # pipeline.py:60-75
    for states, labels, zt in az.replay.training_batches(samples, 256, shuffle=True):
        logits, v = net.predict(states)
        loss = torch.nn.functional.cross_entropy(logits, labels) + torch.nn.functional.mse_loss(v, zt)
        opt.zero_grad(); loss.backward(); opt.step()
    net.eval()
    az.RvsNetwork.from_module(net).attach(eng)                                   # broadcast_weights + load in the multi-GPU case
    lg1, v1 = eng.predict(bl, wh, sd)
    with torch.no_grad():
        tl, tv = net(torch.from_numpy(az.board_ops.encode_planes(bl, wh, sd)).cuda())
    assert np.abs(lg1 - lg0).max() > 1e-3                                        # the engine really runs the new weights
    assert np.abs(lg1 - tl.cpu().numpy()).max() < 0.05 * max(1.0, float(tl.abs().max())) and abs(float(v1[0]) - float(tv[0])) < 0.05
    eng.selfplay(16, plies=64 * 8, temperature=1.0, recycle=True)                # and keeps playing
    assert eng.stats()["overflow"] == 0
    eng.close()
