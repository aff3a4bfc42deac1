"""GPU tier: self-play sample emission.  (a) the SelfPlay mirror reproduces recorded reference
games under the same numpy seed; (b) the batched device self-play equals the C oracle's
orc_self_play_game sample for sample (states, pi as f32, z)."""
import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def az():
    import alphazero_reversi_b200 as m
    return m


@pytest.mark.parametrize("tag,kind", [("e0_t1", "E0"), ("t1_t1", "T1"), ("t1_t05", "T1")])
def test_selfplay_mirror_reproduces_reference_game(az, golden, tag, kind):
    from stubs import StubModel
    gs = golden["selfplay"]
    S, K, T, seed = gs[f"{tag}_cfg"]
    np.random.seed(int(seed))
    sp = az.SelfPlay(StubModel(kind), {"num_simulations": int(S), "c_puct": 1.0, "temperature": float(T)})
    gd = sp.generate_games(1)[0]
    assert np.array_equal(np.array(gd["states"], dtype=np.float32), gs[f"{tag}_states"])
    assert np.array_equal(np.array(gd["current_players"]), gs[f"{tag}_players"])
    assert np.array_equal(np.array(gd["values"], dtype=np.float32), gs[f"{tag}_z"])
    pi = np.array(gd["action_probs"], dtype=np.float64)
    if T == 1.0:
        assert np.array_equal(pi, gs[f"{tag}_pi"])
    else:
        assert np.allclose(pi, gs[f"{tag}_pi"], rtol=1e-14, atol=0)


@pytest.mark.parametrize("evaluator,S,K,T", [(0, 100, 1, 1.0), (1, 100, 64, 1.0), (1, 60, 1, 0.0), (0, 130, 16, 1.0)])
def test_batched_selfplay_vs_oracle(az, evaluator, S, K, T):
    n = 48
    eng = az.Engine(n, S, K, evaluator=evaluator, seed=99)
    moves_log = []
    for ply in range(64):
        eng.search(S, K)
        mv = eng.play(T, recycle=False, want_moves=True)
        moves_log.append(mv.copy())
        if (mv == 255).all():
            break
    st = eng.stats()
    assert st["overflow"] == 0 and st["games_finished"] == n
    states, pi, z = eng.drain_samples()
    assert len(states) == st["samples"]
    # the ring holds whole games (ply order) in completion order: index them by their move lists
    got = {}
    i = 0
    while i < len(states):
        j = i + 1
        while j < len(states) and not (states[j][0].sum() == 2 and states[j][1].sum() == 2):
            j += 1
        got[states[i:j].tobytes()] = (pi[i:j], z[i:j])
        i = j
    assert len(got) <= n
    moves_log = np.array(moves_log)
    for g in range(n):
        samples, win = orc.self_play_game(S, K, evaluator=evaluator, seed=99, game_id=g, temperature=T)
        mine = [int(m) for m in moves_log[:, g] if m != 255]
        assert mine == [s.move for s in samples], g
        st_exp = np.array([orc.planes(s.black, s.white, s.side) for s in samples], dtype=np.float32)
        key = st_exp.tobytes()
        assert key in got, g
        gpi, gz = got[key]
        pi_exp = np.array([orc.action_probs(np.array(s.visits[:]), T).astype(np.float32) for s in samples])
        assert np.array_equal(gpi, pi_exp)
        assert np.array_equal(gz, np.array([s.z for s in samples], dtype=np.float32))


def test_batched_selfplay_api_with_recycling(az):
    """SelfPlay(model, args) with num_parallel_games > 1: reference-format output from the device loop"""
    sp = az.SelfPlay(az.UniformRollout(seed=5), {"num_simulations": 40, "batch_size": 8, "temperature": 1.0,
                                                  "num_parallel_games": 64})
    data = sp.generate_training_data(150)
    assert data["states"].shape[1:] == (3, 8, 8) and data["action_probs"].shape[1] == 65
    assert data["values"].shape == (len(data["states"]), 1)
    assert np.allclose(data["action_probs"].sum(axis=1), 1.0, atol=1e-5)
    assert set(np.unique(data["values"])) <= {-1.0, 0.0, 1.0}
    # pi mass only on legal squares (plane 2)
    legal = data["states"][:, 2].reshape(-1, 64) > 0.5
    assert not (data["action_probs"][:, :64][~legal] > 0).any()


@pytest.mark.parametrize("evaluator,S,T", [(1, 100, 1.0), (0, 60, 1.0), (1, 40, 0.0)])
def test_persistent_selfplay_kernel_vs_oracle(az, evaluator, S, T):
    """rvs_engine_selfplay (one work-conserving launch, games desynchronised) must emit exactly the
    samples of the lockstep path / the oracle: per-game RNG streams do not depend on scheduling"""
    n = 64
    eng = az.Engine(n, S, 1, evaluator=evaluator, seed=321)
    eng.selfplay(S, plies=n * 64, temperature=T, recycle=False)
    st = eng.stats()
    assert st["overflow"] == 0 and st["games_finished"] == n
    states, pi, z = eng.drain_samples()
    got = {}
    i = 0
    while i < len(states):
        j = i + 1
        while j < len(states) and not (states[j][0].sum() == 2 and states[j][1].sum() == 2):
            j += 1
        got[states[i:j].tobytes()] = (pi[i:j], z[i:j])
        i = j
    total = 0
    for g in range(n):
        samples, win = orc.self_play_game(S, 1, evaluator=evaluator, seed=321, game_id=g, temperature=T)
        total += len(samples)
        key = np.array([orc.planes(s.black, s.white, s.side) for s in samples], dtype=np.float32).tobytes()
        assert key in got, g
        gpi, gz = got[key]
        assert np.array_equal(gpi, np.array([orc.action_probs(np.array(s.visits[:]), T).astype(np.float32) for s in samples]))
        assert np.array_equal(gz, np.array([s.z for s in samples], dtype=np.float32))
    assert len(states) == total == st["samples"]


def test_persistent_selfplay_budget_and_recycling(az):
    n, S = 128, 30
    eng = az.Engine(n, S, 1, evaluator=az.EVAL_ROLLOUT, seed=5)
    eng.selfplay(S, plies=n * 100, temperature=1.0, recycle=True)
    st = eng.stats()
    assert st["overflow"] == 0
    assert st["sims"] == n * 100 * S            # exactly the budgeted game-plies were searched
    assert st["games_finished"] >= n            # ~60 plies per game: every slot recycled at least once
    states, pi, z = eng.drain_samples()
    assert len(states) == st["samples"] and np.allclose(pi.sum(axis=1), 1.0, atol=1e-5)


@pytest.mark.parametrize("lpg,n", [(8, 5), (4, 13), (2, 37), (4, 1)])
def test_group_kernels_ragged_sizes_and_lane_counts(az, lpg, n):
    """the several-games-per-warp wave-1 kernels (rvs_treeg.cuh) with 8 / 4 / 2 lanes per game and game
    counts that leave groups of the last warp empty: persistent self-play samples == oracle"""
    try:
        S, T = 40, 1.0
        eng = az.Engine(n, S, 1, evaluator=az.EVAL_ROLLOUT, seed=77)
        eng.set_lanes_per_game(lpg)
        eng.selfplay(S, plies=n * 64, temperature=T, recycle=False)
        st = eng.stats()
        assert st["overflow"] == 0 and st["games_finished"] == n
        pk = eng.drain_packed()
        got = {}
        i = 0
        while i < len(pk):
            j = i + 1
            while j < len(pk) and not (int(pk.black[j]) == orc.START[0] and int(pk.white[j]) == orc.START[1]):
                j += 1
            got[(pk.black[i:j].tobytes(), pk.white[i:j].tobytes(), pk.side[i:j].tobytes())] = (pk.pi[i:j], pk.z[i:j])
            i = j
        for g in range(n):
            samples, win = orc.self_play_game(S, 1, evaluator=1, seed=77, game_id=g, temperature=T)
            key = (np.array([s.black for s in samples], dtype=np.uint64).tobytes(), np.array([s.white for s in samples], dtype=np.uint64).tobytes(),
                   np.array([s.side for s in samples], dtype=np.uint8).tobytes())
            assert key in got, g
            gpi, gz = got[key]
            assert np.array_equal(gpi, np.array([orc.action_probs(np.array(s.visits[:]), T).astype(np.float32) for s in samples]))
            assert np.array_equal(gz, np.array([s.z for s in samples], dtype=np.int8))
        eng.close()
    finally:
        pass


def test_more_games_than_resident_groups(az):
    """12 000 games x 8 lanes = 3000 warps > the 2368 resident one-warp CTAs: groups own several slots
    (search: one after the other; persistent self-play: round-robin).  Visit counts of a strided sample of
    games equal the oracle's, and a budgeted self-play launch spends exactly its budget."""
    try:
        n, S = 12000, 24
        eng = az.Engine(n, S, 1, evaluator=az.EVAL_ROLLOUT, seed=5)
        eng.set_lanes_per_game(8)
        eng.search(S, 1)
        v = eng.root_visits()
        st = eng.stats()
        assert st["overflow"] == 0 and st["sims"] == n * S
        for g in list(range(0, n, 397)) + [n - 1, 9471, 9472, 9473]:
            ov, *_ = orc.mcts_search(orc.START, S, 1, evaluator=1, seed=5, game_id=g, search_id=0)
            assert np.array_equal(v[g], ov), g
        s0 = eng.stats()
        eng.selfplay(S, plies=3 * n + 5, temperature=1.0, recycle=True)
        s1 = eng.stats()
        assert s1["sims"] - s0["sims"] == (3 * n + 5) * S and s1["overflow"] == 0 and s1["stalled"] == 0
        eng.close()
    finally:
        pass
