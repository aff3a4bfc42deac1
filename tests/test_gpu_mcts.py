"""GPU tier (K2): the CUDA tree kernels against (a) visit counts recorded from the live reference
(tests/golden/mcts.npz) and (b) the C oracle on seeded inputs.  Visit counts, root N and root W
bits must match exactly (BASELINE.json: bit-exact under a fixed deterministic evaluator)."""
import ctypes as C

import numpy as np
import pytest

import orc
from test_oracle_mcts import KINDS, _t2_table

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def az():
    import alphazero_reversi_b200 as m
    return m


def _positions(g, idxs):
    pos = g["pos"][idxs]
    return (np.ascontiguousarray(pos[:, 0]), np.ascontiguousarray(pos[:, 1]),
            np.ascontiguousarray(pos[:, 2]).astype(np.uint8))


def test_survey_goldens_fused_e0(az):
    """measured on the live reference, start position"""
    for S, K, exp in ((100, 64, {19: 36, 26: 0, 37: 0, 44: 0}), (100, 1, {19: 13, 26: 78, 37: 4, 44: 4}),
                      (100, 8, {19: 24, 26: 24, 37: 24, 44: 20}), (800, 64, {19: 192, 26: 192, 37: 192, 44: 160})):
        eng = az.Engine(1, S, K, evaluator=az.EVAL_E0)
        eng.search(S, K)
        v = eng.root_visits()[0]
        assert {i: int(v[i]) for i in (19, 26, 37, 44)} == exp and v.sum() == sum(exp.values())
        assert eng.stats()["overflow"] == 0
        eng.close()


def test_reference_cases_e0_batched(az, golden):
    """every E0 case of the reference golden set; cases sharing (S,K,c) run as ONE batched search"""
    g = golden["mcts"]
    groups = {}
    for ci, (kind, pi, S, K) in enumerate(g["cases"]):
        if KINDS[kind] == "E0":
            groups.setdefault((int(S), int(K), float(g["root_w_c_n"][ci][1])), []).append((ci, int(pi)))
    bad = []
    for (S, K, c), items in groups.items():
        eng = az.Engine(len(items), S, K, evaluator=az.EVAL_E0, c_puct=c)
        eng.set_positions(*_positions(g, [pi for _, pi in items]))
        eng.search(S, K)
        v = eng.root_visits()
        assert eng.stats()["overflow"] == 0
        for row, (ci, pi) in enumerate(items):
            if not np.array_equal(v[row], g["visits"][ci]):
                bad.append((ci, pi, S, K))
        eng.close()
    assert not bad, bad


def _run_external(az, eng, S, K, fn):
    """MCTS.search split at the evaluator, host-side evaluator fn(list of (b,w,side)) -> probs, values"""
    eng.begin_search()
    for start in range(0, S, K):
        k = min(K, S - start)
        eng.select(k)
        planes, valid = eng.leaf_planes()
        n = len(valid)
        probs = np.zeros((n, 65), dtype=np.float32)
        values = np.zeros(n, dtype=np.float32)
        idx = np.nonzero(valid)[0]
        if len(idx):
            w = (1 << np.arange(64, dtype=np.uint64))
            pos = []
            for i in idx:
                own = int(((planes[i, 0].reshape(-1) > 0.5).astype(np.uint64) * w).sum())
                opp = int(((planes[i, 1].reshape(-1) > 0.5).astype(np.uint64) * w).sum())
                pos.append((own, opp, 1))  # evaluators only see (own, opp)
            p, v = fn(pos)
            probs[idx] = p
            values[idx] = v
        eng.process(probs, values)


@pytest.mark.parametrize("kind", ["T1", "T2", "E0"])
def test_reference_cases_external_path(az, golden, kind):
    """T1/T2 (position-dependent priors and values) and E0 again through select / leaf_planes /
    process -- the path MCTS(model) uses for arbitrary torch models"""
    g = golden["mcts"]
    sp = g["subset_prior"]
    tab = _t2_table(g)

    def ev(pos):
        if kind == "T1":
            return orc.t1_eval(pos, sp)
        p = np.zeros((len(pos), 65), dtype=np.float32)
        v = np.zeros(len(pos), dtype=np.float32)
        for i, (own, opp, _) in enumerate(pos):
            if kind == "T2":
                p[i], v[i] = tab[(own, opp)]
            else:
                p[i, :] = np.float32(1.0) / np.float32(65.0)
                v[i] = np.float32(bin(own).count("1") - bin(opp).count("1")) / np.float32(64)
        return p, v

    groups = {}
    for ci, (kd, pi, S, K) in enumerate(g["cases"]):
        if KINDS[kd] == kind:
            groups.setdefault((int(S), int(K), float(g["root_w_c_n"][ci][1])), []).append((ci, int(pi)))
    if kind == "E0":  # E0 is already covered by the fused kernel; keep a light sample here
        groups = dict(list(groups.items())[:3])
    bad = []
    for (S, K, c), items in groups.items():
        eng = az.Engine(len(items), S, K, evaluator=az.EVAL_EXTERNAL, c_puct=c)
        eng.set_positions(*_positions(g, [pi for _, pi in items]))
        _run_external(az, eng, S, K, ev)
        v = eng.root_visits()
        assert eng.stats()["overflow"] == 0
        for row, (ci, pi) in enumerate(items):
            if not np.array_equal(v[row], g["visits"][ci]):
                bad.append((ci, pi, S, K, c))
        eng.close()
    assert not bad, bad


def _random_roots(n, seed):
    """reachable positions: oracle random playouts cut at a random ply"""
    rng = np.random.default_rng(seed)
    L = orc.lib()
    bl, wh, sd = [], [], []
    for g in range(n):
        b = orc.make_board(*orc.START)
        st = L.orc_stream_seed(seed, g, 1)
        cut = int(rng.integers(0, 58))
        for _ in range(cut):
            if b.over:
                break
            lm = L.orc_board_legal(C.byref(b), 0)
            moves = [i for i in range(64) if (lm >> i) & 1]
            st = orc.mix64(st + 1)
            L.orc_apply(C.byref(b), moves[st % len(moves)], 0)
        bl.append(b.black); wh.append(b.white); sd.append(b.side)
    return np.array(bl, dtype=np.uint64), np.array(wh, dtype=np.uint64), np.array(sd, dtype=np.uint8)


@pytest.mark.parametrize("evaluator,S,K", [(0, 100, 1), (0, 100, 64), (0, 240, 16), (1, 100, 1), (1, 100, 64), (1, 150, 8)])
def test_batched_search_vs_oracle(az, evaluator, S, K):
    """512 concurrent games from random reachable roots (incl. finished ones); E0 and ROLLOUT
    evaluators; every game's visit vector equals the oracle's"""
    n = 512
    bl, wh, sd = _random_roots(n, 100 + S + K)
    eng = az.Engine(n, S, K, evaluator=evaluator, seed=4242)
    eng.set_positions(bl, wh, sd)
    eng.search(S, K)
    v = eng.root_visits()
    st = eng.stats()
    assert st["overflow"] == 0 and st["sims"] == n * S
    tot_evals = 0
    for g in range(n):
        ov, rn, rw, ne = orc.mcts_search((int(bl[g]), int(wh[g]), int(sd[g])), S, K, evaluator=evaluator,
                                         seed=4242, game_id=g, search_id=0)
        assert np.array_equal(v[g], ov), (g, hex(int(bl[g])), hex(int(wh[g])), int(sd[g]))
        tot_evals += ne
    assert st["evals"] == tot_evals


@pytest.mark.parametrize("evaluator,S,K,alpha,eps,lpg", [(0, 100, 1, 0.03, 0.25, 8), (0, 100, 1, 0.3, 0.25, 4), (1, 100, 1, 0.03, 0.25, 2),
                                                         (0, 128, 64, 0.03, 0.25, 0), (1, 150, 8, 1.0, 0.5, 0), (0, 200, 16, 2.5, 1.0, 0)])
def test_root_dirichlet_noise_vs_oracle(az, evaluator, S, K, alpha, eps, lpg):
    """engine feature (BASELINE config 4): Dirichlet noise mixed into the root priors after the root
    expansion; the oracle restates the sampler, visit counts must stay bit exact.  Covers the
    several-games-per-warp wave-1 kernels (lanes per game forced through rvs_engine_set_lanes_per_game), the fused wave
    kernel, and -- with an E0 evaluator driven from the host -- the external select/process path."""
    n = 192
    bl, wh, sd = _random_roots(n, 500 + S + K)
    orc.set_root_noise(alpha, eps)
    try:
        eng = az.Engine(n, S, K, evaluator=evaluator, seed=777)
        if lpg:
            eng.set_lanes_per_game(lpg)
        eng.set_root_noise(alpha, eps)
        eng.set_positions(bl, wh, sd)
        eng.search(S, K)
        v = eng.root_visits()
        assert eng.stats()["overflow"] == 0
        changed = 0
        for g in range(n):
            pos = (int(bl[g]), int(wh[g]), int(sd[g]))
            ov, *_ = orc.mcts_search(pos, S, K, evaluator=evaluator, seed=777, game_id=g, search_id=0)
            assert np.array_equal(v[g], ov), (g, hex(pos[0]), hex(pos[1]), pos[2])
        eng.close()
        if evaluator == 0:  # same searches through select / leaf_planes / process (the path NN evaluators use)
            m = 48
            ext = az.Engine(m, S, K, evaluator=az.EVAL_EXTERNAL, seed=777)
            ext.set_root_noise(alpha, eps)
            ext.set_positions(bl[:m], wh[:m], sd[:m])

            def ev(pos):
                p = np.full((len(pos), 65), np.float32(1.0) / np.float32(65.0), dtype=np.float32)
                val = np.array([np.float32(bin(o).count("1") - bin(q).count("1")) / np.float32(64) for o, q, _ in pos], dtype=np.float32)
                return p, val
            _run_external(az, ext, S, K, ev)
            assert np.array_equal(ext.root_visits(), v[:m])
            ext.close()
        # the noise really changes searches (otherwise this test proves nothing)
        orc.set_root_noise(0.0, 0.0)
        for g in range(n):
            ov0, *_ = orc.mcts_search((int(bl[g]), int(wh[g]), int(sd[g])), S, K, evaluator=evaluator, seed=777, game_id=g)
            changed += int(not np.array_equal(ov0, v[g]))
        if K <= 16:  # with the reference's wave of 64 the first waves all follow +inf scores: priors cannot matter yet
            assert changed > n // 4, changed
    finally:
        orc.set_root_noise(0.0, 0.0)


@pytest.mark.parametrize("rules,c_puct,lpg,evaluator", [(1, 0.7, 8, 1), (1, 2.5, 4, 0), (1, 1.0, 2, 1), (0, 0.7, 4, 1), (0, 2.5, 2, 0), (0, 1.5, 8, 1)])
def test_wave1_group_kernels_rules_and_cpuct_vs_oracle(az, rules, c_puct, lpg, evaluator):
    """the several-games-per-warp wave-1 search kernels under both rule sets, several exploration
    constants and every lanes-per-game setting: visit counts equal the oracle's"""
    n, S = 96, 120
    bl, wh, sd = _random_roots(n, 31 + lpg + rules)
    eng = az.Engine(n, S, 1, evaluator=evaluator, rules=rules, c_puct=c_puct, seed=2024)
    eng.set_lanes_per_game(lpg)
    eng.set_positions(bl, wh, sd)
    eng.search(S, 1)
    v = eng.root_visits()
    st = eng.stats()
    assert st["overflow"] == 0 and st["sims"] == n * S
    for g in range(n):
        ov, *_ = orc.mcts_search((int(bl[g]), int(wh[g]), int(sd[g])), S, 1, c_puct=c_puct, rules=rules, evaluator=evaluator,
                                 seed=2024, game_id=g)
        assert np.array_equal(v[g], ov), (g, hex(int(bl[g])), hex(int(wh[g])), int(sd[g]))
    eng.close()


def test_strict_rules_search_vs_oracle(az):
    n, S, K = 64, 120, 4
    bl, wh, sd = _random_roots(n, 9)
    eng = az.Engine(n, S, K, evaluator=az.EVAL_ROLLOUT, rules=az.RULES_STRICT, seed=1)
    # random REF-rule roots are still valid disc sets; STRICT legality is what is searched
    eng.set_positions(bl, wh, sd)
    eng.search(S, K)
    v = eng.root_visits()
    for g in range(n):
        ov, *_ = orc.mcts_search((int(bl[g]), int(wh[g]), int(sd[g])), S, K, rules=1, evaluator=1, seed=1, game_id=g)
        assert np.array_equal(v[g], ov)


def test_mcts_mirror_class(az, golden):
    """the drop-in MCTS class: dict result, get_action_probs pi (f64) and the numpy RNG draw"""
    from stubs import StubModel
    g = golden["mcts"]
    game = az.ReversiGame()
    m = az.MCTS(StubModel("E0"), c_puct=1.0, num_simulations=400, batch_size=64)
    counts = m.search(game)
    assert counts == {(2, 3): 64, (3, 2): 64, (4, 5): 80, (5, 4): 128}
    np.random.seed(7)
    a, pi = m.get_action_probs(game, temperature=1.0)
    assert np.array_equal(pi, g["ap_pi"][0]) and pi.dtype == np.float64
    np.random.seed(7)
    exp = np.random.choice(65, p=g["ap_pi"][0])
    assert a == (exp // 8, exp % 8)
    assert (game.board.black, game.board.white) == orc.START[:2]  # caller's game untouched
    # built-in evaluator object instead of a model
    m2 = az.MCTS(az.UniformDiscDiff(), num_simulations=100, batch_size=1)
    assert m2.search(game) == {(2, 3): 13, (3, 2): 78, (4, 5): 4, (5, 4): 4}
    # torch model living on the GPU: leaves stay on the device
    m3 = az.MCTS(StubModel("T1", device="cuda:0"), num_simulations=100, batch_size=8)
    c3 = m3.search(game)
    ci = [i for i, (kd, pi_, S, K) in enumerate(g["cases"]) if KINDS[kd] == "T1" and pi_ == 0 and S == 100 and K == 8][0]
    assert {k: v for k, v in c3.items()} == {divmod(i, 8): int(n) for i, n in enumerate(g["visits"][ci][:64])
                                              if (i // 8, i % 8) in c3}


def test_full_size_search_properties(az):
    """BASELINE configs[1] at full size: 4096 concurrent games x 100 simulations (wave 1, rollout evaluator)
    from random reachable roots.  Size-independent properties for every game (root N = 100, child visits =
    sims that left the root, visits only on legal squares, counters consistent) and the oracle on a strided
    sample of games."""
    n, S = 4096, 100
    bl, wh, sd = _random_roots(n, 4242)
    eng = az.Engine(n, S, 1, evaluator=az.EVAL_ROLLOUT, seed=31337)
    eng.set_positions(bl, wh, sd)
    eng.search(S, 1)
    v = eng.root_visits()
    st = eng.stats()
    assert st["overflow"] == 0 and st["sims"] == n * S and st["stalled"] == 0
    lm = az.board_ops.legal_masks(bl, wh, sd)
    total = v.sum(axis=1)
    for g in range(n):
        legal = [(int(lm[g]) >> q) & 1 for q in range(64)]
        assert all(v[g][q] == 0 for q in range(64) if not legal[q]) and v[g][64] == 0
        if int(lm[g]) == 0:
            assert total[g] == 0              # finished game / no move: every simulation stops at the root
        else:
            assert total[g] == S - 1          # the first simulation expands the root, the other 99 descend
    assert st["evals"] <= n * S and st["nodes"] > 0 and st["tree_bytes"] > 0
    for g in range(0, n, 127):
        ov, *_ = orc.mcts_search((int(bl[g]), int(wh[g]), int(sd[g])), S, 1, evaluator=1, seed=31337, game_id=g)
        assert np.array_equal(v[g], ov), g
    eng.close()


@pytest.mark.parametrize("lpg,evaluator", [(8, 1), (4, 0), (2, 1)])
def test_wave1_800_simulations_vs_oracle(az, lpg, evaluator):
    """config-4 depth of search (800 simulations per move, wave 1): deep trees, multi-chunk child scans,
    node pools of 2 + 34 * 800 rows"""
    n, S = 24, 800
    bl, wh, sd = _random_roots(n, 808 + lpg)
    bl[0], wh[0], sd[0] = orc.START
    eng = az.Engine(n, S, 1, evaluator=evaluator, seed=55)
    eng.set_lanes_per_game(lpg)
    eng.set_positions(bl, wh, sd)
    eng.search(S, 1)
    v = eng.root_visits()
    st = eng.stats()
    assert st["overflow"] == 0 and st["sims"] == n * S
    for g in range(n):
        ov, *_ = orc.mcts_search((int(bl[g]), int(wh[g]), int(sd[g])), S, 1, evaluator=evaluator, seed=55, game_id=g)
        assert np.array_equal(v[g], ov), g
    eng.close()


@pytest.mark.parametrize("lpg", [8, 4, 2])
def test_small_node_pool_overflows_cleanly(az, lpg):
    """A node pool that is too small for the search (nodes_per_game = 48 rows for 100 simulations): the wave-1 group
    kernels create child rows lazily, when a traverse first descends through an expanded node, so the pool runs out in
    the middle of a descent.  The search must flag the overflow, stay inside the pool (the neighbouring games' trees
    are compared with the oracle) and still back every simulation up to the root."""
    n, S = 64, 100
    bl, wh, sd = _random_roots(n, 999)
    small = az.Engine(n, S, 1, evaluator=az.EVAL_ROLLOUT, seed=5, nodes_per_game=48)
    small.set_lanes_per_game(lpg)
    small.set_positions(bl, wh, sd)
    small.search(S, 1)
    v = small.root_visits()
    st = small.stats()
    assert st["overflow"] > 0 and st["sims"] == n * S
    lm = az.board_ops.legal_masks(bl, wh, sd)
    for g in range(n):
        assert int(v[g].sum()) == (S - 1 if int(lm[g]) else 0), g
    small.close()
    # the same engine size with the default (worst-case) pool: no overflow and exact visit counts
    eng = az.Engine(n, S, 1, evaluator=az.EVAL_ROLLOUT, seed=5)
    eng.set_lanes_per_game(lpg)
    eng.set_positions(bl, wh, sd)
    eng.search(S, 1)
    v = eng.root_visits()
    assert eng.stats()["overflow"] == 0
    for g in range(0, n, 7):
        ov, *_ = orc.mcts_search((int(bl[g]), int(wh[g]), int(sd[g])), S, 1, evaluator=1, seed=5, game_id=g)
        assert np.array_equal(v[g], ov), g
    eng.close()
