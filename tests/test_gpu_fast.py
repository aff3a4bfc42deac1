"""GPU tier: RVS_MODE_FAST (virtual-loss PUCT with effective leaf batching) against its specification, the
oracle's independent restatement (oracle/rvs_oracle.c: mcts_search_fast).  The reference has no such mode --
its waves send every simulation down one path (src/mcts/mcts.py:96-100,113,355-392) -- so parity here is
"CUDA == oracle, bit for bit" under deterministic evaluators, plus the properties that make the mode worth
having (a wave spreads over distinct leaves)."""
import numpy as np
import pytest
import torch

import orc
from stubs import perturb_bn
from test_gpu_hardening import _legal_roots

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def az():
    import alphazero_reversi_b200 as m
    return m


@pytest.mark.parametrize("evaluator,S,K,c", [(0, 100, 1, 1.0), (0, 100, 8, 1.0), (0, 200, 16, 1.5), (0, 400, 64, 1.0), (0, 97, 64, 2.5),
                                             (1, 100, 8, 1.0), (1, 300, 16, 0.7), (1, 800, 64, 1.0), (1, 64, 1, 1.0)])
def test_fast_search_vs_oracle(az, evaluator, S, K, c):
    n = 96
    bl, wh, sd = _legal_roots(n, seed=100 + S + K)
    eng = az.Engine(n, S, K, evaluator=evaluator, c_puct=c, seed=4242)
    eng.set_search_mode(az.MODE_FAST)
    eng.set_positions(bl, wh, sd)
    eng.search(S, K)
    v = eng.root_visits()
    st = eng.stats()
    assert st["overflow"] == 0 and st["sims"] == n * S
    uniq = 0
    for g in range(n):
        pos = (int(bl[g]), int(wh[g]), int(sd[g]))
        ov, rn, rw, ne, nu = orc.mcts_search_fast(pos, S, K, c_puct=c, evaluator=evaluator, seed=4242, game_id=g)
        assert np.array_equal(v[g], ov), (g, hex(pos[0]), hex(pos[1]), pos[2], v[g][v[g] > 0], ov[ov > 0])
        uniq += nu
    # every simulation but the root's expansion lands in a root child
    live = np.array([orc.legal(*orc.own_opp(int(bl[g]), int(wh[g]), int(sd[g]))) != 0 for g in range(n)])
    assert (v.sum(axis=1)[live] == S - 1).all()


def test_fast_mode_through_external_path_and_mirror(az):
    """select / leaf_planes / process (the path external models and the NN use) in FAST mode == the fused kernel"""
    n, S, K = 40, 120, 16
    bl, wh, sd = _legal_roots(n, seed=7)
    eng = az.Engine(n, S, K, evaluator=az.EVAL_E0, seed=1)
    eng.set_search_mode(az.MODE_FAST)
    eng.set_positions(bl, wh, sd)
    eng.search(S, K)
    v = eng.root_visits()
    eng.close()
    ext = az.Engine(n, S, K, evaluator=az.EVAL_EXTERNAL, seed=1)
    ext.set_search_mode(az.MODE_FAST)
    ext.set_positions(bl, wh, sd)
    ext.begin_search()
    w = (1 << np.arange(64, dtype=np.uint64))
    start = 0
    while start < S:
        k = 1 if start == 0 else min(K, S - start)
        start += k
        ext.select(k)
        planes, valid = ext.leaf_planes()
        m = len(valid)
        own = ((planes[:, 0].reshape(m, 64) > 0.5).sum(axis=1)).astype(np.float32)
        opp = ((planes[:, 1].reshape(m, 64) > 0.5).sum(axis=1)).astype(np.float32)
        probs = np.full((m, 65), np.float32(1.0) / np.float32(65.0), dtype=np.float32)
        ext.process(probs, ((own - opp) / np.float32(64)).astype(np.float32))
    assert np.array_equal(ext.root_visits(), v)
    ext.close()
    # the reference-API mirror: MCTS(..., search_mode=MODE_FAST) returns the oracle's FAST visit counts
    m = az.MCTS(az.UniformDiscDiff(), num_simulations=100, batch_size=8, search_mode=az.MODE_FAST)
    counts = m.search(az.ReversiGame())
    ov, *_ = orc.mcts_search_fast(orc.START, 100, 8)
    assert counts == {divmod(i, 8): int(x) for i, x in enumerate(ov[:64]) if (orc.legal(orc.START[0], orc.START[1]) >> i) & 1}


def test_fast_mode_spreads_a_wave_over_distinct_leaves(az):
    """the point of the mode: with the built-in network, a wave of K simulations costs close to K unique
    evaluations worth of search (REF mode: about one), and NN self-play in FAST mode produces whole games"""
    torch.manual_seed(42)
    net = az.AlphaZeroNetwork(8, 2, 64)
    with torch.no_grad():
        perturb_bn(net, 43)
    rn = az.RvsNetwork.from_module(net.eval())
    n, S = 64, 400
    bl, wh, sd = _legal_roots(n, seed=11)
    ratios = {}
    for mode in (az.MODE_REF, az.MODE_FAST):
        for K in (8, 16, 64):
            eng = az.Engine(n, S, K, evaluator=az.EVAL_NN, net_blocks=2, net_filters=64)
            eng.set_search_mode(mode)
            rn.attach(eng)
            eng.set_positions(bl, wh, sd)
            eng.search(S, K)
            st = eng.stats()
            assert st["overflow"] == 0 and st["sims"] == n * S
            ratios[(mode, K)] = st["nn_evals"] / st["sims"]
            eng.close()
    print("unique network evaluations per simulation:", {k: round(v, 3) for k, v in ratios.items()})
    for K in (8, 16, 64):
        assert ratios[(az.MODE_FAST, K)] > 2.0 * ratios[(az.MODE_REF, K)]
    assert ratios[(az.MODE_FAST, 8)] > 0.85 and ratios[(az.MODE_FAST, 64)] > 0.5
    sp = az.SelfPlay(rn, {"num_simulations": 48, "batch_size": 16, "temperature": 1.0, "num_parallel_games": 16,
                          "search_mode": az.MODE_FAST})
    data = sp.generate_training_data(16)
    assert data["states"].shape[0] > 16 * 40 and np.allclose(data["action_probs"].sum(axis=1), 1.0, atol=1e-5)


def test_fast_nn_search_fused_equals_external(az):
    """FAST mode on the NN path: fused (compaction, de-duplication, remap) == external path fed the engine's own probs"""
    torch.manual_seed(42)
    net = az.AlphaZeroNetwork(8, 2, 64)
    with torch.no_grad():
        perturb_bn(net, 43)
    rn = az.RvsNetwork.from_module(net.eval())
    n, S, K = 32, 150, 16
    bl, wh, sd = _legal_roots(n, seed=23)
    eng = az.Engine(n, S, K, evaluator=az.EVAL_NN, net_blocks=2, net_filters=64)
    eng.set_search_mode(az.MODE_FAST)
    rn.attach(eng)
    eng.set_positions(bl, wh, sd)
    eng.search(S, K)
    v = eng.root_visits()
    ext = az.Engine(n, S, K, evaluator=az.EVAL_EXTERNAL)
    ext.set_search_mode(az.MODE_FAST)
    ext.set_positions(bl, wh, sd)
    ext.begin_search()
    w = (1 << np.arange(64, dtype=np.uint64))
    start = 0
    while start < S:
        k = 1 if start == 0 else min(K, S - start)
        start += k
        ext.select(k)
        planes, valid = ext.leaf_planes()
        m = len(valid)
        own = ((planes[:, 0].reshape(m, 64) > 0.5).astype(np.uint64) * w).sum(axis=1).astype(np.uint64)
        opp = ((planes[:, 1].reshape(m, 64) > 0.5).astype(np.uint64) * w).sum(axis=1).astype(np.uint64)
        probs, val = eng.predict(own, opp, np.ones(m, dtype=np.uint8), probs=True)
        probs[valid == 0] = 0
        val[valid == 0] = 0
        ext.process(probs, val)
    assert np.array_equal(ext.root_visits(), v)
    eng.close(); ext.close()
