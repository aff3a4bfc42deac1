"""CPU tier: pins the C oracle's MCTS / action-prob / self-play restatement against the live
reference (tests/golden/mcts.npz, selfplay.npz from oracle/gen_golden.py).

The reference's own tests do not cover MCTS; these goldens are outputs of the
unmodified reference run with deterministic stub models (E0/T1/T2, see gen_golden.StubModel)."""
import numpy as np
import pytest

import orc

KINDS = ["E0", "T1", "T2"]


def _t2_table(g):
    return {(int(k[0]), int(k[1])): (g["t2_probs"][i], g["t2_values"][i])
            for i, k in enumerate(g["t2_keys"])}


def make_eval(kind, g):
    if kind == "E0":
        return orc.EVAL_E0, None
    if kind == "T1":
        sp = g["subset_prior"]
        return orc.EVAL_CALLBACK, (lambda pos: orc.t1_eval(pos, sp))
    tab = _t2_table(g)

    def t2(pos):
        p = np.zeros((len(pos), 65), dtype=np.float32)
        v = np.zeros(len(pos), dtype=np.float32)
        for i, (b, w, s) in enumerate(pos):
            p[i], v[i] = tab[orc.own_opp(b, w, s)]  # KeyError == search diverged from the reference
        return p, v
    return orc.EVAL_CALLBACK, t2


def test_uniform_prior_bits(golden):
    g = golden["mcts"]
    assert g["uniform_prior"].view(np.uint32)[0] == 0x3C7C0FC1
    assert np.float32(1.0) / np.float32(65.0) == g["uniform_prior"][0]
    # exactness assumption behind the T1 evaluator: torch softmax over k zero logits == f32(1/k)
    for k in range(1, 66):
        assert g["subset_prior"][k - 1] == np.float32(1.0) / np.float32(k)


def test_survey_goldens():
    """SURVEY.md 8(c) table, measured on the live reference"""
    def vis(pos, S, K):
        v, rn, rw, ne = orc.mcts_search(pos, S, K)
        assert rn == S
        return {i: int(x) for i, x in enumerate(v) if x or i in (19, 26, 37, 44)}
    assert vis(orc.START, 100, 64) == {19: 36, 26: 0, 37: 0, 44: 0}
    assert vis(orc.START, 100, 1) == {19: 13, 26: 78, 37: 4, 44: 4}
    assert vis(orc.START, 100, 8) == {19: 24, 26: 24, 37: 24, 44: 20}
    assert vis(orc.START, 800, 64) == {19: 192, 26: 192, 37: 192, 44: 160}
    p2 = (0x000000081C0A000E, 0x0000001000040211, 1)
    v, *_ = orc.mcts_search(p2, 100, 1)
    assert {i: int(x) for i, x in enumerate(v) if x} == {5: 6, 10: 71, 16: 4, 37: 6, 44: 6, 45: 6}
    v, *_ = orc.mcts_search(p2, 200, 16)
    assert {i: int(x) for i, x in enumerate(v) if x} == {5: 24, 10: 32, 16: 32, 37: 32, 44: 32, 45: 32}


def test_all_reference_cases(golden):
    g = golden["mcts"]
    bad = []
    for ci, (kind, pi, S, K) in enumerate(g["cases"]):
        pos = tuple(int(x) for x in g["pos"][pi])
        ev, fn = make_eval(KINDS[kind], g)
        c = float(g["root_w_c_n"][ci][1])
        v, rn, rw, ne = orc.mcts_search(pos, int(S), int(K), c_puct=c, evaluator=ev, pyfn=fn)
        ok = np.array_equal(v, g["visits"][ci]) and rn == int(g["root_w_c_n"][ci][2]) \
            and np.float32(rw).view(np.uint32) == g["root_w_c_n"][ci][0].view(np.uint32)
        if not ok:
            bad.append((ci, KINDS[kind], int(pi), int(S), int(K), c))
    assert not bad, f"{len(bad)} of {len(g['cases'])} cases differ: {bad[:8]}"


def test_action_probs(golden):
    g = golden["mcts"]
    v, *_ = orc.mcts_search(orc.START, 400, 64)
    pi = orc.action_probs(v, 1.0)
    assert np.array_equal(pi, g["ap_pi"][0])
    assert np.allclose(pi[[19, 26, 37, 44]], np.array([4, 4, 5, 8]) / 21.0)
    pi = orc.action_probs(v, 0.5)
    assert np.allclose(pi, g["ap_pi"][1], rtol=1e-15, atol=0)
    assert orc.action_probs(np.zeros(65, dtype=np.int32), 1.0).sum() == 0.0


@pytest.mark.parametrize("tag,kind", [("e0_t1", "E0"), ("t1_t1", "T1"), ("t1_t05", "T1")])
def test_selfplay_replay(golden, tag, kind):
    """Replays the reference SelfPlay game: at every recorded state the oracle's search must
    give the recorded pi (f64, bit exact at T=1), and the recorded z must follow the winner."""
    import ctypes as C
    gs, gm = golden["selfplay"], golden["mcts"]
    S, K, T, _ = gs[f"{tag}_cfg"]
    ev, fn = make_eval(kind, gm)
    states, pis, players, zs = (gs[f"{tag}_{k}"] for k in ("states", "pi", "players", "z"))
    L = orc.lib()
    b = orc.make_board(*orc.START)
    for p in range(len(states)):
        assert np.array_equal(orc.planes(b.black, b.white, b.side), states[p])
        assert b.side == players[p]
        v, *_ = orc.mcts_search((b.black, b.white, b.side), int(S), int(K), evaluator=ev, pyfn=fn)
        pi = orc.action_probs(v, float(T))
        if T == 1.0:
            assert np.array_equal(pi, pis[p])
        else:
            assert np.allclose(pi, pis[p], rtol=1e-14, atol=0)
        # the move the reference sampled = the one that leads to the next recorded state
        if p + 1 < len(states):
            nxt = None
            for mv in np.nonzero(v)[0]:
                c = orc.Board(b.black, b.white, b.side, b.over, b.winner, b.passes)
                L.orc_apply(C.byref(c), int(mv), 0)
                if np.array_equal(orc.planes(c.black, c.white, c.side), states[p + 1]) \
                        and c.side == players[p + 1]:
                    nxt = c
                    break
            assert nxt is not None
            b = nxt
        else:
            mv = int(np.nonzero(v)[0][0]) if np.count_nonzero(v) == 1 else None
            for mv in np.nonzero(v)[0]:
                c = orc.Board(b.black, b.white, b.side, b.over, b.winner, b.passes)
                L.orc_apply(C.byref(c), int(mv), 0)
                if c.over:
                    exp = [0.0 if c.winner == 0 else (1.0 if pl == c.winner else -1.0) for pl in players]
                    if np.array_equal(np.array(exp, dtype=np.float32), zs):
                        return
            raise AssertionError("no terminal move reproduces the recorded z")
