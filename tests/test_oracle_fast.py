"""CPU tier: the FAST search mode of the oracle (oracle/rvs_oracle.c: mcts_search_fast) against a second,
pure-Python restatement of the same specification written from the prose in the oracle's header comment
(small cases only), plus the properties the mode exists for.  The reference has no such mode -- PARITY
UNPINNED BY THE REFERENCE; the specification is this engine's (DESIGN.md "FAST mode")."""
import ctypes as C

import numpy as np
import pytest

import orc

f32 = np.float32


def _bf16(x):
    u = int(np.array([x], dtype=np.float32).view(np.uint32)[0])
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return np.array([u], dtype=np.uint32).view(np.float32)[0]


def py_fast_search(pos, S, K, c_puct=1.0, evaluator=None):
    """spec: first wave = 1 simulation; descent with VL; leaf VL; W from the perspective of the player who moved into
    the node; first max wins; priors bf16-rounded.  evaluator(board) -> (probs[65], value for the side to move)"""
    L = orc.lib()
    nodes = [dict(N=0, W=f32(0), VL=0, P=f32(1), kids=None, move=255, term=None)]

    def board_copy(b):
        return orc.Board(b.black, b.white, b.side, b.over, b.winner, b.passes)

    root = orc.make_board(*pos)
    done = 0
    evals = unique = 0
    while done < S:
        k = 1 if done == 0 else min(K, S - done)
        done += k
        leaves = []
        for _ in range(k):
            b = board_copy(root)
            path, sides, node = [0], [b.side], 0
            while nodes[node]["kids"] and nodes[node]["term"] is None:
                nd = nodes[node]
                nd["VL"] += 1
                sq = f32(np.sqrt(np.float64(nd["N"] + nd["VL"])))
                best, nxt = f32(-np.inf), None
                for ci in nd["kids"]:
                    ch = nodes[ci]
                    n = ch["N"] + ch["VL"]
                    q = f32(f32(ch["W"] - f32(ch["VL"])) / f32(n)) if n > 0 else f32(0)
                    u = f32(f32(f32(f32(c_puct) * ch["P"]) * sq) / f32(1 + n))
                    sc = f32(q + u)
                    if sc > best:
                        best, nxt = sc, ci
                L.orc_apply(C.byref(b), nodes[nxt]["move"], 0)
                node = nxt
                path.append(node)
                sides.append(b.side)
            nodes[node]["VL"] += 1

            def backup(vb, path=path, sides=sides):
                for i in range(len(path) - 1, -1, -1):
                    x = nodes[path[i]]
                    mover = sides[0] if i == 0 else sides[i - 1]
                    if x["VL"] > 0:
                        x["VL"] -= 1
                    x["N"] += 1
                    x["W"] = f32(x["W"] + (vb if mover == 1 else f32(-vb)))
            if nodes[node]["term"] is not None:
                backup(nodes[node]["term"])
                continue
            leaves.append((node, b, backup))
        pending, seen = [], {}
        for node, b, backup in leaves:
            lm = L.orc_board_legal(C.byref(b), 0)
            if lm == 0:
                tv = f32(0) if not b.over else f32(1) if b.winner == 1 else f32(-1) if b.winner == 2 else f32(0)
                nodes[node]["term"] = tv
                backup(tv)
                continue
            pending.append((node, b, backup, lm))
        for node, b, backup, lm in pending:
            evals += 1
            if node not in seen:
                unique += 1
                seen[node] = evaluator(b)
            probs, v = seen[node]
            if not nodes[node]["kids"]:
                kids = []
                for sq in range(64):
                    if (lm >> sq) & 1:
                        nodes.append(dict(N=0, W=f32(0), VL=0, P=_bf16(probs[sq]), kids=None, move=sq, term=None))
                        kids.append(len(nodes) - 1)
                nodes[node]["kids"] = kids
            backup(f32(v) if b.side == 1 else f32(-f32(v)))
    vis = np.zeros(65, dtype=np.int32)
    for ci in nodes[0]["kids"] or []:
        vis[nodes[ci]["move"]] = nodes[ci]["N"]
    return vis, nodes[0]["N"], nodes[0]["W"], evals, unique


def _e0(b):
    own = bin(b.black if b.side == 1 else b.white).count("1")
    opp = bin(b.white if b.side == 1 else b.black).count("1")
    return np.full(65, f32(1) / f32(65), dtype=np.float32), f32(own - opp) / f32(64)


def _hashed(b):
    """T1-style evaluator: hashed value, non-uniform priors (exercises the bf16 rounding and the first-max rule)"""
    own, opp = orc.own_opp(b.black, b.white, b.side)
    h = orc.mix64((own * 0x9E3779B97F4A7C15) ^ orc.mix64(opp))
    v = f32(((h >> 20) & 0xFFFF) - 32768) / f32(32768)
    raw = np.array([((orc.mix64(h + i) >> 40) & 0xFFFF) + 1 for i in range(65)], dtype=np.float64)
    return (raw / raw.sum()).astype(np.float32), v


ROOTS = [orc.START, (0x000000081C0A000E, 0x0000001000040211, 1)]


@pytest.mark.parametrize("S,K,c", [(60, 1, 1.0), (80, 4, 1.0), (120, 8, 1.5), (150, 16, 0.7), (130, 64, 1.0)])
@pytest.mark.parametrize("kind", ["e0", "hashed"])
def test_c_oracle_fast_equals_python_restatement(S, K, c, kind):
    ev = _e0 if kind == "e0" else _hashed

    def pyfn(pos):
        ps, vs = [], []
        for (bl, wh, sd) in pos:
            p, v = ev(orc.Board(bl, wh, sd, 0, 0, 0))
            ps.append(p); vs.append(v)
        return np.array(ps, dtype=np.float32), np.array(vs, dtype=np.float32)
    for root in ROOTS:
        exp = py_fast_search(root, S, K, c, ev)
        got = orc.mcts_search_fast(root, S, K, c_puct=c, evaluator=orc.EVAL_CALLBACK, pyfn=pyfn)
        assert np.array_equal(got[0], exp[0]), (root, got[0][got[0] > 0], exp[0][exp[0] > 0])
        assert got[1] == exp[1] and np.float32(got[2]) == exp[2] and got[3] == exp[3] and got[4] == exp[4]
        if kind == "e0":  # the built-in E0 evaluator is the same function
            assert np.array_equal(orc.mcts_search_fast(root, S, K, c_puct=c, evaluator=orc.EVAL_E0)[0], exp[0])


def test_fast_mode_properties():
    for K in (1, 8, 16, 64):
        v, rn, rw, ne, nu = orc.mcts_search_fast(orc.START, 800, K, evaluator=orc.EVAL_ROLLOUT, seed=3)
        assert v.sum() == 799 and rn == 800 and 780 <= ne <= 800  # a few simulations end on terminal nodes
        # the mode's purpose: a wave lands on (mostly) distinct leaves -- the reference-compatible wave of 64 evaluates
        # ~1 unique leaf per wave, i.e. ~13 unique leaves in 800 simulations
        assert nu >= (760 if K <= 16 else 600), (K, nu)
    # set_search_mode(1) routes the ordinary entry points through the same code
    orc.set_search_mode(1)
    try:
        a = orc.mcts_search(orc.START, 200, 16, evaluator=orc.EVAL_E0)[0]
    finally:
        orc.set_search_mode(0)
    assert np.array_equal(a, orc.mcts_search_fast(orc.START, 200, 16, evaluator=orc.EVAL_E0)[0])
    assert not np.array_equal(a, orc.mcts_search(orc.START, 200, 16, evaluator=orc.EVAL_E0)[0])
