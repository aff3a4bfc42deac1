"""CPU tier: pins the C oracle (oracle/rvs_oracle.c) against golden vectors produced by the live
reference (oracle/gen_golden.py) and against the reference's own tests (test_game.py:7-126)."""
import ctypes as C
import hashlib

import numpy as np

import orc


def test_reference_owned_pins():
    # test_game.py:7-39 initial layout and initial legal set
    b = orc.make_board(*orc.START)
    assert b.black == (1 << (3 * 8 + 4)) | (1 << (4 * 8 + 3))
    assert b.white == (1 << (3 * 8 + 3)) | (1 << (4 * 8 + 4))
    lm = orc.lib().orc_board_legal(C.byref(b), 0)
    assert lm == sum(1 << (r * 8 + c) for r, c in [(2, 3), (3, 2), (4, 5), (5, 4)])
    # test_game.py:42-57 make_move(2,3) places black, flips (3,3), white to move
    assert orc.lib().orc_apply(C.byref(b), 2 * 8 + 3, 0) == 1
    assert (b.black >> (2 * 8 + 3)) & 1 and (b.black >> (3 * 8 + 3)) & 1 and b.side == 2


def test_endgame_fill(golden):
    g = golden["board"]
    blk, wht = (int(x) for x in g["endgame_in"])
    b = orc.Board(blk, wht, 2, 0, 0, 0)
    ok = orc.lib().orc_apply(C.byref(b), 0, 0)
    eb, ew, eok, eover, ewin = (int(x) for x in g["endgame_out"])
    assert (b.black, b.white, ok, b.over, b.winner) == (eb, ew, eok, eover, ewin)
    assert ewin == 2  # test_game.py:117-124 expects White


def test_perft_ref(golden):
    exp = [int(x) for x in golden["board"]["perft"]]
    assert exp == [4, 12, 56, 244, 1396, 8200, 55134, 391210][:len(exp)]
    for d, e in enumerate(exp, start=1):
        assert orc.perft(d, orc.RULES_REF) == e


def test_perft_strict():
    # true Othello counts
    exp = [4, 12, 56, 244, 1396, 8200, 55092, 390216]
    for d, e in enumerate(exp, start=1):
        assert orc.perft(d, orc.RULES_STRICT) == e


def test_first_last_games(golden):
    g = golden["board"]
    for name in ("first", "last"):
        b = orc.make_board(*orc.START)
        for mv in g[f"{name}_moves"]:
            assert orc.lib().orc_apply(C.byref(b), int(mv), 0) == 1
        assert (b.black, b.white) == tuple(int(x) for x in g[f"{name}_final"])
        assert b.over == 1 and b.winner == int(g[f"{name}_winner"][0])
    assert int(g["first_final"][0]) == 0x012B55E5F4F2F2FE


def test_seeded_games(golden):
    g = golden["board"]
    h = hashlib.sha256()
    L = orc.lib()
    for gi in range(g["seed_moves"].shape[0]):
        b = orc.make_board(*orc.START)
        n = int(g["seed_nply"][gi])
        for p in range(n + 1):
            assert (b.black, b.white, b.side) == (int(g["seed_state"][gi, p, 0]),
                                                  int(g["seed_state"][gi, p, 1]),
                                                  int(g["seed_side"][gi, p]))
            assert L.orc_board_legal(C.byref(b), 0) == int(g["seed_legal"][gi, p])
            assert b.over == int(g["seed_over"][gi, p])
            if p == n:
                break
            assert L.orc_apply(C.byref(b), int(g["seed_moves"][gi, p]), 0) == 1
            h.update(b.black.to_bytes(8, "little") + b.white.to_bytes(8, "little") + bytes([b.side]))
        assert b.winner == int(g["seed_winner"][gi])
    assert h.hexdigest() == "c80c4ef521dd71a6da033d5217b03ccda0d4bed0bce4e4951a16fc8f1977d38b"
    assert h.digest() == bytes(g["seed_sha256"])


def test_random_positions_legal_and_flips(golden):
    g = golden["board"]
    for i in range(len(g["rand_black"])):
        blk, wht, s = int(g["rand_black"][i]), int(g["rand_white"][i]), int(g["rand_side"][i])
        P, O = (blk, wht) if s == 1 else (wht, blk)
        lm = orc.legal(P, O)
        assert lm == int(g["rand_legal"][i])
        m = lm
        while m:
            idx = (m & -m).bit_length() - 1
            m &= m - 1
            assert orc.flips(P, O, idx) == int(g["rand_flips"][i, idx])


def test_illegal_and_over_moves_rejected():
    b = orc.make_board(*orc.START)
    L = orc.lib()
    assert L.orc_apply(C.byref(b), 0, 0) == 0      # (0,0) is not legal at the start
    assert L.orc_apply(C.byref(b), 27, 0) == 0     # occupied
    b.over = 1
    assert L.orc_apply(C.byref(b), 19, 0) == 0     # game.py:47-48


def test_planes(golden):
    g = golden["board"]
    for i in range(len(g["planes_black"])):
        p = orc.planes(g["planes_black"][i], g["planes_white"][i], g["planes_side"][i])
        assert np.array_equal(p, g["planes"][i])
    p = orc.planes(*orc.START)
    assert [int(p[k].sum()) for k in range(3)] == [2, 2, 4]
    assert list(np.nonzero(p[2].reshape(-1))[0]) == [19, 26, 37, 44]
