"""GPU tier (K1/K3): the CUDA board kernels, called through the C ABI, against the golden vectors
of the live reference and against the C oracle on seeded inputs.  Bit-exact everywhere."""
import ctypes as C
import hashlib

import numpy as np
import pytest

import orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def az():
    import alphazero_reversi_b200 as m
    return m


def _pairs(g):
    """(position, legal move) pairs of the random-position golden set"""
    bl, wh, sd, mv, fl = [], [], [], [], []
    for i in range(len(g["rand_black"])):
        m = int(g["rand_legal"][i])
        while m:
            idx = (m & -m).bit_length() - 1
            m &= m - 1
            bl.append(g["rand_black"][i]); wh.append(g["rand_white"][i]); sd.append(g["rand_side"][i])
            mv.append(idx); fl.append(g["rand_flips"][i, idx])
    return (np.array(bl, dtype=np.uint64), np.array(wh, dtype=np.uint64), np.array(sd, dtype=np.uint8),
            np.array(mv, dtype=np.uint8), np.array(fl, dtype=np.uint64))


def test_golden_legal_and_flips(az, golden):
    g = golden["board"]
    lm = az.board_ops.legal_masks(g["rand_black"], g["rand_white"], g["rand_side"])
    assert np.array_equal(lm, g["rand_legal"])
    bl, wh, sd, mv, fl = _pairs(g)
    assert np.array_equal(az.board_ops.flip_masks(bl, wh, sd, mv), fl)


def test_device_pointers_match_host_path(az, golden):
    import torch
    g = golden["board"]
    dev = torch.device("cuda:0")
    tb = torch.from_numpy(g["rand_black"].view(np.int64)).to(dev)
    tw = torch.from_numpy(g["rand_white"].view(np.int64)).to(dev)
    ts = torch.from_numpy(g["rand_side"]).to(dev)
    lm = az.board_ops.legal_masks(tb, tw, ts)
    torch.cuda.synchronize()
    assert np.array_equal(lm.cpu().numpy().view(np.uint64), g["rand_legal"])


def test_seeded_games_stepwise(az, golden):
    """the reference's 100 seeded games (6000 plies) replayed in lockstep through rvs_apply_moves"""
    g = golden["board"]
    G = g["seed_moves"].shape[0]
    bl = np.full(G, orc.START[0], dtype=np.uint64)
    wh = np.full(G, orc.START[1], dtype=np.uint64)
    sd = np.ones(G, dtype=np.uint8)
    fl = np.zeros(G, dtype=np.uint8)
    h = [hashlib.sha256() for _ in range(G)]
    for p in range(int(g["seed_nply"].max())):
        live = g["seed_nply"] > p
        mv = g["seed_moves"][:, p].copy()
        ok, nl = az.board_ops.apply_moves(bl, wh, sd, fl, mv)
        assert np.array_equal(ok.astype(bool), live)  # 255 / finished games are rejected, untouched
        idx = np.nonzero(live)[0]
        assert np.array_equal(bl[idx], g["seed_state"][idx, p + 1, 0])
        assert np.array_equal(wh[idx], g["seed_state"][idx, p + 1, 1])
        assert np.array_equal(sd[idx], g["seed_side"][idx, p + 1])
        assert np.array_equal(nl[idx], g["seed_legal"][idx, p + 1])
        assert np.array_equal(fl[idx] & 1, g["seed_over"][idx, p + 1])
    assert np.array_equal((fl >> 1) & 3, g["seed_winner"])
    # the survey's digest: games in order, states after every ply
    hh = hashlib.sha256()
    for gi in range(G):
        for p in range(1, int(g["seed_nply"][gi]) + 1):
            hh.update(int(g["seed_state"][gi, p, 0]).to_bytes(8, "little") + int(g["seed_state"][gi, p, 1]).to_bytes(8, "little")
                      + bytes([int(g["seed_side"][gi, p])]))
    assert hh.hexdigest() == "c80c4ef521dd71a6da033d5217b03ccda0d4bed0bce4e4951a16fc8f1977d38b"


def test_reference_test_game_py(az, golden):
    """the reference's own tests (test_game.py:7-126) run against the mirror classes"""
    game = az.ReversiGame()
    board = game.get_board_state()
    assert board.shape == (8, 8)
    assert board[3][3] == 2 and board[4][4] == 2 and board[3][4] == 1 and board[4][3] == 1
    assert np.sum(board == 0) == 60
    assert set(game.get_valid_moves()) == {(2, 3), (3, 2), (4, 5), (5, 4)}
    assert game.get_valid_moves() == [(2, 3), (3, 2), (4, 5), (5, 4)]  # ascending bit order
    assert game.make_move(2, 3)
    board = game.get_board_state()
    assert board[2][3] == 1 and board[3][3] == 1 and game.get_current_player() == 2
    assert not game.make_move(0, 0)  # illegal -> False, never raises (game.py:47-48,70)
    # endgame fill (test_game.py:60-126)
    g = golden["board"]
    game = az.ReversiGame(8)
    game.board.black, game.board.white = (int(x) for x in g["endgame_in"])
    game.current_player = game.board.WHITE
    assert game.make_move(0, 0)
    assert game.is_game_over() and game.get_winner() == game.board.WHITE
    eb, ew = int(g["endgame_out"][0]), int(g["endgame_out"][1])
    assert (game.board.black, game.board.white) == (eb, ew)
    assert not game.make_move(0, 1)  # over
    with pytest.raises(ValueError):
        az.ReversiGame(6)


def test_first_last_games_via_mirror(az, golden):
    g = golden["board"]
    for name, pick in (("first", 0), ("last", -1)):
        game = az.ReversiGame()
        moves = []
        while not game.is_game_over():
            r, c = game.get_valid_moves()[pick]
            assert game.make_move(r, c)
            moves.append(r * 8 + c)
        assert moves == list(g[f"{name}_moves"])
        assert (game.board.black, game.board.white) == tuple(int(x) for x in g[f"{name}_final"])
        assert game.get_winner() == int(g[f"{name}_winner"][0])


def test_perft(az, golden):
    exp = [int(x) for x in golden["board"]["perft"]]
    for d, e in enumerate(exp, start=1):
        assert az.board_ops.perft(d) == e
    assert az.board_ops.perft(0) == 1
    # depth 9 differs from the literature's 3005288 because an auto-pass does not consume a ply here
    strict = [4, 12, 56, 244, 1396, 8200, 55092, 390216, orc.perft(9, orc.RULES_STRICT)]
    for d, e in enumerate(strict, start=1):
        assert az.board_ops.perft(d, rules=az.RULES_STRICT) == e
    assert az.board_ops.perft(9) == orc.perft(9)
    assert az.board_ops.perft(10) == orc.perft(10)


@pytest.mark.parametrize("rules", [0, 1])
def test_random_positions_vs_oracle(az, rules):
    rng = np.random.default_rng(5 + rules)
    n = 200000
    occ = rng.integers(0, 2**64, n, dtype=np.uint64) | rng.integers(0, 2**64, n, dtype=np.uint64)
    thin = rng.integers(0, 2**64, n, dtype=np.uint64)
    occ[: n // 3] &= thin[: n // 3]
    pick = rng.integers(0, 2**64, n, dtype=np.uint64)
    bl, wh = occ & pick, occ & ~pick
    sd = rng.integers(1, 3, n, dtype=np.uint8)
    lm = az.board_ops.legal_masks(bl, wh, sd, rules=rules)
    mv = rng.integers(0, 64, n, dtype=np.uint8)
    fl = az.board_ops.flip_masks(bl, wh, sd, mv, rules=rules)
    L = orc.lib()
    for i in range(0, n, 7):
        P, O = (int(bl[i]), int(wh[i])) if sd[i] == 1 else (int(wh[i]), int(bl[i]))
        assert int(lm[i]) == L.orc_legal(P, O, rules)
        assert int(fl[i]) == L.orc_flips(P, O, int(mv[i]), rules)


@pytest.mark.parametrize("rules", [0, 1])
def test_random_playouts_vs_oracle(az, rules):
    n = 20000
    ob, ow, owin, opl = orc.random_playouts(n, 12345, rules)
    bl, wh, wi, pl, total = az.board_ops.random_playouts(n, 12345, rules=rules)
    assert np.array_equal(bl, ob) and np.array_equal(wh, ow)
    assert np.array_equal(wi, owin) and np.array_equal(pl, opl)
    assert total == int(opl.sum())


def test_random_playouts_full_size_properties(az):
    """BASELINE config-1 scale (1M games): totals agree with per-game outputs, games are legal
    final positions (no legal move for either side), and a strided sample matches the oracle."""
    n = 1 << 20
    bl, wh, wi, pl, total = az.board_ops.random_playouts(n, 777)
    assert total == int(pl.astype(np.int64).sum())
    assert not np.any(bl & wh)
    ones = np.ones(n, dtype=np.uint8)
    assert not az.board_ops.legal_masks(bl, wh, ones).any()
    assert not az.board_ops.legal_masks(bl, wh, ones * 2).any()
    L = orc.lib()
    for g in range(0, n, 4099):
        b = orc.make_board(*orc.START)
        p = L.orc_random_playout(C.byref(b), L.orc_stream_seed(777, g, 0), 0)
        assert (b.black, b.white, b.winner, p) == (int(bl[g]), int(wh[g]), int(wi[g]), int(pl[g]))
    # no-output mode (used by the bench) counts the same number of board-steps
    *_, total2 = az.board_ops.random_playouts(n, 777, outputs=False)
    assert total2 == total


def test_encode_planes(az, golden):
    g = golden["board"]
    p = az.board_ops.encode_planes(g["planes_black"], g["planes_white"], g["planes_side"])
    assert p.dtype == np.float32 and np.array_equal(p, g["planes"])
    # bf16 NHWC16 carries the same three planes in channels 0..2
    q = az.board_ops.encode_planes(g["planes_black"], g["planes_white"], g["planes_side"], layout=1)
    assert q.shape == (len(p), 8, 8, 16)
    assert np.array_equal(q[..., :3] == 0x3F80, np.transpose(p, (0, 2, 3, 1)) == 1.0)
    assert not q[..., 3:].any() and set(np.unique(q)) <= {0, 0x3F80}
    game = az.ReversiGame()
    s = game.get_canonical_state()
    assert s.shape == (3, 8, 8) and [int(s[k].sum()) for k in range(3)] == [2, 2, 4]


def test_edge_cases(az):
    e = np.zeros(0, dtype=np.uint64)
    assert len(az.board_ops.legal_masks(e, e, np.zeros(0, dtype=np.uint8))) == 0
    # full board / empty board / single colour
    bl = np.array([0, 2**64 - 1, 0, 0x00000000FFFFFFFF], dtype=np.uint64)
    wh = np.array([0, 0, 2**64 - 1, 0xFFFFFFFF00000000], dtype=np.uint64)
    sd = np.array([1, 2, 1, 2], dtype=np.uint8)
    lm = az.board_ops.legal_masks(bl, wh, sd)
    assert not lm.any()
    fl = np.zeros(4, dtype=np.uint8)
    ok, nl = az.board_ops.apply_moves(bl, wh, sd, fl, np.array([0, 5, 64, 255], dtype=np.uint8))
    assert not ok.any() and not nl.any()
    with pytest.raises(az.RvsError):
        az.board_ops.perft(-1)
