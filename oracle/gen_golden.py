#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference from /root/reference.

TEST INFRASTRUCTURE ONLY.  Run once in the build container (the GPU box has no
/root/reference); the outputs are committed.  Pins: numpy 2.3.5 (NEP-50 f32 tree
arithmetic) and torch 2.11.0 (softmax / network numerics).

    CUDA_VISIBLE_DEVICES="" python oracle/gen_golden.py [--only board,mcts,net,selfplay]

Reference entry points exercised (paths relative to /root/reference):
  src/game/board.py:70-133,135-251   Board.get_valid_moves / make_move
  src/game/game.py:36-70,131-162     ReversiGame.make_move / get_canonical_state
  src/mcts/mcts.py:322-407,642-694   MCTS.search / get_action_probs
  src/model/network.py:80-158        AlphaZeroNetwork.forward / predict
  src/self_play/self_play.py:51-145  SelfPlay.generate_games
"""
import argparse
import contextlib
import hashlib
import io
import os
import random
import sys
import tempfile

os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")
sys.path.insert(0, "/root/reference")

import numpy as np
import torch

from src.game.board import Board  # noqa: E402
from src.game.game import ReversiGame  # noqa: E402
from src.mcts.mcts import MCTS  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
M64 = (1 << 64) - 1


def mix64(x):
    x &= M64
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & M64
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & M64
    x ^= x >> 31
    return x


# ----------------------------------------------------------------------------- board
def perft(board, depth):
    if depth == 0 or board.game_over:
        return 1
    n = 0
    for (r, c) in board.get_valid_moves():
        b = board.copy()
        assert b.make_move(r, c)
        n += perft(b, depth - 1)
    return n


def play_policy(pick):
    g = ReversiGame()
    moves = []
    while not g.is_game_over():
        vm = g.get_valid_moves()
        r, c = pick(vm)
        assert g.make_move(r, c)
        moves.append(r * 8 + c)
    return g, moves


def gen_board(perft_depth):
    out = {}
    out["perft"] = np.array([perft(Board(), d) for d in range(1, perft_depth + 1)], dtype=np.uint64)
    print("perft", out["perft"])
    for name, pick in (("first", lambda vm: vm[0]), ("last", lambda vm: vm[-1])):
        g, moves = play_policy(pick)
        out[f"{name}_moves"] = np.array(moves, dtype=np.uint8)
        out[f"{name}_final"] = np.array([g.board.black, g.board.white], dtype=np.uint64)
        out[f"{name}_winner"] = np.array([g.get_winner()], dtype=np.int64)
        print(name, len(moves), hex(g.board.black), hex(g.board.white), g.get_winner())

    # 100 seeded games, state + legal mask + per-move flips after every ply
    rng = random.Random(12345)
    h = hashlib.sha256()
    G, maxp = 100, 64
    moves = np.full((G, maxp), 255, dtype=np.uint8)
    st = np.zeros((G, maxp + 1, 2), dtype=np.uint64)  # black, white (index 0 = start)
    side = np.zeros((G, maxp + 1), dtype=np.uint8)
    legal = np.zeros((G, maxp + 1), dtype=np.uint64)
    over = np.zeros((G, maxp + 1), dtype=np.uint8)
    winner = np.zeros(G, dtype=np.uint8)
    nply = np.zeros(G, dtype=np.int32)
    wins = {0: 0, 1: 0, 2: 0}
    total = 0
    for gi in range(G):
        g = ReversiGame()
        p = 0
        while True:
            st[gi, p] = (g.board.black, g.board.white)
            side[gi, p] = g.current_player
            lm = 0
            for (r, c) in g.get_valid_moves():
                lm |= 1 << (r * 8 + c)
            legal[gi, p] = lm
            over[gi, p] = int(g.is_game_over())
            if g.is_game_over():
                break
            vm = g.get_valid_moves()
            r, c = vm[rng.randrange(len(vm))]
            assert g.make_move(r, c)
            moves[gi, p] = r * 8 + c
            p += 1
            h.update(g.board.black.to_bytes(8, "little") + g.board.white.to_bytes(8, "little")
                     + bytes([g.current_player]))
        nply[gi] = p
        total += p
        winner[gi] = g.get_winner()
        wins[g.get_winner()] += 1
    print("seeded games", total, wins, h.hexdigest())
    out.update(seed_moves=moves, seed_state=st, seed_side=side, seed_legal=legal,
               seed_over=over, seed_winner=winner, seed_nply=nply,
               seed_sha256=np.frombuffer(h.digest(), dtype=np.uint8))

    # arbitrary (mostly unreachable) disjoint bitboards: legal mask + flips of every legal move
    rr = random.Random(777)
    NPOS = 3000
    pb = np.zeros(NPOS, dtype=np.uint64)
    pw = np.zeros(NPOS, dtype=np.uint64)
    ps = np.zeros(NPOS, dtype=np.uint8)
    pl = np.zeros(NPOS, dtype=np.uint64)
    pf = np.zeros((NPOS, 64), dtype=np.uint64)
    for i in range(NPOS):
        dens = rr.choice([0.15, 0.3, 0.5, 0.7, 0.9, 0.97])
        occ = 0
        for b in range(64):
            if rr.random() < dens:
                occ |= 1 << b
        blk = occ & rr.getrandbits(64)
        wht = occ & ~blk
        bd = Board()
        bd.black, bd.white = blk, wht
        player = rr.choice([1, 2])
        bd.current_player = player
        bd._update_board_state()
        lm = 0
        for (r, c) in bd.get_valid_moves(player):
            lm |= 1 << (r * 8 + c)
            fl = 0
            for (fr, fc) in bd._get_flipped_pieces((r, c), player):
                fl |= 1 << (fr * 8 + fc)
            # cross-check that make_move applies exactly this flip mask
            b2 = bd.copy()
            assert b2.make_move(r, c, player)
            mine, theirs = (b2.black, b2.white) if player == 1 else (b2.white, b2.black)
            P, O = (blk, wht) if player == 1 else (wht, blk)
            assert mine == P ^ ((1 << (r * 8 + c)) | fl) and theirs == O ^ fl
            pf[i, r * 8 + c] = fl
        pb[i], pw[i], ps[i], pl[i] = blk, wht, player, lm
    out.update(rand_black=pb, rand_white=pw, rand_side=ps, rand_legal=pl, rand_flips=pf)

    # canonical planes at a few reachable positions
    rng = random.Random(99)
    cb, cw, cs, cp = [], [], [], []
    for _ in range(6):
        g = ReversiGame()
        n = rng.randrange(0, 58)
        for _ in range(n):
            if g.is_game_over():
                break
            vm = g.get_valid_moves()
            g.make_move(*vm[rng.randrange(len(vm))])
        cb.append(g.board.black); cw.append(g.board.white); cs.append(g.current_player)
        cp.append(g.get_canonical_state())
    out.update(planes_black=np.array(cb, dtype=np.uint64), planes_white=np.array(cw, dtype=np.uint64),
               planes_side=np.array(cs, dtype=np.uint8), planes=np.array(cp, dtype=np.float32))

    # reference-owned pins from test_game.py:60-126 (endgame fill position)
    g = ReversiGame(8)
    blk, wht = 0x2, 0
    for i in range(8):
        for j in range(8):
            if i > 0 or j > 1:
                if (i + j) % 2 == 0:
                    wht |= 1 << (i * 8 + j)
                else:
                    blk |= 1 << (i * 8 + j)
    g.board.black, g.board.white = blk, wht
    g.current_player = 2
    g.board._update_board_state()
    ok = g.make_move(0, 0)
    out["endgame_in"] = np.array([blk, wht], dtype=np.uint64)
    out["endgame_out"] = np.array([g.board.black, g.board.white, int(ok), int(g.is_game_over()),
                                   g.get_winner()], dtype=np.uint64)
    print("endgame", ok, g.is_game_over(), g.get_winner(), g.get_score())
    np.savez_compressed(os.path.join(OUT, "board.npz"), **out)


# ----------------------------------------------------------------------------- mcts
def planes_to_bits(x):
    own = opp = 0
    p0 = x[0].reshape(-1)
    p1 = x[1].reshape(-1)
    for i in range(64):
        if p0[i] > 0.5:
            own |= 1 << i
        if p1[i] > 0.5:
            opp |= 1 << i
    return own, opp


class StubModel:
    """model duck-type the reference MCTS needs (mcts.py:211,235,501)."""

    def __init__(self, kind, log=None):
        self.kind = kind
        self.log = log
        self._p = torch.nn.Parameter(torch.zeros(1))

    def parameters(self):
        return iter([self._p])

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def predict(self, x):
        x = x.detach().cpu().numpy()
        B = x.shape[0]
        logits = np.zeros((B, 65), dtype=np.float32)
        values = np.zeros((B,), dtype=np.float32)
        for b in range(B):
            own, opp = planes_to_bits(x[b])
            if self.kind == "E0":
                values[b] = np.float32(bin(own).count("1") - bin(opp).count("1")) / np.float32(64)
            else:
                h = mix64((own * 0x9E3779B97F4A7C15) ^ mix64(opp))
                values[b] = np.float32(((h >> 20) & 0xFFFF) - 32768) / np.float32(32768)
                if self.kind == "T1":  # uniform over a hashed subset (+ pass), exact f32(1/k)
                    sub = mix64(h ^ 0xC2B2AE3D27D4EB4F)
                    for i in range(64):
                        logits[b, i] = 0.0 if (sub >> i) & 1 else -np.inf
                else:  # "T2": rich logits; softmax outputs are stored in the fixture
                    for i in range(65):
                        logits[b, i] = np.float32(((mix64(h + i) >> 40) & 0xFF)) / np.float32(32) - np.float32(4)
        lt = torch.from_numpy(logits)
        if self.log is not None:
            pr = torch.softmax(lt, dim=1).numpy()
            for b in range(B):
                own, opp = planes_to_bits(x[b])
                self.log[(own, opp)] = (pr[b].copy(), values[b])
        return lt, torch.from_numpy(values)


def midgame_positions():
    """deterministic set of root positions: start + seeded-game positions at several plies"""
    pos = [(0x0000000810000000, 0x0000001008000000, 1)]
    pos.append((0x000000081C0A000E, 0x0000001000040211, 1))
    rng = random.Random(2024)
    for target in (6, 13, 21, 30, 38, 45, 51, 55, 57, 58, 59):
        g = ReversiGame()
        for _ in range(target):
            if g.is_game_over():
                break
            vm = g.get_valid_moves()
            g.make_move(*vm[rng.randrange(len(vm))])
        if not g.is_game_over():
            pos.append((g.board.black, g.board.white, g.current_player))
    return pos


def set_position(blk, wht, side):
    g = ReversiGame()
    g.board.black, g.board.white = blk, wht
    g.board.current_player = side
    g.current_player = side
    g.board._update_board_state()
    return g


def gen_mcts():
    out = {}
    cases = []  # (kind, pos_idx, S, K, c_puct)
    pos = midgame_positions()
    for kind in ("E0", "T1"):
        for pi in range(len(pos)):
            for (S, K) in ((100, 1), (100, 8), (100, 64), (200, 16), (400, 64)):
                cases.append((kind, pi, S, K, 1.0))
    for pi in (0, 1, 4, 7):
        cases.append(("E0", pi, 800, 64, 1.0))
        cases.append(("T1", pi, 800, 1, 1.0))
        cases.append(("T1", pi, 300, 1, 2.5))
        cases.append(("T1", pi, 300, 4, 0.7))
    table = {}
    for pi in (0, 3, 6, 9):
        cases.append(("T2", pi, 150, 1, 1.0))
        cases.append(("T2", pi, 150, 8, 1.5))
    rows, vis, rootw = [], [], []
    for (kind, pi, S, K, c) in cases:
        blk, wht, side = pos[pi]
        g = set_position(blk, wht, side)
        model = StubModel(kind, table if kind == "T2" else None)
        m = MCTS(model, c_puct=c, num_simulations=S, batch_size=K)
        counts = m.search(g)
        v = np.zeros(65, dtype=np.int32)
        for (r, cc), n in counts.items():
            v[r * 8 + cc] = n
        rows.append((["E0", "T1", "T2"].index(kind), pi, S, K))
        vis.append(v)
        rootw.append((np.float32(m.root.value_sum), np.float32(c), np.float32(m.root.visit_count)))
        print(kind, pi, S, K, c, dict((k, int(x)) for k, x in enumerate(v) if x), float(m.root.value_sum))
    out["pos"] = np.array([(b, w, s) for (b, w, s) in pos], dtype=np.uint64)
    out["cases"] = np.array(rows, dtype=np.int32)
    out["visits"] = np.array(vis, dtype=np.int32)
    out["root_w_c_n"] = np.array(rootw, dtype=np.float32)
    keys = sorted(table.keys())
    out["t2_keys"] = np.array(keys, dtype=np.uint64).reshape(-1, 2)
    out["t2_probs"] = np.array([table[k][0] for k in keys], dtype=np.float32)
    out["t2_values"] = np.array([table[k][1] for k in keys], dtype=np.float32)
    # prior actually produced by torch softmax for all-zero logits and for subset sizes
    out["uniform_prior"] = torch.softmax(torch.zeros(1, 65), dim=1).numpy()[0, :1]
    ks = []
    for k in range(1, 66):
        lt = torch.full((1, 65), -np.inf)
        lt[0, :k] = 0.0
        ks.append(torch.softmax(lt, dim=1).numpy()[0, 0])
    out["subset_prior"] = np.array(ks, dtype=np.float32)

    # get_action_probs (mcts.py:642-694): pi for T in {1, 0.5, 0} at the start position
    g = ReversiGame()
    ap = []
    for T in (1.0, 0.5):
        m = MCTS(StubModel("E0"), c_puct=1.0, num_simulations=400, batch_size=64)
        np.random.seed(7)
        a, p = m.get_action_probs(g, temperature=T)
        ap.append(p)
    out["ap_pi"] = np.array(ap, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "mcts.npz"), **out)
    print("t2 table", len(keys))


# ----------------------------------------------------------------------------- self-play
def gen_selfplay():
    from src.self_play.self_play import SelfPlay
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for tag, kind, S, T, seed in (("e0_t1", "E0", 100, 1.0, 11), ("t1_t1", "T1", 100, 1.0, 123),
                                       ("t1_t05", "T1", 130, 0.5, 5)):
            # temperature 0 cannot be pinned: the reference's argmax path returns np.int64
            # coordinates and Board.make_move then raises OverflowError (board.py:213)
            np.random.seed(seed)
            sp = SelfPlay(StubModel(kind), {"num_simulations": S, "c_puct": 1.0, "temperature": T,
                                            "save_dir": td})
            with contextlib.redirect_stdout(io.StringIO()):
                games = sp.generate_games(1)
            gd = games[0]
            out[f"{tag}_states"] = np.array(gd["states"], dtype=np.float32)
            out[f"{tag}_pi"] = np.array(gd["action_probs"], dtype=np.float64)
            out[f"{tag}_players"] = np.array(gd["current_players"], dtype=np.int32)
            out[f"{tag}_z"] = np.array(gd["values"], dtype=np.float32)
            out[f"{tag}_cfg"] = np.array([S, 64, T, seed], dtype=np.float64)
            print(tag, len(gd["states"]), gd["values"][:4])
    np.savez_compressed(os.path.join(OUT, "selfplay.npz"), **out)


# ----------------------------------------------------------------------------- network
def perturb_bn(model, seed):
    """deterministic non-trivial BN statistics so that BN folding is really tested"""
    gen = torch.Generator().manual_seed(seed)
    for name, m in model.named_modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=gen) * 0.2)
            m.running_var.copy_(torch.rand(m.num_features, generator=gen) * 1.5 + 0.25)
            m.weight.data.copy_(torch.rand(m.num_features, generator=gen) * 0.5 + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=gen) * 0.1)
    model.value_fc2.weight.data.mul_(0.05)  # keep tanh out of saturation


def tame(model):
    """a deep random-init tower saturates (20x256: |logits| ~ 1e3, value = +-1): damp the residual branches
    and the policy head so that priors and values of the deep network are informative test targets"""
    for blk in model.res_blocks:
        blk.bn2.weight.data.mul_(0.1)
        blk.bn2.bias.data.mul_(0.1)
    model.policy_fc.weight.data.mul_(0.1)


def net_positions():
    pos = midgame_positions()
    rng = random.Random(5)
    for _ in range(19):
        g = ReversiGame()
        for _ in range(rng.randrange(1, 59)):
            if g.is_game_over():
                break
            vm = g.get_valid_moves()
            g.make_move(*vm[rng.randrange(len(vm))])
        if not g.is_game_over():
            pos.append((g.board.black, g.board.white, g.current_player))
    return pos


def gen_net(which=(("5x128", 5, 128), ("2x64", 2, 64)), fname="net.npz"):
    from src.model.network import AlphaZeroNetwork
    out = {}
    pos = net_positions()
    planes = np.array([set_position(*p).get_canonical_state() for p in pos], dtype=np.float32)
    out["pos"] = np.array(pos, dtype=np.uint64)
    for tag, nb, nf in which:
        torch.manual_seed(42)
        net = AlphaZeroNetwork(8, nb, nf)
        sd = {k: v for k, v in net.state_dict().items()}
        out[f"{tag}_wsum"] = np.array([float(v.double().sum()) for v in sd.values()], dtype=np.float64)
        out[f"{tag}_wabs"] = np.array([float(v.double().abs().sum()) for v in sd.values()], dtype=np.float64)
        net.eval()
        with torch.no_grad():
            lg, vl = net.predict(torch.from_numpy(planes))
        out[f"{tag}_fresh_logits"] = lg.numpy()
        out[f"{tag}_fresh_values"] = vl.numpy()
        torch.manual_seed(42)
        net = AlphaZeroNetwork(8, nb, nf)
        with torch.no_grad():
            perturb_bn(net, 43)
        net.eval()
        with torch.no_grad():
            lg, vl = net.predict(torch.from_numpy(planes))
            x16 = torch.from_numpy(planes)
            with torch.autocast("cpu", dtype=torch.bfloat16):
                lg16, vl16 = net.forward(x16)
        out[f"{tag}_bn_logits"] = lg.numpy()
        out[f"{tag}_bn_values"] = vl.numpy()
        out[f"{tag}_bn_logits_autocast"] = lg16.float().numpy()
        out[f"{tag}_bn_values_autocast"] = vl16.float().numpy()
        print(tag, lg.abs().max().item(), vl[:5].tolist())
        if nb >= 20:  # third variant for the deep tower: perturbed BN + damped residual branches (unsaturated outputs)
            torch.manual_seed(42)
            net = AlphaZeroNetwork(8, nb, nf)
            with torch.no_grad():
                perturb_bn(net, 43)
                tame(net)
            net.eval()
            with torch.no_grad():
                lg, vl = net.predict(torch.from_numpy(planes))
                with torch.autocast("cpu", dtype=torch.bfloat16):
                    lg16, vl16 = net.forward(torch.from_numpy(planes))
            out[f"{tag}_tamed_logits"] = lg.numpy()
            out[f"{tag}_tamed_values"] = vl.numpy()
            out[f"{tag}_tamed_logits_autocast"] = lg16.float().numpy()
            out[f"{tag}_tamed_values_autocast"] = vl16.float().numpy()
            print(tag, "tamed", lg.abs().max().item(), vl[:5].tolist())
    np.savez_compressed(os.path.join(OUT, fname), **out)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="board,mcts,selfplay,net")
    ap.add_argument("--perft-depth", type=int, default=8)
    a = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    todo = a.only.split(",")
    if "board" in todo:
        gen_board(a.perft_depth)
    if "mcts" in todo:
        gen_mcts()
    if "selfplay" in todo:
        gen_selfplay()
    if "net" in todo:
        gen_net()
    if "net20" in todo:  # BASELINE config 4's network at full depth: AlphaZeroNetwork(8, 20, 256) (network.py:80-117)
        gen_net((("20x256", 20, 256),), "net20.npz")
