"""ctypes binding of the CPU oracle (oracle/liborc.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN = os.path.join(ROOT, "tests", "golden")

RULES_REF, RULES_STRICT = 0, 1
EVAL_E0, EVAL_ROLLOUT, EVAL_CALLBACK = 0, 1, 2
START = (0x0000000810000000, 0x0000001008000000, 1)
M64 = (1 << 64) - 1


class Board(C.Structure):
    _fields_ = [("black", C.c_uint64), ("white", C.c_uint64), ("side", C.c_uint8),
                ("over", C.c_uint8), ("winner", C.c_uint8), ("passes", C.c_uint8)]


class Sample(C.Structure):
    _fields_ = [("black", C.c_uint64), ("white", C.c_uint64), ("side", C.c_uint8),
                ("z", C.c_int8), ("move", C.c_uint8), ("pad", C.c_uint8),
                ("visits", C.c_int32 * 65)]


EVAL_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(Board), C.c_int, C.POINTER(C.c_float),
                      C.POINTER(C.c_float))

_lib = None


def build():
    src = os.path.join(ORACLE_DIR, "rvs_oracle.c")
    so = os.path.join(ORACLE_DIR, "liborc.so")
    if (not os.path.exists(so)) or os.path.getmtime(so) < max(
            os.path.getmtime(src), os.path.getmtime(os.path.join(ORACLE_DIR, "rvs_oracle.h"))):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return so


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        u64, i32, i64, u8p = C.c_uint64, C.c_int, C.c_int64, C.POINTER(C.c_uint8)
        L.orc_legal.restype = u64
        L.orc_legal.argtypes = [u64, u64, i32]
        L.orc_set_root_noise.restype = None
        L.orc_set_root_noise.argtypes = [C.c_double, C.c_float]
        L.orc_dirichlet.restype = None
        L.orc_dirichlet.argtypes = [C.c_double, i32, u64, C.POINTER(C.c_float)]
        L.orc_det_log.restype = C.c_double
        L.orc_det_log.argtypes = [C.c_double]
        L.orc_det_exp.restype = C.c_double
        L.orc_det_exp.argtypes = [C.c_double]
        L.orc_flips.restype = u64
        L.orc_flips.argtypes = [u64, u64, i32, i32]
        L.orc_apply.restype = i32
        L.orc_apply.argtypes = [C.POINTER(Board), i32, i32]
        L.orc_board_legal.restype = u64
        L.orc_board_legal.argtypes = [C.POINTER(Board), i32]
        L.orc_perft.restype = u64
        L.orc_perft.argtypes = [C.POINTER(Board), i32, i32]
        L.orc_planes.argtypes = [C.POINTER(Board), i32, C.POINTER(C.c_float)]
        L.orc_mix64.restype = u64
        L.orc_mix64.argtypes = [u64]
        L.orc_stream_seed.restype = u64
        L.orc_stream_seed.argtypes = [u64, u64, u64]
        L.orc_random_playout.restype = i32
        L.orc_random_playout.argtypes = [C.POINTER(Board), u64, i32]
        L.orc_random_playouts.argtypes = [i64, u64, i32, C.POINTER(u64), C.POINTER(u64), u8p, u8p]
        L.orc_mcts_search.restype = i32
        L.orc_mcts_search.argtypes = [C.POINTER(Board), i32, i32, C.c_float, i32, i32, EVAL_FN,
                                      C.c_void_p, u64, u64, u64, C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32), C.POINTER(C.c_float),
                                      C.POINTER(C.c_int64)]
        L.orc_set_search_mode.restype = None
        L.orc_set_search_mode.argtypes = [i32]
        L.orc_mcts_search_fast.restype = i32
        L.orc_mcts_search_fast.argtypes = [C.POINTER(Board), i32, i32, C.c_float, i32, i32, EVAL_FN,
                                           C.c_void_p, u64, u64, u64, C.POINTER(C.c_int32),
                                           C.POINTER(C.c_int32), C.POINTER(C.c_float),
                                           C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.orc_action_probs.argtypes = [C.POINTER(C.c_int32), C.c_double, C.POINTER(C.c_double)]
        L.orc_self_play_game.restype = i32
        L.orc_self_play_game.argtypes = [i32, i32, C.c_float, i32, i32, EVAL_FN, C.c_void_p, u64,
                                         u64, C.c_double, C.POINTER(Sample), i32, u8p]
        L.orc_search_batch.restype = i32
        L.orc_search_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, i32, i32, i32, C.c_float, i32, i32,
                                       u64, u64, u64, C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        _lib = L
    return _lib


def make_board(black, white, side, rules=RULES_REF):
    """board with over/winner derived like Board would have at the time the game ended"""
    b = Board(int(black), int(white), int(side), 0, 0, 0)
    L = lib()
    if L.orc_board_legal(C.byref(b), rules) == 0:
        o = Board(int(black), int(white), 3 - int(side), 0, 0, 0)
        if L.orc_board_legal(C.byref(o), rules) == 0:
            nb, nw = bin(int(black)).count("1"), bin(int(white)).count("1")
            b.over = 1
            b.winner = 1 if nb > nw else 2 if nw > nb else 0
    return b


def legal(P, O, rules=RULES_REF):
    return lib().orc_legal(int(P), int(O), rules)


def flips(P, O, idx, rules=RULES_REF):
    return lib().orc_flips(int(P), int(O), int(idx), rules)


def perft(depth, rules=RULES_REF, pos=START):
    b = make_board(*pos, rules=rules)
    return lib().orc_perft(C.byref(b), depth, rules)


def planes(black, white, side, rules=RULES_REF):
    b = Board(int(black), int(white), int(side), 0, 0, 0)
    out = np.zeros(192, dtype=np.float32)
    lib().orc_planes(C.byref(b), rules, out.ctypes.data_as(C.POINTER(C.c_float)))
    return out.reshape(3, 8, 8)


def random_playouts(n, seed, rules=RULES_REF):
    bl = np.zeros(n, dtype=np.uint64)
    wh = np.zeros(n, dtype=np.uint64)
    wi = np.zeros(n, dtype=np.uint8)
    pl = np.zeros(n, dtype=np.uint8)
    lib().orc_random_playouts(n, seed, rules, bl.ctypes.data_as(C.POINTER(C.c_uint64)),
                              wh.ctypes.data_as(C.POINTER(C.c_uint64)),
                              wi.ctypes.data_as(C.POINTER(C.c_uint8)),
                              pl.ctypes.data_as(C.POINTER(C.c_uint8)))
    return bl, wh, wi, pl


def _wrap_eval(pyfn):
    """pyfn(list[(black, white, side)]) -> (probs[n,65] f32, values[n] f32)"""
    def cb(ctx, leaves, n, probs, values):
        pos = [(leaves[i].black, leaves[i].white, leaves[i].side) for i in range(n)]
        p, v = pyfn(pos)
        p = np.ascontiguousarray(p, dtype=np.float32).reshape(n, 65)
        v = np.ascontiguousarray(v, dtype=np.float32).reshape(n)
        C.memmove(probs, p.ctypes.data, n * 65 * 4)
        C.memmove(values, v.ctypes.data, n * 4)
    return EVAL_FN(cb)


def mcts_search(pos, num_sims, wave, c_puct=1.0, rules=RULES_REF, evaluator=EVAL_E0, pyfn=None,
                seed=0, game_id=0, search_id=0):
    b = make_board(*pos, rules=rules)
    vis = (C.c_int32 * 65)()
    rn = C.c_int32(0)
    rw = C.c_float(0)
    ne = C.c_int64(0)
    fn = _wrap_eval(pyfn) if pyfn is not None else EVAL_FN()
    rc = lib().orc_mcts_search(C.byref(b), num_sims, wave, c_puct, rules, evaluator, fn, None, seed,
                               game_id, search_id, vis, C.byref(rn), C.byref(rw), C.byref(ne))
    if rc < 0:
        raise RuntimeError(f"orc_mcts_search failed rc={rc}")
    return np.array(vis[:], dtype=np.int32), rn.value, np.float32(rw.value), ne.value


def mcts_search_fast(pos, num_sims, wave, c_puct=1.0, rules=RULES_REF, evaluator=EVAL_E0, pyfn=None,
                     seed=0, game_id=0, search_id=0):
    """the engine's FAST search mode (virtual-loss PUCT; specification in oracle/rvs_oracle.c -- not reference
    behaviour).  Returns (visits, root_n, root_w, evals consumed, unique evaluations)."""
    b = make_board(*pos, rules=rules)
    vis = (C.c_int32 * 65)()
    rn = C.c_int32(0)
    rw = C.c_float(0)
    ne = C.c_int64(0)
    nu = C.c_int64(0)
    fn = _wrap_eval(pyfn) if pyfn is not None else EVAL_FN()
    rc = lib().orc_mcts_search_fast(C.byref(b), num_sims, wave, c_puct, rules, evaluator, fn, None, seed,
                                    game_id, search_id, vis, C.byref(rn), C.byref(rw), C.byref(ne), C.byref(nu))
    if rc < 0:
        raise RuntimeError(f"orc_mcts_search_fast failed rc={rc}")
    return np.array(vis[:], dtype=np.int32), rn.value, np.float32(rw.value), ne.value, nu.value


def set_search_mode(mode):
    """0 = reference-compatible (graded), 1 = FAST: routes mcts_search / self_play_game / search_batch"""
    lib().orc_set_search_mode(int(mode))


def set_root_noise(alpha, eps):
    """engine feature restated by the oracle: Dirichlet noise on the root priors (0 = off)"""
    lib().orc_set_root_noise(float(alpha), float(eps))


def action_probs(visits, temperature):
    v = (C.c_int32 * 65)(*[int(x) for x in visits])
    pi = (C.c_double * 65)()
    lib().orc_action_probs(v, temperature, pi)
    return np.array(pi[:], dtype=np.float64)


def self_play_game(num_sims, wave, c_puct=1.0, rules=RULES_REF, evaluator=EVAL_E0, pyfn=None,
                   seed=0, game_id=0, temperature=1.0):
    out = (Sample * 64)()
    win = C.c_uint8(0)
    fn = _wrap_eval(pyfn) if pyfn is not None else EVAL_FN()
    n = lib().orc_self_play_game(num_sims, wave, c_puct, rules, evaluator, fn, None, seed, game_id,
                                 temperature, out, 64, C.byref(win))
    if n < 0:
        raise RuntimeError(f"orc_self_play_game failed rc={n}")
    return [out[i] for i in range(n)], win.value


# ---- the deterministic stub evaluators of oracle/gen_golden.py, restated ------------
def mix64(x):
    x &= M64
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & M64
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & M64
    x ^= x >> 31
    return x


def own_opp(black, white, side):
    return (black, white) if side == 1 else (white, black)


def t1_eval(pos, subset_prior):
    """T1: uniform prior over a hashed subset of squares (+pass); value from the hash"""
    n = len(pos)
    probs = np.zeros((n, 65), dtype=np.float32)
    values = np.zeros(n, dtype=np.float32)
    for i, (b, w, s) in enumerate(pos):
        own, opp = own_opp(b, w, s)
        h = mix64((own * 0x9E3779B97F4A7C15) ^ mix64(opp))
        values[i] = np.float32(((h >> 20) & 0xFFFF) - 32768) / np.float32(32768)
        sub = mix64(h ^ 0xC2B2AE3D27D4EB4F)
        k = bin(sub).count("1") + 1
        p = subset_prior[k - 1]
        for q in range(64):
            if (sub >> q) & 1:
                probs[i, q] = p
        probs[i, 64] = p
    return probs, values


def t1_hash_value(b, w, s):
    own, opp = own_opp(b, w, s)
    h = mix64((own * 0x9E3779B97F4A7C15) ^ mix64(opp))
    return h


def search_batch(black, white, side, num_sims, wave, c_puct=1.0, rules=RULES_REF, evaluator=EVAL_ROLLOUT,
                 seed=0, game0=0, search_id=0, threads=1):
    """oracle search over a batch of roots, split over `threads` host threads (ctypes drops the GIL).
    Returns (visits [n,65] int32, evals, board_steps)."""
    from concurrent.futures import ThreadPoolExecutor
    n = len(black)
    black = np.ascontiguousarray(black, dtype=np.uint64)
    white = np.ascontiguousarray(white, dtype=np.uint64)
    side = np.ascontiguousarray(side, dtype=np.uint8)
    vis = np.zeros((n, 65), dtype=np.int32)
    L = lib()
    bounds = np.linspace(0, n, max(1, threads) + 1).astype(int)

    def work(t):
        lo, hi = int(bounds[t]), int(bounds[t + 1])
        ev, st = C.c_int64(0), C.c_int64(0)
        if hi > lo:
            rc = L.orc_search_batch(black[lo:].ctypes.data, white[lo:].ctypes.data, side[lo:].ctypes.data, hi - lo,
                                    num_sims, wave, c_puct, rules, evaluator, seed, game0 + lo, search_id,
                                    vis[lo:].ctypes.data, C.byref(ev), C.byref(st))
            if rc:
                raise RuntimeError(f"orc_search_batch rc={rc}")
        return ev.value, st.value
    if threads <= 1:
        res = [work(0)]
    else:
        with ThreadPoolExecutor(max_workers=threads) as ex:
            res = list(ex.map(work, range(threads)))
    return vis, sum(r[0] for r in res), sum(r[1] for r in res)
