"""K1/K3 batch operations over the C ABI.  Arguments are numpy arrays (host memory; the library
stages them) or contiguous torch CUDA tensors (used in place).  dtypes: black/white uint64
(torch: int64 bit patterns), side/move/flags uint8."""
import ctypes as C

import numpy as np

from . import _lib as L


def _torch():
    import torch
    return torch


def _like(x, shape, np_dtype, torch_dtype_name):
    if hasattr(x, "data_ptr"):
        t = _torch()
        return t.empty(shape, dtype=getattr(t, torch_dtype_name), device=x.device)
    return np.empty(shape, dtype=np_dtype)


def _p(x, dtype=None):
    return L.ptr(x, dtype)[0]


def _u64(x):
    return L.ptr(x, "uint64")[0]


def _u8(x):
    return L.ptr(x, "uint8")[0]


def legal_masks(black, white, side, rules=L.RULES_REF, stream=None):
    """Board.get_valid_moves for n positions -> uint64 masks (reference: src/game/board.py:70-133)"""
    n = len(black)
    out = _like(black, (n,), np.uint64, "int64")
    mem = L.mem_of(black, white, side, out)
    L.check(L.lib().rvs_legal_masks(_u64(black), _u64(white), _u8(side), _u64(out), n, rules, mem,
                                    stream if stream is not None else L.current_stream()))
    return out


def flip_masks(black, white, side, move, rules=L.RULES_REF, stream=None):
    """flip scan of Board.make_move (src/game/board.py:190-219)"""
    n = len(black)
    out = _like(black, (n,), np.uint64, "int64")
    mem = L.mem_of(black, white, side, move, out)
    L.check(L.lib().rvs_flip_masks(_u64(black), _u64(white), _u8(side), _u8(move), _u64(out), n, rules, mem,
                                   stream if stream is not None else L.current_stream()))
    return out


def apply_moves(black, white, side, flags, move, rules=L.RULES_REF, want_legal=True, stream=None):
    """ReversiGame.make_move in place (src/game/game.py:36-70); returns (ok, next_legal)"""
    n = len(black)
    ok = _like(black, (n,), np.uint8, "uint8")
    nl = _like(black, (n,), np.uint64, "int64") if want_legal else None
    mem = L.mem_of(black, white, side, flags, move, ok)
    L.check(L.lib().rvs_apply_moves(_u64(black), _u64(white), _u8(side), _u8(flags), _u8(move), _u8(ok), _u64(nl), n,
                                    rules, mem, stream if stream is not None else L.current_stream()))
    return ok, nl


def random_playouts(n_games, seed, first_game=0, rules=L.RULES_REF, device=None, outputs=True, stream=None):
    """n uniform-random games from the start (BASELINE config 1).  Returns
    (black, white, winner, plies, total_plies); arrays are None when outputs=False."""
    total = C.c_int64(0)
    if outputs:
        if device is not None:
            t = _torch()
            bl = t.empty(n_games, dtype=t.int64, device=device)
            wh = t.empty(n_games, dtype=t.int64, device=device)
            wi = t.empty(n_games, dtype=t.uint8, device=device)
            pl = t.empty(n_games, dtype=t.uint8, device=device)
            mem = L.MEM_DEVICE
        else:
            bl = np.empty(n_games, dtype=np.uint64)
            wh = np.empty(n_games, dtype=np.uint64)
            wi = np.empty(n_games, dtype=np.uint8)
            pl = np.empty(n_games, dtype=np.uint8)
            mem = L.MEM_HOST
    else:
        bl = wh = wi = pl = None
        mem = L.MEM_DEVICE
    L.check(L.lib().rvs_random_playouts(n_games, seed, first_game, rules, _p(bl), _p(wh), _p(wi), _p(pl),
                                        C.byref(total), mem,
                                        stream if stream is not None else L.current_stream()))
    return bl, wh, wi, pl, total.value


def perft(depth, black=0x0000000810000000, white=0x0000001008000000, side=1, rules=L.RULES_REF, stream=None):
    out = C.c_uint64(0)
    L.check(L.lib().rvs_perft(black, white, side, depth, rules, C.byref(out),
                              stream if stream is not None else L.current_stream()))
    return out.value


def encode_planes(black, white, side, layout=L.PLANES_F32_NCHW, rules=L.RULES_REF, stream=None):
    """ReversiGame.get_canonical_state for n positions (src/game/game.py:131-162)"""
    n = len(black)
    if layout == L.PLANES_F32_NCHW:
        out = _like(black, (n, 3, 8, 8), np.float32, "float32")
    else:
        out = _like(black, (n, 8, 8, 16), np.uint16, "bfloat16")
    mem = L.mem_of(black, white, side, out)
    L.check(L.lib().rvs_encode_planes(_u64(black), _u64(white), _u8(side), _p(out), n, layout, rules, mem,
                                      stream if stream is not None else L.current_stream()))
    return out
