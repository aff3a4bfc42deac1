"""CPU tier: the kernels' bit formulas (csrc/rvs_board.cuh, compiled for the host by g++ through
csrc/rvs_hostcheck.cpp -- a test artefact) against the oracle on random and adversarial inputs.
Catches formula bugs before any GPU time is spent; the GPU tier repeats this through the C-ABI."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

import orc

CSRC = os.path.join(orc.ROOT, "alphazero-reversi_b200", "csrc")


@pytest.fixture(scope="module")
def hc():
    out = os.path.join(orc.ROOT, "build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "rvs_hostcheck.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so,
                           os.path.join(CSRC, "rvs_hostcheck.cpp")])
    L = C.CDLL(so)
    L.hc_legal.restype = C.c_uint64
    L.hc_legal.argtypes = [C.c_uint64, C.c_uint64, C.c_int]
    L.hc_flips.restype = C.c_uint64
    L.hc_flips.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int]
    L.hc_try_move.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint8),
                              C.POINTER(C.c_uint8), C.c_int, C.c_int, C.POINTER(C.c_uint64)]
    L.hc_playout.argtypes = [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint8),
                             C.POINTER(C.c_uint8), C.c_uint64, C.c_int]
    L.hc_legal_sliced.restype = C.c_uint64
    L.hc_legal_sliced.argtypes = [C.c_uint64, C.c_uint64, C.c_int]
    L.hc_flips_sliced.restype = C.c_uint64
    L.hc_flips_sliced.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int]
    L.hc_nth_set_bit.argtypes = [C.c_uint64, C.c_int]
    L.hc_stream_seed.restype = C.c_uint64
    L.hc_stream_seed.argtypes = [C.c_uint64] * 3
    return L


def rand_positions(n, seed):
    rr = random.Random(seed)
    for _ in range(n):
        dens = rr.choice([0.05, 0.2, 0.4, 0.6, 0.8, 0.95, 1.0])
        occ = 0
        for b in range(64):
            if rr.random() < dens:
                occ |= 1 << b
        style = rr.random()
        if style < 0.2:    # long runs of one colour: exercises 6/7-long lines and row wrap
            P = occ & rr.choice([0xFF, 0xFF00, 0x8181818181818181, 0x0101010101010101, 1 << rr.randrange(64)])
        else:
            P = occ & rr.getrandbits(64)
        yield P, occ & ~P


@pytest.mark.parametrize("rules", [orc.RULES_REF, orc.RULES_STRICT])
def test_legal_and_flips_match_oracle(hc, rules):
    for P, O in rand_positions(20000, 1 + rules):
        lm = orc.legal(P, O, rules)
        assert hc.hc_legal(P, O, rules) == lm
        m = lm
        while m:
            idx = (m & -m).bit_length() - 1
            m &= m - 1
            assert hc.hc_flips(P, O, idx, rules) == orc.flips(P, O, idx, rules)


@pytest.mark.parametrize("rules", [orc.RULES_REF, orc.RULES_STRICT])
def test_flips_on_every_empty_square(hc, rules):
    # flips must agree even for squares outside the legal mask (rvs_flip_masks accepts any square)
    for P, O in rand_positions(3000, 77 + rules):
        E = ~(P | O) & orc.M64
        while E:
            idx = (E & -E).bit_length() - 1
            E &= E - 1
            assert hc.hc_flips(P, O, idx, rules) == orc.flips(P, O, idx, rules)


@pytest.mark.parametrize("rules", [orc.RULES_REF, orc.RULES_STRICT])
def test_direction_sliced_formulas_match_oracle(hc, rules):
    """the warp-cooperative path: per-direction parts (bit-reversed boards for right shifts) ORed"""
    for P, O in rand_positions(12000, 31 + rules):
        assert hc.hc_legal_sliced(P, O, rules) == orc.legal(P, O, rules)
        E = ~(P | O) & orc.M64
        k = 0
        while E and k < 12:
            idx = (E & -E).bit_length() - 1
            E &= E - 1
            k += 1
            assert hc.hc_flips_sliced(P, O, idx, rules) == orc.flips(P, O, idx, rules)
        m = orc.legal(P, O, rules)
        while m:
            idx = (m & -m).bit_length() - 1
            m &= m - 1
            assert hc.hc_flips_sliced(P, O, idx, rules) == orc.flips(P, O, idx, rules)


@pytest.mark.parametrize("rules", [orc.RULES_REF, orc.RULES_STRICT])
def test_playouts_match_oracle(hc, rules):
    L = orc.lib()
    for g in range(3000):
        st = L.orc_stream_seed(12345, g, 0)
        assert st == hc.hc_stream_seed(12345, g, 0)
        b = orc.make_board(*orc.START)
        n = L.orc_random_playout(C.byref(b), st, rules)
        bl, wh = C.c_uint64(orc.START[0]), C.c_uint64(orc.START[1])
        sd, fl = C.c_uint8(1), C.c_uint8(0)
        n2 = hc.hc_playout(C.byref(bl), C.byref(wh), C.byref(sd), C.byref(fl), st, rules)
        assert (n, b.black, b.white, b.over, b.winner) == (n2, bl.value, wh.value, fl.value & 1, (fl.value >> 1) & 3)


def test_try_move_contract(hc):
    bl, wh = C.c_uint64(orc.START[0]), C.c_uint64(orc.START[1])
    sd, fl, nl = C.c_uint8(1), C.c_uint8(0), C.c_uint64(0)
    assert hc.hc_try_move(C.byref(bl), C.byref(wh), C.byref(sd), C.byref(fl), 0, 0, C.byref(nl)) == 0
    assert (bl.value, wh.value, sd.value) == orc.START
    assert hc.hc_try_move(C.byref(bl), C.byref(wh), C.byref(sd), C.byref(fl), 19, 0, C.byref(nl)) == 1
    b = orc.make_board(*orc.START)
    orc.lib().orc_apply(C.byref(b), 19, 0)
    assert (bl.value, wh.value, sd.value) == (b.black, b.white, b.side)
    assert nl.value == orc.lib().orc_board_legal(C.byref(b), 0)
    fl = C.c_uint8(1)  # game over -> rejected (game.py:47-48)
    assert hc.hc_try_move(C.byref(bl), C.byref(wh), C.byref(sd), C.byref(fl), 18, 0, C.byref(nl)) == 0


def test_nth_set_bit(hc):
    rr = random.Random(3)
    for _ in range(2000):
        m = rr.getrandbits(64) | (1 << rr.randrange(64))
        bits = [i for i in range(64) if (m >> i) & 1]
        k = rr.randrange(len(bits))
        assert hc.hc_nth_set_bit(m, k) == bits[k]
