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
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-o", so,
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
    L.hc_flips_carry.restype = C.c_uint64
    L.hc_flips_carry.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_int]
    L.hc_nth_set_bit.argtypes = [C.c_uint64, C.c_int]
    L.hc_stream_seed.restype = C.c_uint64
    L.hc_stream_seed.argtypes = [C.c_uint64] * 3
    L.hc_det_log.restype = C.c_double
    L.hc_det_log.argtypes = [C.c_double]
    L.hc_det_exp.restype = C.c_double
    L.hc_det_exp.argtypes = [C.c_double]
    L.hc_dirichlet.argtypes = [C.c_double, C.c_int, C.c_uint64, C.POINTER(C.c_float)]
    L.hc_noise_mix.restype = C.c_float
    L.hc_noise_mix.argtypes = [C.c_float] * 3
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
def test_ray_table_carry_flips_match_oracle(hc, rules):
    """the 8-lane groups' flips (ray of the move square + carry ripple, rvs_board.cuh: flip_ray / flip_carry):
    every EMPTY square of every position -- legal moves and the phantom ones the search never plays -- including
    dense endgame boards where runs reach the 6-cell limit"""
    import random
    rng = random.Random(77 + rules)
    pos = list(rand_positions(6000, 57 + rules))
    for _ in range(3000):  # dense boards: 50-63 discs
        occ = orc.M64
        for _ in range(rng.randrange(1, 14)):
            occ &= ~(1 << rng.randrange(64))
        P = rng.getrandbits(64) & occ
        pos.append((P, occ & ~P))
    n = 0
    for P, O in pos:
        E = ~(P | O) & orc.M64
        while E:
            idx = (E & -E).bit_length() - 1
            E &= E - 1
            assert hc.hc_flips_carry(P, O, idx, rules) == orc.flips(P, O, idx, rules), (hex(P), hex(O), idx)
            n += 1
    assert n > 100000


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


# ---- Dirichlet root noise (csrc/rvs_noise.cuh): product formulas vs the oracle's restatement ----
def test_det_log_exp_match_oracle_and_libm(hc):
    import math
    L = orc.lib()
    rng = np.random.default_rng(5)
    xs = np.concatenate([rng.uniform(1e-300, 1.0, 2000), rng.uniform(1.0, 1e6, 500), 10.0 ** rng.uniform(-300, 300, 500),
                         [1.0, 2.0, 0.5, math.sqrt(2.0), 1.4142135623730951, 2.0 ** -53]])
    for x in xs:
        a, b = hc.hc_det_log(float(x)), L.orc_det_log(float(x))
        assert a == b, x                                   # bit exact between the two implementations
        assert abs(a - math.log(x)) <= 4e-16 * max(1.0, abs(math.log(x))) + 1e-15, (x, a, math.log(x))
    for x in np.concatenate([-rng.uniform(0, 700, 3000), [0.0, -1e-300, -689.9, -690.0, -800.0]]):
        a, b = hc.hc_det_exp(float(x)), L.orc_det_exp(float(x))
        assert a == b, x
        ref = math.exp(x) if x > -690.0 else 0.0
        assert abs(a - ref) <= 2e-13 * ref + 1e-300, (x, a, ref)  # single-constant ln2 reduction: ~1e-16 * |k|


@pytest.mark.parametrize("alpha", [0.03, 0.3, 1.0, 2.5])
def test_dirichlet_matches_oracle_and_distribution(hc, alpha):
    L = orc.lib()
    k = 10
    n = 3000
    a = np.zeros(k, dtype=np.float32)
    b = np.zeros(k, dtype=np.float32)
    acc = np.zeros((n, k))
    for i in range(n):
        st = L.orc_stream_seed(99, i, 0xD1000000 + 7)
        hc.hc_dirichlet(alpha, k, st, a.ctypes.data_as(C.POINTER(C.c_float)))
        L.orc_dirichlet(alpha, k, st, b.ctypes.data_as(C.POINTER(C.c_float)))
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), i
        acc[i] = a
    assert np.allclose(acc.sum(axis=1), 1.0, atol=1e-5) and (acc >= 0).all()
    # Dirichlet(alpha) moments: E = 1/k, Var = (k-1) / (k^2 (k alpha + 1))
    var = (k - 1) / (k * k * (k * alpha + 1))
    se = np.sqrt(var / n)
    assert np.abs(acc.mean(axis=0) - 1.0 / k).max() < 5 * se
    assert abs(acc.var(axis=0).mean() - var) < 0.15 * var
    # ragged sizes incl. a single child
    for kk in (1, 2, 33, 64):
        aa = np.zeros(kk, dtype=np.float32); bb = np.zeros(kk, dtype=np.float32)
        hc.hc_dirichlet(alpha, kk, 12345, aa.ctypes.data_as(C.POINTER(C.c_float)))
        L.orc_dirichlet(alpha, kk, 12345, bb.ctypes.data_as(C.POINTER(C.c_float)))
        assert np.array_equal(aa, bb) and abs(float(aa.sum()) - 1.0) < 1e-5
