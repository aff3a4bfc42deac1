"""GPU tier: round-2 hardening of the C ABI and the mirrors -- per-handle network state under two host
threads, input validation, RVS_MEM_HOST_ASYNC normalisation, dtype checks of the binding, independent RNG
streams across set_positions calls, the game limit of the batched self-play mirror, the asynchronous drain and
the caller's-current-device contract."""
import ctypes as C
import threading

import numpy as np
import pytest
import torch

import orc
from stubs import perturb_bn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def az():
    import alphazero_reversi_b200 as m
    return m


def _net(az, seed, nb=2, nf=64):
    torch.manual_seed(seed)
    net = az.AlphaZeroNetwork(8, nb, nf)
    with torch.no_grad():
        perturb_bn(net, seed + 1)
    return net.eval()


def _positions(n, seed):
    bl, wh, wi, pl = orc.random_playouts(n, seed)
    rng = np.random.default_rng(seed)
    occ = rng.integers(0, 2**64, n, dtype=np.uint64)
    pick = rng.integers(0, 2**64, n, dtype=np.uint64)
    return occ & pick, occ & ~pick, rng.integers(1, 3, n).astype(np.uint8)


def test_two_threads_two_handles_two_networks(az):
    """the arena's case: two weight sets in one process, each driven by its own host thread.  The fused
    last-layer epilogue takes the folded head weights as a kernel parameter; they are per network now (a
    process-wide static copy raced here in round 1).  Each thread's outputs must equal its network's
    single-threaded outputs bit for bit, on every iteration."""
    n = 256
    nets = [az.RvsNetwork.from_module(_net(az, 42)), az.RvsNetwork.from_module(_net(az, 1234))]
    bl, wh, sd = _positions(n, 9)
    ref = []
    for rn in nets:
        e = az.Engine(n, 8, 1, evaluator=az.EVAL_NN, net_blocks=2, net_filters=64)
        rn.attach(e)
        ref.append(e.predict(bl, wh, sd, probs=True))
        e.close()
    assert np.abs(ref[0][0] - ref[1][0]).max() > 1e-3  # the two networks really differ
    engines = []
    for rn in nets:
        e = az.Engine(n, 8, 1, evaluator=az.EVAL_NN, net_blocks=2, net_filters=64)
        rn.attach(e)
        engines.append(e)
    errors = []
    barrier = threading.Barrier(2)

    def work(i):
        try:
            st = torch.cuda.Stream()
            barrier.wait()
            for it in range(40):
                p, v = engines[i].predict(bl, wh, sd, probs=True, stream=st.cuda_stream)
                if not (np.array_equal(p, ref[i][0]) and np.array_equal(v, ref[i][1])):
                    errors.append((i, it, float(np.abs(p - ref[i][0]).max())))
                    return
        except Exception as ex:  # noqa: BLE001
            errors.append((i, repr(ex)))

    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
    # and two NN searches in flight at once give what each gives alone
    rb, rw, rs = bl[:32], wh[:32], sd[:32]
    alone = []
    for i in range(2):
        e = az.Engine(32, 40, 1, evaluator=az.EVAL_NN, net_blocks=2, net_filters=64, seed=5)
        nets[i].attach(e)
        e.set_positions(*_legal_roots(32))
        e.search(40, 1)
        alone.append(e.root_visits())
        e.close()
    both = [None, None]

    def search(i):
        try:
            st = torch.cuda.Stream()
            e = az.Engine(32, 40, 1, evaluator=az.EVAL_NN, net_blocks=2, net_filters=64, seed=5)
            nets[i].attach(e)
            e.set_positions(*_legal_roots(32), stream=st.cuda_stream)
            barrier.wait()
            e.search(40, 1, stream=st.cuda_stream)
            both[i] = e.root_visits(stream=st.cuda_stream)
            e.close()
        except Exception as ex:  # noqa: BLE001
            errors.append((i, repr(ex)))

    ts = [threading.Thread(target=search, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
    assert np.array_equal(both[0], alone[0]) and np.array_equal(both[1], alone[1])
    for e in engines:
        e.close()


def _legal_roots(n, seed=17):
    import ctypes as CC
    rng = np.random.default_rng(seed)
    L = orc.lib()
    out = []
    for i in range(n):
        b = orc.make_board(*orc.START)
        for _ in range(int(rng.integers(0, 40))):
            lm = L.orc_board_legal(CC.byref(b), 0)
            if not lm:
                break
            bits = [q for q in range(64) if (lm >> q) & 1]
            L.orc_apply(CC.byref(b), bits[int(rng.integers(0, len(bits)))], 0)
        out.append((b.black, b.white, b.side))
    return (np.array([r[0] for r in out], dtype=np.uint64), np.array([r[1] for r in out], dtype=np.uint64),
            np.array([r[2] for r in out], dtype=np.uint8))


def test_set_positions_validates_and_draws_fresh_streams(az):
    eng = az.Engine(16, 60, 1, evaluator=az.EVAL_ROLLOUT, seed=21)
    bl, wh, sd = _legal_roots(16)
    bad_side = sd.copy(); bad_side[3] = 0
    with pytest.raises(az.RvsError, match="position 3"):
        eng.set_positions(bl, wh, bad_side)
    overlap = wh.copy(); overlap[5] |= bl[5]
    with pytest.raises(az.RvsError, match="overlap"):
        eng.set_positions(bl, overlap, sd)
    # device inputs are validated on the device: the bad slot is parked and counted, the others search
    tb = torch.from_numpy(bl.view(np.int64)).cuda(); tw = torch.from_numpy(overlap.view(np.int64)).cuda(); ts = torch.from_numpy(sd).cuda()
    eng.set_positions(tb, tw, ts)
    eng.search(60, 1)
    st = eng.stats()
    v = eng.root_visits()
    assert st["bad_positions"] == 1 and v[5].sum() == 0 and (np.delete(v, 5, axis=0).sum(axis=1) == 59).all()
    eng.close()
    # successive set_positions calls on one handle play game ids g, g + G, g + 2G, ...: the rollout streams of
    # two searches of the SAME roots differ (they were identical in round 1), and each equals the oracle's
    eng = az.Engine(16, 60, 1, evaluator=az.EVAL_ROLLOUT, seed=21)
    vs = []
    for epoch in range(3):
        eng.set_positions(bl, wh, sd)
        eng.search(60, 1)
        v = eng.root_visits()
        vs.append(v)
        for g in range(16):
            ov, *_ = orc.mcts_search((int(bl[g]), int(wh[g]), int(sd[g])), 60, 1, evaluator=1, seed=21, game_id=g + 16 * epoch)
            assert np.array_equal(v[g], ov), (epoch, g)
    assert not np.array_equal(vs[0], vs[1]) and not np.array_equal(vs[1], vs[2])
    eng.reset()  # back to epoch 0
    eng.set_positions(bl, wh, sd)
    eng.search(60, 1)
    assert np.array_equal(eng.root_visits(), vs[0])
    eng.close()


def test_binding_rejects_wrong_dtypes(az):
    bl, wh, sd = _legal_roots(4)
    with pytest.raises(TypeError):
        az.board_ops.legal_masks(bl, wh, sd.astype(np.int64))      # int64 side would be read as 8 uint8 values
    with pytest.raises(TypeError):
        az.board_ops.legal_masks(bl.astype(np.float64), wh, sd)
    eng = az.Engine(4, 10, 1)
    with pytest.raises(TypeError):
        eng.set_positions(bl, wh, sd.astype(np.int32))
    eng.close()


def test_host_async_means_host_everywhere(az):
    """RVS_MEM_HOST_ASYNC on entry points that do not implement the asynchronous mode behaves exactly like
    RVS_MEM_HOST (staged copies, outputs complete on return) instead of dereferencing host pointers on the device"""
    L = az._lib
    lib = L.lib()
    bl, wh, sd = _legal_roots(64)
    ref = az.board_ops.legal_masks(bl, wh, sd)
    out = np.zeros(64, dtype=np.uint64)
    assert lib.rvs_legal_masks(bl.ctypes.data, wh.ctypes.data, sd.ctypes.data, out.ctypes.data, 64, 0, L.MEM_HOST_ASYNC, None) == 0
    assert np.array_equal(out, ref)
    assert lib.rvs_legal_masks(bl.ctypes.data, wh.ctypes.data, sd.ctypes.data, out.ctypes.data, 64, 0, 7, None) < 0
    eng = az.Engine(64, 10, 2, evaluator=az.EVAL_EXTERNAL)
    eng.set_positions(bl, wh, sd)
    b2 = np.zeros(64, np.uint64); w2 = np.zeros(64, np.uint64); s2 = np.zeros(64, np.uint8); f2 = np.zeros(64, np.uint8)
    assert lib.rvs_engine_get_positions(eng._h, b2.ctypes.data, w2.ctypes.data, s2.ctypes.data, f2.ctypes.data, 64, L.MEM_HOST_ASYNC, None) == 0
    assert np.array_equal(b2, bl) and np.array_equal(s2, sd)
    eng.begin_search()
    eng.select(2)
    planes = np.zeros((128, 3, 8, 8), np.float32); valid = np.zeros(128, np.uint8)
    assert lib.rvs_engine_leaf_planes(eng._h, planes.ctypes.data, valid.ctypes.data, L.MEM_HOST_ASYNC, None) == 0
    p_ref, v_ref = eng.leaf_planes()
    assert np.array_equal(planes, p_ref) and np.array_equal(valid, v_ref)
    probs = np.full((128, 65), 1 / 65, np.float32); vals = np.zeros(128, np.float32)
    assert lib.rvs_engine_process(eng._h, probs.ctypes.data, vals.ctypes.data, L.MEM_HOST_ASYNC, None) == 0
    assert eng.stats()["nodes"] > 0
    eng.close()
    # predict: outputs are written (they were silently dropped in round 1)
    e = az.Engine(64, 8, 1, evaluator=az.EVAL_NN, net_blocks=2, net_filters=64)
    az.RvsNetwork.from_module(_net(az, 42)).attach(e)
    lg_ref, v_ref = e.predict(bl, wh, sd)
    lg = np.zeros((64, 65), np.float32); v = np.zeros(64, np.float32)
    assert lib.rvs_engine_predict(e._h, bl.ctypes.data, wh.ctypes.data, sd.ctypes.data, 64, lg.ctypes.data, v.ctypes.data, L.MEM_HOST_ASYNC, None) == 0
    assert np.array_equal(lg, lg_ref) and np.array_equal(v, v_ref)
    e.close()


def test_selfplay_mirror_returns_exactly_the_games_started(az):
    """SelfPlay(num_parallel_games > 1).generate_games(n): exactly n games are STARTED (ids 0..n-1, a slot restarts
    only below the game limit), so the result is the full set of those games -- no surplus searches, no bias
    toward games that finish early -- and equals the oracle's games 0..n-1 as a set"""
    S, n, slots = 24, 21, 8
    sp = az.SelfPlay(az.UniformRollout(seed=31), {"num_simulations": S, "batch_size": 1, "temperature": 1.0,
                                                    "num_parallel_games": slots, "seed": 31})
    games = sp.generate_games(n)
    assert len(games) == n
    got = sorted(tuple(np.asarray(g["states"]).astype(np.int8).tobytes() for _ in (0,)) for g in games)
    exp = []
    for gid in range(n):
        samples, win = orc.self_play_game(S, 1, evaluator=1, seed=31, game_id=gid, temperature=1.0)
        planes = np.stack([orc.planes(int(s.black), int(s.white), int(s.side)) for s in samples])
        exp.append((planes.astype(np.int8).tobytes(),))
    assert got == sorted(exp)
    # engine level: with a limit the slots park instead of recycling, and nothing beyond the limit is searched
    eng = az.Engine(slots, S, 1, evaluator=az.EVAL_ROLLOUT, seed=31)
    with pytest.raises(az.RvsError):
        eng.set_option(az._lib.OPT_GAME_LIMIT, 3)  # below n_games
    eng.set_option(az._lib.OPT_GAME_LIMIT, n)
    for _ in range(40):
        eng.selfplay(S, plies=slots * 16, temperature=1.0, recycle=True)
    st = eng.stats()
    assert st["games_finished"] == n
    eng.close()


def test_drain_packed_async_matches_sync(az):
    outs = []
    for mode in ("sync", "async", "async_small"):
        eng = az.Engine(32, 20, 1, evaluator=az.EVAL_ROLLOUT, seed=3)
        eng.selfplay(20, plies=32 * 70, temperature=1.0, recycle=False)
        if mode == "sync":
            outs.append(eng.drain_packed().numpy())
        elif mode == "async":
            cnt = torch.zeros(1, dtype=torch.int64).pin_memory()
            pk, cnt = eng.drain_packed_async(32 * 64, "cuda:0", count_out=cnt)
            torch.cuda.synchronize()
            k = int(cnt[0])
            outs.append(az.PackedSamples(pk.black[:k], pk.white[:k], pk.side[:k], pk.z[:k], pk.pi[:k]).numpy())
            assert len(eng.drain_packed()) == 0
        else:  # capacity below the number pending: the oldest `cap` come out, the rest stay in order
            parts = []
            for _ in range(6):
                pk, cnt = eng.drain_packed_async(500, "cuda:0")
                k = int(cnt.cpu()[0])
                parts.append(az.PackedSamples(pk.black[:k], pk.white[:k], pk.side[:k], pk.z[:k], pk.pi[:k]).numpy())
            outs.append(az.PackedSamples.concat(parts))
        eng.close()
    # The three engines play the same games sample for sample, but a finished game claims its ring rows with an
    # atomic, so the ORDER of the games in the ring follows completion time and differs from run to run: compare
    # the samples as a multiset (rows sorted by their bytes).
    def canon(p):
        rows = np.concatenate([p.black.view(np.uint8).reshape(len(p), 8), p.white.view(np.uint8).reshape(len(p), 8),
                               p.side.reshape(len(p), 1).view(np.uint8), p.z.reshape(len(p), 1).view(np.uint8),
                               np.ascontiguousarray(p.pi).view(np.uint8).reshape(len(p), -1)], axis=1)
        return rows[np.lexsort(rows.T[::-1])]

    a = outs[0]
    assert len(a) > 32 * 50
    for b in outs[1:]:
        assert len(b) == len(a)
        assert np.array_equal(canon(a), canon(b))


def test_calls_leave_the_current_device_alone(az):
    if torch.cuda.device_count() < 2:
        # one GPU: still check that the default engine device is the current device and that calls keep it
        torch.cuda.set_device(0)
        e = az.Engine(4, 10, 1)
        assert e.device == torch.cuda.current_device()
        e.search(10, 1)
        assert torch.cuda.current_device() == 0
        e.close()
        return
    torch.cuda.set_device(1)
    e1 = az.Engine(4, 10, 1)              # default device = the caller's current device
    assert e1.device == 1
    e0 = az.Engine(4, 10, 1, device=0)    # an engine elsewhere does not move the caller
    e0.search(10, 1)
    assert torch.cuda.current_device() == 1
    e1.search(10, 1)
    assert e1.root_visits().sum() > 0 and e0.root_visits().sum() > 0
    with pytest.raises(ValueError):
        e0.set_positions(torch.zeros(4, dtype=torch.int64, device="cuda:1"), torch.zeros(4, dtype=torch.int64, device="cuda:1"),
                         torch.ones(4, dtype=torch.uint8, device="cuda:1"))
    e0.close(); e1.close()
    torch.cuda.set_device(0)
