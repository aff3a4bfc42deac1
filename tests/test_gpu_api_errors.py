"""GPU tier: error behaviour of the C ABI (include/rvs_b200.h): every misuse returns a negative code with a
message (raised as RvsError by the binding), never crashes, and leaves the handle usable."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def az():
    import alphazero_reversi_b200 as m
    return m


def test_engine_argument_errors(az):
    L = az._lib
    lib = L.lib()
    with pytest.raises(az.RvsError):
        az.Engine(0, 10, 1)                                  # no games
    with pytest.raises(az.RvsError):
        az.Engine(4, 70000, 1)                               # max_sims > 65535
    with pytest.raises(az.RvsError):
        az.Engine(4, 10, 1, rules=7)
    eng = az.Engine(8, 20, 4, evaluator=az.EVAL_E0)
    with pytest.raises(az.RvsError):
        eng.search(21, 1)                                    # num_sims above the pool size
    with pytest.raises(az.RvsError):
        eng.search(10, 5)                                    # wave above max_wave
    with pytest.raises(az.RvsError):
        eng.set_positions(np.zeros(9, np.uint64), np.zeros(9, np.uint64), np.ones(9, np.uint8))   # more than n_games
    with pytest.raises(az.RvsError):
        eng.select(2)                                        # select without begin_search
    with pytest.raises(az.RvsError):
        eng.set_lanes_per_game(3)
    with pytest.raises(az.RvsError):
        eng.set_root_noise(0.3, 1.5)
    with pytest.raises(az.RvsError):
        eng.set_root_noise(0.0, 0.25)                        # alpha out of range while noise is on
    with pytest.raises(az.RvsError):
        eng.predict(np.zeros(1, np.uint64), np.zeros(1, np.uint64), np.ones(1, np.uint8))        # no weights loaded
    with pytest.raises(az.RvsError):
        eng.load_weights(np.zeros(10, dtype=np.float32))     # wrong blob size for the configured network
    assert L.lib().rvs_last_error()                          # a message is available
    # null handle / null outputs through the raw ABI
    assert lib.rvs_engine_search(None, 10, 1, None) < 0
    assert lib.rvs_engine_stats_get(eng._h, None, None) < 0
    cnt = C.c_int64(0)
    # the handle still works after all of the above
    eng.search(20, 4)
    v = eng.root_visits()
    assert v.sum() > 0 and eng.stats()["overflow"] == 0
    # drain with too small a capacity reports -4 and keeps the samples
    eng2 = az.Engine(4, 10, 1, evaluator=az.EVAL_ROLLOUT, seed=1)
    eng2.selfplay(10, plies=4 * 64, temperature=1.0, recycle=False)
    n = eng2.stats()["samples"]
    assert n > 100
    st = np.empty((8, 3, 8, 8), np.float32); pi = np.empty((8, 65), np.float32); z = np.empty(8, np.float32)
    rc = lib.rvs_engine_drain_samples(eng2._h, st.ctypes.data, pi.ctypes.data, z.ctypes.data, 8, C.byref(cnt), L.MEM_HOST, None)
    assert rc == -4 and cnt.value == n
    s2, p2, z2 = eng2.drain_samples()
    assert len(s2) == n
    assert len(eng2.drain_samples()[0]) == 0                 # ring is empty afterwards
    eng.close(); eng2.close()
    eng.close()                                              # double close is harmless


def test_external_evaluator_protocol_errors(az):
    eng = az.Engine(2, 16, 4, evaluator=az.EVAL_EXTERNAL)
    with pytest.raises(az.RvsError):
        eng.search(16, 4)                                    # the external evaluator has no fused search
    with pytest.raises(az.RvsError):
        eng.selfplay(16, plies=4)
    eng.begin_search()
    eng.select(4)
    with pytest.raises(az.RvsError):
        eng.select(4)                                        # previous wave not processed
    planes, valid = eng.leaf_planes()
    eng.process(np.full((8, 65), 1 / 65, np.float32), np.zeros(8, np.float32))
    with pytest.raises(az.RvsError):
        eng.process(np.full((8, 65), 1 / 65, np.float32), np.zeros(8, np.float32))   # nothing selected
    eng.close()
