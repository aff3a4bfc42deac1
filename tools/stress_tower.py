"""stress: whole-network kernel vs the chain of per-layer launches on random batch sizes and networks (bit-identical
outputs expected every time: the cross-layer protocol of conv_tower_kernel -- published store counters, in-place weight
swaps -- has no other checker, compute-sanitizer is not available on the pool).  usage: python tools/stress_tower.py [iters]"""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import alphazero_reversi_b200 as az
from stubs import perturb_bn
L = az._lib
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(7)
bad = 0
for it in range(iters):
    nb = int(rng.integers(1, 9))
    boards = int(rng.choice([rng.integers(1, 600), rng.integers(600, 3000), rng.integers(3000, 12000)]))
    torch.manual_seed(it)
    net = az.AlphaZeroNetwork(8, nb, 128)
    with torch.no_grad():
        perturb_bn(net, it)
    net.eval()
    occ = rng.integers(0, 2**63, boards, dtype=np.int64); pick = rng.integers(0, 2**63, boards, dtype=np.int64)
    bl, wh = (occ & pick).astype(np.uint64), (occ & ~pick).astype(np.uint64)
    sd = rng.integers(1, 3, boards).astype(np.uint8)
    eng = az.Engine(boards, 8, 1, evaluator=az.EVAL_NN, net_blocks=nb, net_filters=128)
    az.RvsNetwork.from_module(net).attach(eng)
    eng.set_option(L.OPT_NET_TOWER, 0)
    ref = eng.predict(bl, wh, sd)
    eng.set_option(L.OPT_NET_TOWER, 1)
    for rep in range(4):
        out = eng.predict(bl, wh, sd)
        if not (np.array_equal(out[0], ref[0]) and np.array_equal(out[1], ref[1])):
            bad += 1
            print(f"MISMATCH it {it} blocks {nb} boards {boards} rep {rep}: max |dlogit| {np.abs(out[0] - ref[0]).max()}", flush=True)
    eng.close()
print(f"{iters} configurations x 4 repetitions, mismatches: {bad}")
sys.exit(1 if bad else 0)
