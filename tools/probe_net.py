"""probe: K4 forward throughput (predict on B positions) and NN-evaluated search step time"""
import sys, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import alphazero_reversi_b200 as az

nb, nf = int(sys.argv[1]) if len(sys.argv) > 1 else 5, int(sys.argv[2]) if len(sys.argv) > 2 else 128
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
torch.manual_seed(42)
net = az.AlphaZeroNetwork(8, nb, nf).eval()
rn = az.RvsNetwork.from_module(net)
flops = 2 * (64 * 27 * nf + 2 * nb * 64 * 9 * nf * nf + 64 * nf * 2 + 128 * 65 + 64 * nf + 64 * 256 + 256)
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
occ = rng.integers(0, 2**63, B, dtype=np.int64); pick = rng.integers(0, 2**63, B, dtype=np.int64)
bl = torch.from_numpy(occ & pick).to(dev); wh = torch.from_numpy(occ & ~pick).to(dev)
sd = torch.ones(B, dtype=torch.uint8, device=dev)
eng = az.Engine(B, 100, 1, evaluator=az.EVAL_NN, net_blocks=nb, net_filters=nf)
rn.attach(eng)
import os
if os.environ.get('RVS_TOWER') is not None: eng.set_option(az._lib.OPT_NET_TOWER, int(os.environ['RVS_TOWER']))
for _ in range(3): eng.predict(bl, wh, sd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): eng.predict(bl, wh, sd)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(f"predict {nb}x{nf} B={B}: {ms:.3f} ms  {B/ms/1e3:.2f} M evals/s  {B*flops/ms/1e9:.1f} TFLOP/s")
if len(sys.argv) > 4 and sys.argv[4] == "predict":
    sys.exit(0)
# NN search: 100 sims wave 1 from the start position
eng.search(100, 1); torch.cuda.synchronize()
e0.record(); eng.search(100, 1); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"search 100 sims wave 1, {B} games: {ms:.2f} ms  {B*100/ms/1e3:.2f} M sims/s")
