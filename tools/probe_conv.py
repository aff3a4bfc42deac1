"""probe (debug variant built with -DRVS_CONV_PROBE, tools/build_variant.py): where the warp roles of the 128-filter
tcgen05 convolution wait.  usage: python tools/probe_conv.py build/variants/librvs_probe.so [tower(0|1) blocks boards]
Cycle counts per CTA, averaged over the CTAs, of the persistent tower kernel (tower=1: all layers) or of the last
per-layer launch (tower=0)."""
import ctypes as C
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import alphazero_reversi_b200 as az

az._lib.LIB_PATH = sys.argv[1]
tower = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nb = int(sys.argv[3]) if len(sys.argv) > 3 else 5
B = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
torch.manual_seed(42)
rn = az.RvsNetwork.from_module(az.AlphaZeroNetwork(8, nb, 128).eval())
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
occ = rng.integers(0, 2**63, B, dtype=np.int64); pick = rng.integers(0, 2**63, B, dtype=np.int64)
bl = torch.from_numpy(occ & pick).to(dev); wh = torch.from_numpy(occ & ~pick).to(dev)
sd = torch.ones(B, dtype=torch.uint8, device=dev)
eng = az.Engine(B, 100, 1, evaluator=az.EVAL_NN, net_blocks=nb, net_filters=128)
eng.set_option(az._lib.OPT_NET_TOWER, tower)
rn.attach(eng)
for _ in range(3): eng.predict(bl, wh, sd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): eng.predict(bl, wh, sd)
e1.record(); torch.cuda.synchronize()
print(f"tower={tower} forward {e0.elapsed_time(e1) / 10:.3f} ms")
L = az._lib.lib()
buf = (C.c_longlong * (148 * 16))()
L.rvs_debug_conv_probe.argtypes = [C.c_void_p]
assert L.rvs_debug_conv_probe(buf) == 0
a = np.frombuffer(buf, dtype=np.int64).reshape(148, 16)
lead, peer = a[0::2], a[1::2]
print("tiles per CTA:", lead[:, 7].min(), "..", lead[:, 7].max())
names = [("MMA total", 4, lead), ("MMA wait weights", 1, lead), ("MMA wait acc free", 2, lead), ("MMA wait stage full", 3, lead),
         ("producer wait stage free (leader)", 0, lead), ("producer wait stage free (peer)", 0, peer)]
if tower:
    names += [("producer wait own stores (leader)", 5, lead), ("producer wait weights free (leader)", 6, lead),
              ("epilogue total (leader)", 11, lead), ("epilogue wait acc full (leader)", 10, lead), ("epilogue in proxy fences", 8, lead),
              ("epilogue proxy fences (count)", 9, lead), ("epilogue residual issue", 15, lead), ("epilogue TMEM loads + arrive", 12, lead),
              ("epilogue math + stores", 13, lead), ("epilogue publish", 14, lead)]
else:
    names += [("epilogue total (leader)", 6, lead), ("epilogue wait acc full (leader)", 5, lead)]
for name, col, rows in names:
    v = rows[:, col]
    print(f"  {name:36s} mean {v.mean():9.0f}  min {v.min():8d}  max {v.max():8d} cycles   per tile {v.mean() / max(1, lead[:, 7].mean()):7.1f}")

if tower:
    buf2 = (C.c_longlong * (148 * 48))()
    L.rvs_debug_conv_layers.argtypes = [C.c_void_p]
    assert L.rvs_debug_conv_layers(buf2) == 0
    t = np.frombuffer(buf2, dtype=np.int64).reshape(148, 48)[0::2]
    nl = 2 * nb + 2
    d = np.diff(t[:, :nl], axis=1)
    print("MMA issuer, cycles per layer (mean over CTA pairs):", " ".join(f"{x:.0f}" for x in d.mean(axis=0)))
