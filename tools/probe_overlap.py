"""probe: does splitting the NN-evaluated search into two half-batches on two streams pay?  (one 4096-game
engine vs two 2048-game engines driven concurrently)"""
import os, sys, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import alphazero_reversi_b200 as az
torch.manual_seed(42)
rn = az.RvsNetwork.from_module(az.AlphaZeroNetwork(8, 5, 128).eval())
def mk(G, seed):
    e = az.Engine(G, 100, 1, evaluator=az.EVAL_NN, seed=seed, net_blocks=5, net_filters=128)
    rn.attach(e)
    return e
one = mk(4096, 1)
one.search(100, 1); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); one.search(100, 1); e1.record(); torch.cuda.synchronize()
print(f"one engine 4096 games: {e0.elapsed_time(e1):.2f} ms")
one.close()
a, b = mk(2048, 2), mk(2048, 3)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
for e, s in ((a, sa), (b, sb)):
    e.search(100, 1, stream=s.cuda_stream)
torch.cuda.synchronize()
import time
def run(e, s):
    e.search(100, 1, stream=s.cuda_stream)
t0 = time.perf_counter()
ta = threading.Thread(target=run, args=(a, sa)); tb = threading.Thread(target=run, args=(b, sb))
ta.start(); tb.start(); ta.join(); tb.join()
torch.cuda.synchronize()
print(f"two engines 2 x 2048 games on two streams (two host threads): {(time.perf_counter() - t0) * 1e3:.2f} ms")
