"""BASELINE config 5 in small: sharded self-play over the ranks of a torchrun launch, weights broadcast from
rank 0, packed samples gathered on rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_sharded_selfplay.py [games] [sims] [rollout|nn]
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import alphazero_reversi_b200 as az
from alphazero_reversi_b200 import dist as azd

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank = dist.get_rank() if world > 1 else 0
games = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 50
kind = sys.argv[3] if len(sys.argv) > 3 else "rollout"
nb = int(sys.argv[4]) if len(sys.argv) > 4 else 2
nf = int(sys.argv[5]) if len(sys.argv) > 5 else 64
slots = int(sys.argv[6]) if len(sys.argv) > 6 else 4096
if kind == "nn":
    torch.manual_seed(42 + rank)        # ranks start from DIFFERENT weights: the broadcast must make them equal
    model = az.RvsNetwork.from_module(az.AlphaZeroNetwork(8, nb, nf).eval())
else:
    model = az.UniformRollout(seed=9)
t0 = time.perf_counter()
res = azd.sharded_self_play(model, {"num_simulations": sims, "batch_size": 1, "temperature": 1.0, "seed": 123}, games,
                            slots_per_rank=slots)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
if rank == 0:
    n = len(res)
    starts = int(((res.black == 0x0000000810000000) & (res.white == 0x0000001008000000) & (res.side == 1)).sum())
    assert starts >= games, (starts, games)
    assert n >= games * 50 and set(torch.unique(res.z).tolist()) <= {-1, 0, 1}
    assert torch.allclose(res.pi.sum(dim=1), torch.ones(n, device=res.pi.device), atol=1e-5)
    td = az.replay.to_training_data(res)
    assert td["states"].shape == (n, 3, 8, 8)
    print(f"sharded self-play ok: world {world}, {starts} games, {n} samples gathered on rank 0 in {dt:.2f} s ({kind} {nb}x{nf}, {sims} sims/move, "
          f"{slots} slots per rank) = {n * sims / dt / 1e6:.1f} M sims/s incl. engine setup, weight broadcast and gather")
else:
    assert res is None
if world > 1:
    dist.destroy_process_group()
