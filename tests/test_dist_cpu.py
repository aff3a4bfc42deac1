"""CPU tier: the N>1 host logic (game sharding, weight broadcast, sample gather) on the gloo
backend with world_size 2 -- the same code path NCCL runs on the GPU box."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import alphazero_reversi_b200 as az
    from alphazero_reversi_b200 import dist as azd
    # weights: rank 0 holds the trainer's flat state dict
    torch.manual_seed(42)
    flat, nb, nf = az.pack_state_dict(az.AlphaZeroNetwork(8, 1, 64).state_dict())
    mine = flat.clone() if rank == 0 else torch.zeros_like(flat)
    azd.broadcast_weights(mine, src=0)
    ok_w = torch.equal(mine, flat)
    # samples: rank r contributes 3 + 2r samples with recognisable content
    n = 3 + 2 * rank
    st = torch.full((n, 3, 8, 8), float(rank + 1))
    pi = torch.full((n, 65), 1.0 / 65)
    z = torch.arange(n, dtype=torch.float32) + 100 * rank
    out = azd.gather_samples(st, pi, z, dst=0)
    # packed samples (replay.PackedSamples): bit patterns incl. the sign bit of the boards and z = -1
    gp = torch.Generator().manual_seed(5 + rank)
    pk = az.PackedSamples(torch.randint(-2**63, 2**63 - 1, (n,), generator=gp, dtype=torch.int64),
                          torch.randint(-2**63, 2**63 - 1, (n,), generator=gp, dtype=torch.int64),
                          torch.randint(1, 3, (n,), generator=gp).to(torch.uint8),
                          (torch.randint(0, 3, (n,), generator=gp) - 1).to(torch.int8), torch.rand((n, 65), generator=gp))
    gk = azd.gather_packed(pk, dst=0)
    packed_ok = None
    if gk is not None:
        parts = []
        for r in range(world):
            g2 = torch.Generator().manual_seed(5 + r)
            m = 3 + 2 * r
            parts.append((torch.randint(-2**63, 2**63 - 1, (m,), generator=g2, dtype=torch.int64),
                          torch.randint(-2**63, 2**63 - 1, (m,), generator=g2, dtype=torch.int64),
                          torch.randint(1, 3, (m,), generator=g2).to(torch.uint8),
                          (torch.randint(0, 3, (m,), generator=g2) - 1).to(torch.int8), torch.rand((m, 65), generator=g2)))
        exp = [torch.cat([p[i] for p in parts]) for i in range(5)]
        packed_ok = all(torch.equal(a, b) for a, b in zip((gk.black, gk.white, gk.side, gk.z, gk.pi), exp))
    # ragged case: rank 1 has nothing to contribute (more ranks than games); exact-size transfers, no padding
    info = {}
    empty = az.PackedSamples(torch.empty(0, dtype=torch.int64), torch.empty(0, dtype=torch.int64), torch.empty(0, dtype=torch.uint8),
                             torch.empty(0, dtype=torch.int8), torch.empty((0, 65)))
    ge = azd.gather_packed(pk if rank == 0 else empty, dst=0, info=info)
    if rank == 0:
        packed_ok = packed_ok and len(ge) == 3 and torch.equal(ge.black, pk.black) and info["bytes_received"] == 0
    info2 = {}
    ge2 = azd.gather_packed(empty if rank == 0 else pk, dst=0, info=info2)
    if rank == 0:
        packed_ok = packed_ok and len(ge2) == 5 and info2["bytes_received"] == 5 * 280
    else:
        assert ge2 is None and info2["bytes_sent"] == 5 * 280
    first, count = azd.shard_range(65536 + 1, rank, world)
    q.put((rank, ok_w, None if out is None else (out[0].shape[0], out[0][:, 0, 0, 0].tolist(), out[2].tolist()),
           first, count, azd.rank_seed(7, rank), packed_ok))
    dist.destroy_process_group()


def test_two_rank_gloo_broadcast_and_gather():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, w0, g0, f0, c0, s0, p0), (r1, w1, g1, f1, c1, s1, p1) = res
    assert p0 is True and p1 is None      # packed gather: bit-identical rows on rank 0 only
    assert w0 and w1                      # broadcast delivered identical weights
    assert g1 is None and g0[0] == 8      # 3 + 5 samples gathered on rank 0 only
    assert g0[1] == [1.0] * 3 + [2.0] * 5
    assert g0[2] == [0.0, 1.0, 2.0, 100.0, 101.0, 102.0, 103.0, 104.0]
    assert (f0, c0, f1, c1) == (0, 32769, 32769, 32768) and f1 == f0 + c0
    assert s0 != s1


def test_shard_range_covers_all_games():
    from alphazero_reversi_b200 import dist as azd
    for n, w in ((65536, 8), (10, 4), (3, 8), (4096, 1)):
        spans = [azd.shard_range(n, r, w) for r in range(w)]
        assert sum(c for _, c in spans) == n
        assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(w - 1))
