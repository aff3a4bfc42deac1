"""Multi-GPU plumbing for sharded self-play: games are independent, so every rank
owns a contiguous block of game slots with its own trees, RNG streams and sample ring; the only
exchanges are a weight broadcast per generation and a replay-sample gather.  Both are plain
torch.distributed collectives (NCCL over NVLink on the GPU box, gloo in the CPU test tier) and
neither is inside the search loop.

The reference has no counterpart: its `num_parallel_games` is accepted and ignored
(src/self_play/self_play.py:30) and `torch.distributed` is imported but unused (src/mcts/mcts.py:11).
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_games: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous block of global game ids owned by `rank`: (first_game, count); sizes differ by <= 1"""
    base, rem = divmod(n_games, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def rank_seed(seed: int, rank: int) -> int:
    """per-rank engine seed so that shards draw disjoint RNG streams"""
    return (seed * 0x9E3779B97F4A7C15 + (rank + 1) * 0xD1B54A32D192ED03) & 0x7FFFFFFFFFFFFFFF


def broadcast_weights(flat: torch.Tensor, src: int = 0) -> torch.Tensor:
    """trainer rank -> every self-play rank; `flat` = network.pack_state_dict(...)[0] (in place)"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(flat, src=src)
    return flat


def gather_samples(states: torch.Tensor, pi: torch.Tensor, z: torch.Tensor, dst: int = 0):
    """variable-length (states [n,3,8,8], pi [n,65], z [n]) from every rank -> concatenated on `dst`
    (None elsewhere).  Counts are all-gathered first, payloads padded to the maximum count."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return states, pi, z
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = states.device
    n = torch.tensor([states.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    mx = max(max(counts), 1)
    # one packed row per sample: 192 planes + 65 pi + 1 z
    row = torch.zeros((mx, 258), dtype=torch.float32, device=dev)
    k = states.shape[0]
    if k:
        row[:k, :192] = states.reshape(k, 192)
        row[:k, 192:257] = pi
        row[:k, 257] = z
    bufs = [torch.empty_like(row) for _ in range(world)] if rank == dst else None
    dist.gather(row, bufs, dst=dst)
    if rank != dst:
        return None
    parts = [bufs[r][:counts[r]] for r in range(world)]
    allr = torch.cat(parts, dim=0)
    return allr[:, :192].reshape(-1, 3, 8, 8), allr[:, 192:257].contiguous(), allr[:, 257].contiguous()


PACKED_ROW_WORDS = 70  # int32 words per packed sample: 2+2 boards, 1 side|z, 65 pi bits = 280 bytes


def pack_rows(s):
    """PackedSamples (torch) -> [n, 70] int32 rows (one contiguous message per rank)"""
    k = len(s)
    row = torch.empty((k, PACKED_ROW_WORDS), dtype=torch.int32, device=s.side.device)
    if k:
        row[:, 0:2] = s.black.view(torch.int32).reshape(k, 2)
        row[:, 2:4] = s.white.view(torch.int32).reshape(k, 2)
        row[:, 4] = s.side.to(torch.int32) | ((s.z.to(torch.int32) & 0xFF) << 8)
        row[:, 5:70] = s.pi.view(torch.int32)
    return row


def unpack_rows(allr):
    from .replay import PackedSamples
    zz = ((allr[:, 4] >> 8) & 0xFF).to(torch.uint8).view(torch.int8)
    return PackedSamples(allr[:, 0:2].contiguous().view(torch.int64).reshape(-1), allr[:, 2:4].contiguous().view(torch.int64).reshape(-1),
                         (allr[:, 4] & 0xFF).to(torch.uint8), zz, allr[:, 5:70].contiguous().view(torch.float32))


def gather_packed(s, dst: int = 0, info: Optional[dict] = None):
    """replay.PackedSamples (torch, on this rank's device) from every rank -> concatenated on `dst`
    (None elsewhere).  One 280-byte row per sample (black, white, side, z, pi) instead of the 1032 bytes
    of the trainer format: 3.7x less NVLink traffic; planes are re-derived on `dst` by the K3 kernel.

    One all_gather_into_tensor of the counts (a single host read), then EXACT-size point-to-point
    transfers straight into the slices of one preallocated destination buffer: no max-padding, no
    per-rank .item() round trips, no concatenation copy.  `info` (optional dict) receives `rows`,
    `bytes_received` (on dst) and `bytes_sent`."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        if info is not None:
            info.update(rows=len(s), bytes_received=0, bytes_sent=0)
        return s
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = s.side.device
    n = torch.tensor([len(s)], dtype=torch.int64, device=dev)
    counts_t = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_t, n)
    counts = counts_t.cpu().tolist()  # the one host synchronisation of the gather
    row = pack_rows(s)
    total = int(sum(counts))
    if info is not None:
        info.update(rows=total if rank == dst else len(s),
                    bytes_received=(total - counts[dst]) * PACKED_ROW_WORDS * 4 if rank == dst else 0,
                    bytes_sent=0 if rank == dst else len(s) * PACKED_ROW_WORDS * 4)
    if rank != dst:
        if len(s):  # batched P2P: not serialised with the other collectives of the group
            for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, row, dst)]):
                w.wait()
        return None
    allr = torch.empty((total, PACKED_ROW_WORDS), dtype=torch.int32, device=dev)
    off = [0]
    for c in counts:
        off.append(off[-1] + int(c))
    ops = []
    for r in range(world):
        if counts[r] == 0:
            continue
        view = allr[off[r]:off[r + 1]]
        if r == dst:
            view.copy_(row)
        else:
            ops.append(dist.P2POp(dist.irecv, view, r))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return unpack_rows(allr)


def sharded_self_play(model, args: dict, total_games: int, dst: int = 0, slots_per_rank: int = 4096):
    """BASELINE config 5: `total_games` self-play games sharded over the ranks of the default process group
    (one process per GPU).  Rank `dst` is the trainer rank: its weights are broadcast first (NCCL), every rank
    then plays its contiguous block of games with its own RNG streams, trees and sample ring -- no collective
    inside the search loop -- and the packed samples are gathered on `dst`.

    model: RvsNetwork (weights taken from rank `dst`) or a built-in evaluator object (UniformRollout, ...).
    args:  the reference's SelfPlay args (`num_simulations`, `c_puct`, `temperature`, `batch_size`, `seed`, and the
           opt-in `apply_dirichlet_noise` / `dirichlet_alpha` / `dirichlet_epsilon`).
    Returns replay.PackedSamples on rank `dst` (every finished game of every rank, whole games only), None elsewhere.
    """
    from . import _lib as L
    from .engine import Engine
    from .replay import PackedSamples
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    dev = torch.device("cuda", torch.cuda.current_device())
    _, count = shard_range(total_games, rank, world)
    S = args.get("num_simulations", 800)
    K = max(1, args.get("batch_size", 1))
    T = args.get("temperature", 1.0)
    ev = model.evaluator
    if count == 0:  # more ranks than games: this rank only takes part in the collectives
        if ev == L.EVAL_NN:
            flat = model.flat.to(dev) if rank == dst else torch.empty_like(model.flat, device=dev)
            broadcast_weights(flat, src=dst)
        empty = PackedSamples(torch.empty(0, dtype=torch.int64, device=dev), torch.empty(0, dtype=torch.int64, device=dev),
                              torch.empty(0, dtype=torch.uint8, device=dev), torch.empty(0, dtype=torch.int8, device=dev),
                              torch.empty((0, 65), dtype=torch.float32, device=dev))
        return gather_packed(empty, dst=dst) if world > 1 else empty
    slots = max(1, min(slots_per_rank, count))
    eng = Engine(slots, S, K, evaluator=ev, c_puct=args.get("c_puct", 1.0), seed=rank_seed(args.get("seed", 0), rank),
                 device=dev.index, sample_capacity=64 * slots * 3, net_blocks=getattr(model, "net_blocks", 0),
                 net_filters=getattr(model, "net_filters", 0))
    if ev == L.EVAL_NN:
        flat = model.flat.to(dev) if rank == dst else torch.empty_like(model.flat, device=dev)
        broadcast_weights(flat, src=dst)
        eng.load_weights(flat)
    if args.get("apply_dirichlet_noise", False):
        eng.set_root_noise(args.get("dirichlet_alpha", 0.3), args.get("dirichlet_epsilon", 0.25))
    eng.set_option(L.OPT_GAME_LIMIT, count)  # exactly `count` games are started on this rank (no surplus searches)
    parts = []
    persistent = K == 1 and ev in (L.EVAL_E0, L.EVAL_ROLLOUT, L.EVAL_NN)
    finished = 0
    while finished < count:
        if persistent:
            eng.selfplay(S, plies=16 * slots, temperature=T, recycle=True)
        else:
            eng.search(S, K)
            eng.play(T, recycle=True)
        st = eng.stats()
        if st["overflow"] or st["stalled"] or st["samples_dropped"]:
            raise L.RvsError(f"engine error counters non-zero: {st}")
        if st["games_finished"] > finished:
            finished = st["games_finished"]
            parts.append(eng.drain_packed(device=dev))
    eng.close()
    mine = PackedSamples(*(torch.cat([getattr(p_, f) for p_ in parts]) for f in ("black", "white", "side", "z", "pi")))
    return gather_packed(mine, dst=dst) if world > 1 else mine
