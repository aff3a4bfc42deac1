"""Packed self-play samples: the replay-file, NCCL-gather and trainer-ingest format.

The reference keeps a finished game as a dict of Python lists -- `states` (3,8,8) f32 planes,
`action_probs` (65,) f64, `current_players`, `values` (src/self_play/self_play.py:72-77) -- pickles
one `torch.save` file per game (self_play.py:129-131) and later turns the lists into the three f32
arrays the training step consumes (src/trainer/pipeline.py:172-252, 272-311).  The engine keeps the
same information packed: the position BEFORE the move as black / white bitboards + side to move,
z as int8 and pi as f32[65] (277 B per sample instead of 1032 B); the canonical planes are a pure
function of (black, white, side) and are re-derived on the device by the K3 encode kernel.

    PackedSamples                  container (numpy on the host or torch on a device)
    save_replay / load_replay      one flat binary file per generation
    to_reference_games             -> the reference's list of game dicts (drop-in for generate_games)
    from_reference_games           <- reference game dicts / its per-game .pt files
    to_training_data               -> pipeline._prepare_training_data's dict, on the device
    training_batches               -> the (states, argmax(pi), z) batches of pipeline._train_epoch
"""
import struct
from dataclasses import dataclass
from typing import Dict, Iterator, List

import numpy as np

from . import board_ops

MAGIC = b"RVSRPL01"
START_BLACK, START_WHITE = 0x0000000810000000, 0x0000001008000000


def _is_torch(x):
    return hasattr(x, "is_cuda")


@dataclass
class PackedSamples:
    black: object   # [n] uint64 (numpy) / int64 bit pattern (torch)
    white: object
    side: object    # [n] uint8, 1 = BLACK, 2 = WHITE (the reference's current_players)
    z: object       # [n] int8 in {-1, 0, +1}, from the side to move's perspective
    pi: object      # [n, 65] float32

    def __len__(self):
        return int(self.side.shape[0])

    def numpy(self) -> "PackedSamples":
        if not _is_torch(self.side):
            return self
        return PackedSamples(self.black.cpu().numpy().view(np.uint64), self.white.cpu().numpy().view(np.uint64),
                             self.side.cpu().numpy(), self.z.cpu().numpy(), self.pi.cpu().numpy())

    def to(self, device) -> "PackedSamples":
        import torch
        if _is_torch(self.side):
            return PackedSamples(*(t.to(device) for t in (self.black, self.white, self.side, self.z, self.pi)))
        return PackedSamples(torch.from_numpy(self.black.view(np.int64)).to(device), torch.from_numpy(self.white.view(np.int64)).to(device),
                             torch.from_numpy(self.side).to(device), torch.from_numpy(self.z).to(device), torch.from_numpy(self.pi).to(device))

    def states(self):
        """canonical planes [n,3,8,8] f32 (ReversiGame.get_canonical_state, src/game/game.py:131-162),
        computed by the K3 kernel where the samples live"""
        return board_ops.encode_planes(self.black, self.white, self.side)

    @staticmethod
    def concat(parts: List["PackedSamples"]) -> "PackedSamples":
        parts = [p.numpy() for p in parts]
        return PackedSamples(*(np.concatenate([getattr(p, f) for p in parts]) for f in ("black", "white", "side", "z", "pi")))


def save_replay(path: str, s: PackedSamples) -> None:
    """header: magic, n (int64 LE); then black u64[n], white u64[n], side u8[n], z i8[n], pi f32[n,65]"""
    s = s.numpy()
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<q", len(s)))
        for a, dt in ((s.black, "<u8"), (s.white, "<u8"), (s.side, "u1"), (s.z, "i1"), (s.pi, "<f4")):
            f.write(np.ascontiguousarray(a).astype(dt, copy=False).tobytes())


def load_replay(path: str) -> PackedSamples:
    with open(path, "rb") as f:
        head = f.read(16)
        if len(head) != 16 or head[:8] != MAGIC:
            raise ValueError(f"{path}: not a packed replay file")
        n = struct.unpack("<q", head[8:])[0]
        if n < 0:
            raise ValueError(f"{path}: negative sample count")
        def rd(dt, count):
            b = f.read(np.dtype(dt).itemsize * count)
            if len(b) != np.dtype(dt).itemsize * count:
                raise ValueError(f"{path}: truncated")
            return np.frombuffer(b, dtype=dt).copy()
        return PackedSamples(rd("<u8", n), rd("<u8", n), rd("u1", n), rd("i1", n), rd("<f4", n * 65).reshape(n, 65))


def to_reference_games(s: PackedSamples) -> List[Dict]:
    """the list `SelfPlay.generate_games` returns (self_play.py:72-77, 133): the ring stores each
    finished game's plies contiguously, ply 0 (the start position) first"""
    s = s.numpy()
    st = np.asarray(s.states()) if len(s) else np.zeros((0, 3, 8, 8), dtype=np.float32)
    games, cur = [], None
    for i in range(len(s)):
        if cur is None or (int(s.black[i]) == START_BLACK and int(s.white[i]) == START_WHITE and int(s.side[i]) == 1):
            cur = {"states": [], "action_probs": [], "current_players": [], "values": []}
            games.append(cur)
        cur["states"].append(st[i])
        cur["action_probs"].append(s.pi[i].astype(np.float64))
        cur["current_players"].append(int(s.side[i]))
        cur["values"].append(float(s.z[i]))
    return games


def from_reference_games(games: List[Dict]) -> PackedSamples:
    """reference game dicts (or the dicts its per-game `torch.save` files hold) -> packed samples"""
    w = (np.uint64(1) << np.arange(64, dtype=np.uint64))
    bl, wh, sd, z, pi = [], [], [], [], []
    for g in games:
        n = min(len(g.get("states", [])), len(g.get("action_probs", [])), len(g.get("values", [])))  # pipeline.py:189
        for i in range(n):
            planes = np.asarray(g["states"][i], dtype=np.float32).reshape(3, 64)
            own = int(((planes[0] > 0.5).astype(np.uint64) * w).sum())
            opp = int(((planes[1] > 0.5).astype(np.uint64) * w).sum())
            player = int(g["current_players"][i]) if len(g.get("current_players", [])) > i else 1
            bl.append(own if player == 1 else opp)
            wh.append(opp if player == 1 else own)
            sd.append(player)
            z.append(int(round(float(g["values"][i]))))
            pi.append(np.asarray(g["action_probs"][i], dtype=np.float32))
    return PackedSamples(np.array(bl, dtype=np.uint64), np.array(wh, dtype=np.uint64), np.array(sd, dtype=np.uint8),
                         np.array(z, dtype=np.int8), np.array(pi, dtype=np.float32).reshape(-1, 65))


def to_training_data(s: PackedSamples) -> Dict:
    """what pipeline._prepare_training_data returns (pipeline.py:226-252): states [n,3,8,8] f32,
    policy_targets [n,65] f32, value_targets [n,1] f32 -- as torch tensors on the device the samples
    are on (no host round trip) or as numpy arrays for host samples"""
    st = s.states()
    if _is_torch(s.side):
        return {"states": st, "policy_targets": s.pi, "value_targets": s.z.to(st.dtype).reshape(-1, 1)}
    return {"states": np.asarray(st), "policy_targets": s.pi, "value_targets": s.z.astype(np.float32).reshape(-1, 1)}


def training_batches(s: PackedSamples, batch_size: int, shuffle: bool = True, generator=None) -> Iterator:
    """the batches pipeline._train_epoch feeds to the model (pipeline.py:276-311): shuffled
    (states f32, policy class = argmax(pi), value target [b]) tensors, built on the samples' device"""
    import torch
    if not _is_torch(s.side):
        s = s.to("cuda" if torch.cuda.is_available() else "cpu")
    n = len(s)
    order = torch.randperm(n, device=s.side.device, generator=generator) if shuffle else torch.arange(n, device=s.side.device)
    labels = s.pi.argmax(dim=1)  # hard labels (pipeline.py:308-311)
    for i in range(0, n, batch_size):
        idx = order[i:i + batch_size]
        part = PackedSamples(s.black[idx], s.white[idx], s.side[idx], s.z[idx], s.pi[idx])
        yield part.states(), labels[idx], part.z.to(torch.float32)
