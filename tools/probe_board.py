"""probe: K1 / K3 streaming kernels on positions resident in HBM -- achieved GB/s against the measured copy peak.
usage: python tools/probe_board.py [n_positions]      (prints one JSON line)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import alphazero_reversi_b200 as az

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 25
dev = torch.device("cuda:0")
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
# mid-game positions: random playouts stopped nowhere are final positions; build mid-game ones by applying
# random legal moves to start positions on the device
bl, wh, wi, pl, _ = az.board_ops.random_playouts(n, 5, device=dev)          # final positions (legal mask mostly 0)
g = torch.Generator(device=dev).manual_seed(1)
occ = torch.randint(-2**63, 2**63 - 1, (n,), generator=g, dtype=torch.int64, device=dev)
pick = torch.randint(-2**63, 2**63 - 1, (n,), generator=g, dtype=torch.int64, device=dev)
bl, wh = occ & pick, occ & ~pick                                            # arbitrary disc sets: every code path runs
sd = torch.randint(1, 3, (n,), generator=g, device=dev).to(torch.uint8)
fl = torch.zeros(n, dtype=torch.uint8, device=dev)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

out = {}
lm = az.board_ops.legal_masks(bl, wh, sd)
ms = timed(lambda: az.board_ops.legal_masks(bl, wh, sd))
out["legal_masks"] = {"bytes_per_position": 25, "ms": ms, "gbs": n * 25 / ms / 1e6}
# first legal square of every position as the move
low = lm & -lm
mv = torch.where(lm != 0, (torch.log2((low.to(torch.float64)).abs().clamp(min=1)).round()).to(torch.uint8), torch.full_like(sd, 255))
mv = torch.where(low < 0, torch.full_like(mv, 63), mv)
ms = timed(lambda: az.board_ops.flip_masks(bl, wh, sd, mv))
out["flip_masks"] = {"bytes_per_position": 26, "ms": ms, "gbs": n * 26 / ms / 1e6}
b2, w2, s2, f2 = bl.clone(), wh.clone(), sd.clone(), fl.clone()
ms = timed(lambda: az.board_ops.apply_moves(b2, w2, s2, f2, mv))
out["apply_moves"] = {"bytes_per_position": 46, "ms": ms, "gbs": n * 46 / ms / 1e6, "board_steps_per_sec": n / ms * 1e3}
m = n // 8
ms = timed(lambda: az.board_ops.encode_planes(bl[:m], wh[:m], sd[:m]))
out["encode_planes_f32"] = {"bytes_per_position": 17 + 768, "ms": ms, "gbs": m * 785 / ms / 1e6}
for k in out:
    out[k]["frac_of_hbm_peak"] = out[k]["gbs"] / peak
out["n_positions"] = n
out["hbm_peak_gbs"] = peak
print(json.dumps(out))
