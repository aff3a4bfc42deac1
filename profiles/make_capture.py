#!/usr/bin/env python3
"""Turn the exports of tools/capture_profiles.sh (gpurun_out/cap_r2/: ncu raw / source CSVs, launch lists, bench
lines) into the committed round-2 evidence under profiles/: one text summary per kernel, the launch lists, and
profiles/traffic_r2.json -- the ncu figures bench.py reads for `roofline`, stamped with the commit and the hash of
the kernel sources they were captured on (bench.py: kernel_source_hash; a mismatch makes it emit traffic_stale).
usage: python profiles/make_capture.py [gpurun_out/cap_r2]"""
import collections
import csv
import gzip
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "cap_r2")
OUT = os.path.join(ROOT, "profiles")
sys.path.insert(0, ROOT)
from bench import kernel_source_hash  # noqa: E402

WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__average_warp_latency_per_inst_issued.ratio", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TO_US = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}


def raw_rows(name):
    rows = list(csv.reader(open(os.path.join(SRC, name + ".raw.csv"), errors="replace")))
    hdr, unit = rows[0], rows[1]
    return [({h: v for h, v in zip(hdr, r)}, {h: u for h, u in zip(hdr, unit)}) for r in rows[2:] if len(r) == len(hdr)]


def num(d, u, key, table=None):
    v = float(d[key].replace(",", ""))
    return v * table[u[key]] if table else v


def source_stats(name, kernel_regex=None):
    """opcode mix and stall samples per kernel from the cuda,sass source page"""
    per = collections.OrderedDict()
    cur = None
    hdr = None
    with gzip.open(os.path.join(SRC, name + ".src.csv.gz"), "rt", errors="replace") as f:
        for r in csv.reader(f):
            if not r:
                continue
            if r[0] == "Function Name":
                cur = per.setdefault(r[1], {"ops": collections.Counter(), "stalls": collections.Counter(), "lines": collections.Counter(),
                                            "line_inst": collections.Counter(), "inst": 0, "file": None})
                continue
            if r[0] == "File Path":
                curfile = os.path.basename(r[1])
                continue
            if r[0] == "Line No":
                hdr = r
                ix = {h: i for i, h in enumerate(hdr)}
                continue
            if cur is None or hdr is None or len(r) < len(hdr):
                continue
            if r[0] == "":  # a SASS row
                sass = r[3].split()
                if not sass or sass[0] in ("...", "-"):
                    continue
                try:
                    n = int(r[ix["Instructions Executed"]] or 0)
                except ValueError:
                    continue
                op = (sass[1] if sass[0].startswith("@") and len(sass) > 1 else sass[0]).split(".")[0]
                cur["ops"][op] += n
                cur["inst"] += n
                for h in hdr:
                    if h.startswith("stall_") and "(Not" not in h:
                        try:
                            cur["stalls"][h] += int(r[ix[h]] or 0)
                        except ValueError:
                            pass
            else:  # a source line with its aggregated samples
                try:
                    key = (curfile, int(r[0]), r[1].strip()[:96])
                    cur["lines"][key] += int(r[ix["# Samples"]] or 0)
                    cur["line_inst"][key] += int(r[ix["Instructions Executed"]] or 0)
                except ValueError:
                    pass
    return per


def sims_in_launch(name):
    """simulations executed by the captured group-kernel launch, from the capture itself: warp-level executions of
    `if (act) ++cx.sims;` (rvs_treeg.cuh simulate_one_g) x games per warp (32 / lanes per game, third template argument)"""
    src = source_stats(name)
    for fn, st in src.items():
        m = re.search(r"k1g_kernel<\(int\)\d+, \(int\)\d+, \(int\)(\d+)", fn) or re.search(r"k1g_kernel<\d+, \d+, (\d+)", fn)
        if not m:
            continue
        lpg = int(m.group(1))
        # lines of simulate_one_g that execute once per simulation; the compiler attributes ONE warp instruction to at
        # least one of them in every instantiation (and two to some), so the smallest positive count is the number of
        # warp-level simulations
        cands = [v for (fl, ln, txt), v in st["line_inst"].items()
                 if fl == "rvs_treeg.cuh" and v > 0 and any(a in txt for a in ("++cx.sims", "backup_path_g(cx, plen, v, act)",
                                                                              "expand_node_g(cx, node, lm, kUniformPrior, eval)"))]
        if cands:
            return min(cands) * (32 // lpg), lpg
    raise SystemExit(f"{name}: cannot find the simulation counter line")


def short(name):
    return re.sub(r"\(.*", "", name).replace("rvs::", "").replace("(anonymous namespace)::", "").replace("unnamed>::", "").replace("void ", "").strip()


def summarise(name, out_name, title, units=None, unit_name="unit", top_lines=14):
    rows = raw_rows(name)
    src = source_stats(name)
    res = []
    with open(os.path.join(OUT, out_name), "w") as f:
        f.write(f"{title}\n(ncu --set full --clock-control none --import-source on; exported on the GPU box by tools/capture_profiles.sh,\n summarised by profiles/make_capture.py; kernel sources hash {kernel_source_hash()})\n")
        for (d, u), (fn, st) in zip(rows, list(src.items()) + [(None, None)] * len(rows)):
            f.write(f"\n=== {short(d['Kernel Name'])}   grid {d['launch__grid_size']} x block {d['launch__block_size']}\n")
            for k in WANT:
                if k in d and d[k] not in ("", "n/a"):
                    f.write(f"  {k:66s} {d[k]:>16s} {u[k]}\n")
            inst = num(d, u, "smsp__inst_executed.sum")
            us = num(d, u, "gpu__time_duration.sum", TO_US)
            dram = num(d, u, "dram__bytes_read.sum", TO_BYTES) + num(d, u, "dram__bytes_write.sum", TO_BYTES)
            rec = {"kernel": short(d["Kernel Name"]), "us": us, "dram_bytes": dram, "warp_inst": inst,
                   "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                   "alu_pipe_pct": float(d.get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active") or 0),
                   "tensor_pipe_pct": float(d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") or 0)}
            f.write(f"  DRAM bytes (read + write)                                          {dram / 1e6:16.2f} MB = {dram / us / 1e3:.1f} GB/s\n")
            if units:
                f.write(f"  warp instructions per {unit_name:44s} {inst / units:16.1f}   ({units:.0f} {unit_name}s in this launch)\n")
                rec["warp_inst_per_unit"] = inst / units
                rec["dram_bytes_per_unit"] = dram / units
            res.append(rec)
        for fn, st in src.items():
            if not st["inst"]:
                continue
            f.write(f"\n--- {short(fn)}: SASS opcode mix (executed warp instructions)\n")
            for k, v in st["ops"].most_common(16):
                f.write(f"  {k:10s} {v:13d} {100 * v / st['inst']:5.1f}%\n")
            ts = sum(st["stalls"].values())
            f.write("--- warp stall samples\n")
            for k, v in st["stalls"].most_common(8):
                f.write(f"  {k:26s} {v:9d} {100 * v / max(ts, 1):5.1f}%\n")
            tl = sum(st["lines"].values())
            if tl:
                f.write("--- source lines with the most samples\n")
                for (fl, ln, txt), v in st["lines"].most_common(top_lines):
                    f.write(f"  {100 * v / tl:5.2f}%  {fl}:{ln}  {txt}\n")
    return res


def launches(csv_name, out_name, cmd):
    r = subprocess.run([sys.executable, os.path.join(OUT, "launch_summary.py"), os.path.join(SRC, csv_name), cmd], capture_output=True, text=True)
    open(os.path.join(OUT, out_name), "w").write(r.stdout)


def main():
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    n1, lpg1 = sims_in_launch("k1g")
    n16, lpg16 = sims_in_launch("k1g_16k")
    k1g = summarise("k1g", "ncu_selfplay_k1g_r2.txt", f"selfplay_k1g_kernel<REF,ROLLOUT,{lpg1} lanes/game>: 4096 games, one launch of {n1 // 409600} step(s) = {n1 // 100} game-plies x 100 simulations (bench.py --steps-per-launch 2; simulations counted from the capture's own source counters)",
                    units=n1, unit_name="simulation")[0]
    k16 = summarise("k1g_16k", "ncu_selfplay_k1g_16k_r2.txt", f"selfplay_k1g_kernel, 16384 games ({lpg16} lanes per game), launch of {n16 // 100} game-plies x 100 simulations (tools/probe_selfplay.py 16384 0 2)",
                    units=n16, unit_name="simulation")[0]
    t128 = summarise("tower128", "ncu_conv_tower_128_r2.txt", "conv_tower_kernel<128>: the whole 5x128 network up to the head planes (first layer + 10 tower layers) in ONE persistent launch on 4096 boards (tools/probe_net.py 5 128 4096 predict)")[0]
    c128 = summarise("conv128", "ncu_conv_tc2_128_r2.txt", "conv3x3_tc2_kernel<128,128> (the per-layer path, RVS_OPT_NET_TOWER = 0): two consecutive tower layers (plain, then with residual) of a 5x128 forward on 4096 boards (RVS_TOWER=0 tools/probe_net.py 5 128 4096 predict)")
    c256 = summarise("conv256", "ncu_conv_tc2s_256_r2.txt", "conv3x3_tc2s_kernel (256 filters, streamed weights): two consecutive tower layers of a 20x256 forward on 4096 boards (tools/probe_net.py 20 256 4096 predict)")
    aux = summarise("nnaux", "ncu_nn_aux_r2.txt", "the small kernels of the 5x128 NN search on 4096 games, two half-batches of 2048 (tools/probe_nn_wave.py): heads, tree step")
    brd = summarise("board", "ncu_board_k1_r2.txt", "K1 streaming kernels on 4 Mi arbitrary disc sets (tools/probe_board.py 4194304)")
    launches("launches_bench_r2.csv", "launches_bench_r2.txt", "python bench.py --steps 2 --warmup 1 --min-seconds 0 --no-cpu --no-big")
    launches("launches_nnwave_r2.csv", "launches_nnwave_r2.txt", "four waves of the 5x128 NN search, 4096 games as two pipelined half-batches (tools/probe_nn_wave.py, --launch-skip 800 -c 24)")
    for f in ("bench_r2_1gpu.json", "bench_r2_reference_arm.json"):
        if os.path.exists(os.path.join(SRC, f)):
            shutil.copy(os.path.join(SRC, f), os.path.join(OUT, f))
    boards = 4096
    cap = {
        "note": "ncu --set full --clock-control none; per-launch figures of the dominant kernels (profiles/ncu_*_r2.txt); bench.py reads "
                "selfplay_k1g_kernel for `roofline` and reports traffic_stale when kernel_source_hash differs from the sources it runs",
        "commit": commit, "kernel_source_hash": kernel_source_hash(),
        "selfplay_k1g_kernel": {"sims_in_launch": n1, "steps_in_launch": n1 / 409600, "lanes_per_game": lpg1, "dram_bytes_per_step": k1g["dram_bytes"] / (n1 / 409600),
                                "warp_inst_per_sim": k1g["warp_inst_per_unit"], "issue_active_pct": k1g["issue_active_pct"],
                                "alu_pipe_active_pct": k1g["alu_pipe_pct"], "kernel_us_under_ncu": k1g["us"],
                                "source": "profiles/ncu_selfplay_k1g_r2.txt"},
        "selfplay_k1g_kernel_16384": {"sims_in_launch": n16, "lanes_per_game": lpg16, "warp_inst_per_sim": k16["warp_inst_per_unit"],
                                      "issue_active_pct": k16["issue_active_pct"], "alu_pipe_active_pct": k16["alu_pipe_pct"],
                                      "dram_bytes_per_sim": k16["dram_bytes_per_unit"], "source": "profiles/ncu_selfplay_k1g_16k_r2.txt"},
        "conv_tower_kernel_128": {"dram_bytes_per_launch": t128["dram_bytes"], "us": t128["us"], "tensor_pipe_active_pct": t128["tensor_pipe_pct"], "boards": boards,
                                  "flop": boards * 2 * (64 * 9 * 16 * 128 + 10 * 64 * 9 * 128 * 128),
                                  "algorithmic_bytes": boards * 64 * (64 * 2 + 192 * 4 / 64) + 11 * 9 * 128 * 128 * 2,
                                  "note": "algorithmic bytes = input tiles in + head planes out + weights once: inter-layer activations need not leave the chip "
                                          "(they go through L2: depth-first tile groups); flop counts the first layer at the K = 16 it runs"},
        "conv3x3_tc2_kernel_128": [{"dram_bytes_per_launch": c["dram_bytes"], "us": c["us"], "tensor_pipe_active_pct": c["tensor_pipe_pct"],
                                    "boards": boards, "algorithmic_bytes": boards * 64 * 128 * 2 * (2 + i) + 9 * 128 * 128 * 2,
                                    "layer": "plain" if i == 0 else "with residual"} for i, c in enumerate(c128)],
        "conv3x3_tc2s_kernel_256": [{"dram_bytes_per_launch": c["dram_bytes"], "us": c["us"], "tensor_pipe_active_pct": c["tensor_pipe_pct"],
                                     "boards": boards, "algorithmic_bytes": boards * 64 * 256 * 2 * (2 + i) + 9 * 256 * 256 * 2,
                                     "layer": "plain" if i == 0 else "with residual"} for i, c in enumerate(c256)],
        "nn_wave_small_kernels": [{"kernel": c["kernel"], "us": c["us"], "issue_active_pct": c["issue_active_pct"]} for c in aux],
        "board_kernels": [{"kernel": c["kernel"], "us": c["us"], "alu_pipe_active_pct": c["alu_pipe_pct"], "dram_bytes": c["dram_bytes"]} for c in brd],
    }
    json.dump(cap, open(os.path.join(OUT, "traffic_r2.json"), "w"), indent=1)
    print(json.dumps(cap, indent=1)[:3000])


if __name__ == "__main__":
    main()
