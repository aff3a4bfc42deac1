#!/usr/bin/env python3
"""SASS evidence for the tensor-core path: opcode counts per kernel of the shipped library and an excerpt around the
first tcgen05 MMA of each convolution kernel.   usage: python profiles/sass_summary.py > profiles/sass_tcgen05_r2.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "alphazero-reversi_b200", "librvs_b200.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "UCGABAR_ARV", "MEMBAR", "CCTL", "HMMA", "IMMA"]
arch = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
print("cuobjdump -sass alphazero-reversi_b200/librvs_b200.so  (built by alphazero-reversi_b200/build.py: -gencode arch=compute_100a,code=sm_100a)")
print("embedded cubins:", ", ".join(sorted(set(re.findall(r"sm_\d+a?", arch)))))
fn = None
count = collections.OrderedDict()
lines = collections.OrderedDict()
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        fn = fn.replace("rvs::", "").replace("(anonymous namespace)::", "")
        fn = re.sub(r"^void ", "", re.sub(r"\(.*", "", fn))
        while fn in count:
            fn += "'"
        count[fn] = collections.Counter()
        lines[fn] = []
        continue
    if fn is None:
        continue
    m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(@!?U?P\d\s+)?([A-Z][A-Z0-9_]*)", ln)
    if m:
        count[fn][m.group(2)] += 1
        count[fn]["*"] += 1
        lines[fn].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln).rstrip())
tot = collections.Counter()
print(f"\n{'kernel':64s} {'instr':>7s} " + " ".join(f"{o:>8s}" for o in OPS))
for fn, c in count.items():
    if any(c[o] for o in OPS[:7]):
        print(f"{fn[:64]:64s} {c['*']:7d} " + " ".join(f"{c[o]:8d}" for o in OPS))
    for o in OPS:
        tot[o] += c[o]
print(f"{'whole library (' + str(len(count)) + ' kernels)':64s} {sum(c['*'] for c in count.values()):7d} " + " ".join(f"{tot[o]:8d}" for o in OPS))
print("\nUTCHMMA = tcgen05.mma (.2CTA: cta_group::2), UTMALDG = cp.async.bulk.tensor (TMA load), LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,")
print("SYNCS = mbarrier ops.  MEMBAR / CCTL: only in the kernel prologues / epilogues (barrier.cluster at start and end) since round 2 --")
print("the per-tile barrier traffic uses CTA-scope semantics (rvs_conv_tc.cu: mbar_wait_cluster / mbar_arrive_leader).")
for fn, ls in lines.items():
    if "conv3x3_tc2" not in fn:
        continue
    idx = next((i for i, l in enumerate(ls) if "UTCHMMA" in l), None)
    if idx is None:
        continue
    print(f"\n--- {fn}: around the first tcgen05.mma")
    for l in ls[max(0, idx - 14): idx + 10]:
        print(l)
    tma = next((i for i, l in enumerate(ls) if "UTMALDG" in l), None)
    print("--- first TMA load")
    for l in ls[max(0, tma - 3): tma + 3]:
        print(l)
    ld = next((i for i, l in enumerate(ls) if "LDTM" in l), None)
    print("--- first tcgen05.ld")
    for l in ls[max(0, ld - 2): ld + 3]:
        print(l)
