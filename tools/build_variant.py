"""Builds an experimental variant of the library for A/B runs on the GPU box:
    python tools/build_variant.py <name> [-DFLAG ...]   ->  build/variants/librvs_<name>.so
(tools/ab_net.sh / ab_selfplay.sh swap it in for the product library; probes take RVS_LIB=<path>).
Only the translation units that see the flags are rebuilt per variant; build/ is not committed."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "alphazero-reversi_b200", "csrc")
SOURCES = ["rvs_board.cu", "rvs_engine.cu", "rvs_net.cu", "rvs_conv_tc.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def main():
    name, flags = sys.argv[1], sys.argv[2:]
    out = os.path.join(ROOT, "build", "variants")
    obj = os.path.join(out, "obj_" + name)
    os.makedirs(obj, exist_ok=True)
    base = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
            "-diag-suppress", "550"] + flags
    only = os.environ.get("RVS_ONLY")  # e.g. RVS_ONLY=rvs_conv_tc.cu: the other objects come from build/obj (product flags)
    objs, jobs = [], []
    for s in SOURCES:
        if only and s not in only.split(","):
            objs.append(os.path.join(ROOT, "build", "obj", s.replace(".cu", ".o")))
            continue
        o = os.path.join(obj, s.replace(".cu", ".o"))
        objs.append(o)
        jobs.append([NVCC] + base + ["-c", os.path.join(CSRC, s), "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            raise SystemExit("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)

    with ThreadPoolExecutor(max_workers=4) as ex:
        list(ex.map(run, jobs))
    lib = os.path.join(out, f"librvs_{name}.so")
    run([NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"])
    print(lib)


if __name__ == "__main__":
    main()
