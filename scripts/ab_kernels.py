"""A/B builds of the kernels with different compile-time switches (cache policy of the streaming accesses, clamp
of the fp16 table index, ...) and their per-pass bandwidth.

  python scripts/ab_kernels.py build  name=-DFLAG=1,-DOTHER=2 ...   # here (nvcc cross-compiles): ab/lib_<name>.so
  python scripts/ab_kernels.py run [kernel_perf args]               # on the GPU box: scripts/kernel_perf.py per library

ab/ is git-ignored but travels to the GPU box with the snapshot."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "ldpc_decoder_b200", "csrc")
AB = os.path.join(ROOT, "ab")


def build(specs):
    os.makedirs(AB, exist_ok=True)
    procs = []
    for spec in specs:
        name, _, flags = spec.partition("=")
        out = os.path.join(AB, f"lib_{name}.so")
        cmd = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++14", "-lineinfo",
               "-Xcompiler", "-fPIC", "-shared", "-o", out, "ldpc_b200.cu", "ldpc_code.cpp"] + [f for f in flags.split(",") if f]
        procs.append((name, subprocess.Popen(cmd, cwd=CSRC)))
    for name, p in procs:
        if p.wait() != 0:
            raise SystemExit(f"build of {name} failed")
        print("built", name, flush=True)


def run(args):
    libs = sorted(f for f in os.listdir(AB) if f.startswith("lib_") and f.endswith(".so"))
    for lib in libs:
        env = dict(os.environ, LDPC_B200_LIB=os.path.join(AB, lib))
        print(f"### {lib}", flush=True)
        subprocess.call([sys.executable, os.path.join(ROOT, "scripts", "kernel_perf.py")] + args, env=env)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "build":
        build(sys.argv[2:])
    else:
        run(sys.argv[2:] if len(sys.argv) > 1 else [])
