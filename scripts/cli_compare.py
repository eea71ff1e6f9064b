"""Run the reference's own command-line program (oracle/_ref/ref_ldpc_decoder_cuda_*, built from /root/reference) and
the B200 CLI (ldpc_decoder_b200/ldpc_decoder_b200_cli) with the same flags on the same alist, and compare summaries."""
import os, re, subprocess, sys, tempfile, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ldpc_decoder_b200 import alist

KEYS = {
    "errors": r"Total # of errors:\s+(\S+)", "ber": r"Bit error rate \(BER\):\s+(\S+)",
    "max_err": r"Maximum # of errors / frame:\s+(\S+)", "fer_target": r"Frames with more than \d+ errors:\s+(\d+)",
    "fer_any": r"Frames with at least one error:\s+(\d+)", "iters": r"Max/min/average number of iterations per vector: (\S+)",
    "frames": r"# of frames decoded:\s+(\d+)", "mbits": r"Mbits processed:\s+(\S+)",
    "throughput": r"Throughput including transfers and finish: (\S+)", "decoding_throughput": r"Decoding throughput: (\S+)",
}


def summary(text):
    out = {k: (re.search(p, text).group(1) if re.search(p, text) else None) for k, p in KEYS.items()}
    m = re.search(r"\* Channel and code description\n\n(.*?)\n\n\n\* Test result", text, re.S)
    out["description"] = m.group(1) if m else None
    return out


def run(log2n, sigma, prec, extra=("-p", "8", "-m", "2", "-e", "15", "-i", "120"), path=None):
    if path is None:
        gen = alist.make_readme_shaped_code(log2n, seed=1)
        path = os.path.join(tempfile.mkdtemp(), f"standin_{log2n}.alist")
        alist.write_alist(gen, path)
    args = ["-f", path, "-c", "1", "-n", str(sigma)] + list(extra)
    ref_bin = os.path.join(ROOT, "oracle", "_ref", f"ref_ldpc_decoder_cuda_{prec}")
    new_bin = os.path.join(ROOT, "ldpc_decoder_b200", "ldpc_decoder_b200_cli")
    ref = subprocess.run([ref_bin] + args, capture_output=True, text=True, timeout=900).stdout
    new = subprocess.run([new_bin] + args + ["-x", "16" if prec == "fp16" else "32"], capture_output=True, text=True, timeout=900).stdout
    return summary(ref), summary(new), ref, new


def run_dropin(path, sigma, prec, extra=("-p", "8", "-m", "2", "-e", "15", "-i", "120")):
    """The reference's OWN main.cpp compiled unchanged against include/ldpc_decoder_gpu.h and linked with
    libldpc_b200.so (oracle/Makefile refdropin)."""
    args = ["-f", path, "-c", "1", "-n", str(sigma)] + list(extra)
    exe = os.path.join(ROOT, "oracle", "_ref", f"dropin_ldpc_decoder_{prec}")
    out = subprocess.run([exe] + args, capture_output=True, text=True, timeout=900).stdout
    return summary(out), out


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    res = {}
    gen = alist.make_readme_shaped_code(20, seed=1)
    path = os.path.join(tempfile.mkdtemp(), "standin_20.alist")
    alist.write_alist(gen, path)
    for prec in ("fp16", "fp32"):
        r, n, rt, nt = run(20, 0.82, prec, path=path)
        dsum, dt = run_dropin(path, 0.82, prec)
        res[prec] = {"reference_cli": r, "b200_cli": n, "reference_main_cpp_on_dropin_header": dsum}
        print(prec, "dropin   :", {k: v for k, v in dsum.items() if k != "description"})
        open(os.path.join(ROOT, "gpurun_out", f"r06_cli_{prec}_reference_main_on_dropin.txt"), "w").write(dt)
        print(prec, "reference:", {k: v for k, v in r.items() if k != "description"})
        print(prec, "b200     :", {k: v for k, v in n.items() if k != "description"})
        print(prec, "description identical:", r["description"] == n["description"])
        open(os.path.join(ROOT, "gpurun_out", f"r06_cli_{prec}_reference.txt"), "w").write(rt)
        open(os.path.join(ROOT, "gpurun_out", f"r06_cli_{prec}_b200.txt"), "w").write(nt)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "r06_cli_compare.json"), "w"), indent=1)
