#!/usr/bin/env python
"""bench.py -- decoded Mb/s of the B200 flood decoder on BASELINE.json's AWGN workload.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
  python bench.py --impl reference ...                      (CPU arm: the reference algorithm on host cores)

A step = one decode() call over `loading_factor * frames_in_flight` synthetic frames of the README-shaped
rate-0.5 code (BASELINE.json configs[2]: AWGN, fp16 messages, 2^12 frames in flight, loading factor 8).
`value` is measured with inputs, syndromes and results resident in HBM; `e2e` through the same C-ABI call with
HOST buffers (H2D of channel values and syndromes, D2H of packed decisions inside the timed region).
Frames come from a recycled pool of distinct synthetic frames (frame f -> column f % pool), see DESIGN.md.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

MBIT = float(1 << 20)  # the reference's Mbit (test_report.cpp:107-111)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    # workload knobs (defaults = BASELINE.json configs[2])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp32"])
    ap.add_argument("--log2-frames", type=int, default=12)
    ap.add_argument("--loading", type=int, default=8)
    ap.add_argument("--log2-n", type=int, default=20)
    ap.add_argument("--sigma", type=float, default=0.824,
                    help="AWGN sigma; the stand-in code's threshold is ~0.826 (the absent sample code: 0.95, run at 0.94)")
    ap.add_argument("--imax", type=int, default=120)
    ap.add_argument("--check-every", type=int, default=10)
    ap.add_argument("--pool", type=int, default=128, help="distinct synthetic frames (recycled)")
    ap.add_argument("--cpu-frames", type=int, default=32, help="frames of the bounded CPU sample")
    ap.add_argument("--early-retire", action="store_true",
                    help="EXTENSION: LDPC_B200_FLAG_EARLY_RETIRE (not the reference's end-of-input behaviour)")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # median over the samples under load (upper half of power draw)
        if sm:
            order = np.argsort(power)
            load = [sm[i] for i in order[len(order) // 2:]]
            return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons),
                    "samples": len(sm), "power_w_max": float(max(power))}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}


def make_workload(args):
    from ldpc_decoder_b200 import alist, harness

    gen = alist.make_readme_shaped_code(args.log2_n, seed=1)
    dt = np.float16 if args.precision == "fp16" else np.float32
    inp, syn, bits = harness.make_frames(gen, args.pool, "awgn", args.sigma, seed=1234, dtype=dt)
    prm = harness.channel_param("awgn", args.sigma)
    return gen, inp, syn, bits, prm


def cpu_sample(args, gen, n_threads=None):
    """The reference algorithm on host cores: CPU restatement of the reference's OpenCL path (oracle/, kind
    'port': fp32, OpenCL clamp 2e^-11, even-count refills; CN/VN passes vectorised over the frame slots with
    AVX2 + libmvec, OpenMP over all cores -- what a CPU OpenCL runtime does with the float2 kernels) on a bounded
    sample of the same code and noise."""
    from ldpc_decoder_b200 import alist, harness
    from oracle import binding as ob

    n_threads = n_threads or (os.cpu_count() or 1)
    n = max(2, args.cpu_frames & ~1)
    inp, syn, bits = harness.make_frames(gen, n, "awgn", args.sigma, seed=99, dtype=np.float32)
    ocode = ob.OracleCode(text=alist.alist_text(gen))
    log2p = int(np.log2(n))
    dec = ob.OracleDecoder(ocode, ob.FP32, log2p, pre_threshold=ob.PRE_THRESHOLD_OPENCL, n_threads=n_threads, fast=True)
    t = time.perf_counter()
    res, iters, rep = dec.decode(ob.AWGN, harness.channel_param("awgn", args.sigma), args.imax, args.check_every,
                                 inp, syn, simd_pairs=True)
    dt_s = time.perf_counter() - t
    return {
        "value": n * gen.n_vars / MBIT / dt_s, "unit": "Mb/s", "cores": n_threads, "kind": "port",
        "sample": f"{n} frames of the same code (N={gen.n_vars}) at sigma={args.sigma}, fp32, {1 << log2p} in flight, "
                  f"imax={args.imax}, avg_iter={rep['avg_iter']:.1f}, {dt_s:.1f}s; SIMD-over-frames port of the "
                  f"reference's OpenCL kernels (AVX2 + libmvec), {n_threads} OpenMP threads",
    }, dt_s


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from ldpc_decoder_b200 import alist

    gen = alist.make_readme_shaped_code(args.log2_n, seed=1)
    cpu_sample(args, gen)  # warm-up (thread pool, page faults) -- bounded, counts as W
    times, last = [], None
    for _ in range(max(1, args.steps)):
        last, dt_s = cpu_sample(args, gen)
        times.append(dt_s)
    n = max(2, args.cpu_frames & ~1)
    value = n * gen.n_vars / MBIT / (sum(times) / len(times))
    last["value"] = value
    line = {
        "impl": "reference", "metric": "decoded Mb/s", "value": value, "unit": "Mb/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, gen, None),
        "cpu_baseline": last,
        "e2e": {"value": value, "unit": "Mb/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, gen, P):
    return {
        "workload": "BASELINE.json configs[2]: rate-0.5 AWGN code (README-shaped stand-in, the sample alist is "
                    "absent from the reference tree), fp16 messages, 2^12 frames in flight, loading factor 8",
        "code": {"n_vars": gen.n_vars, "n_checks": gen.n_checks, "n_edges": gen.n_edges, "n_punctured": gen.n_erased},
        "sigma": args.sigma,
        "sigma_note": "the stand-in ensemble's threshold is ~0.826 (the absent sample code: 0.95, README run at 0.94); "
                      "0.824 gives converging frames with ~67 average iterations",
        "num_iter_max": args.imax, "check_every": args.check_every,
        "frames_in_flight": P, "loading_factor": args.loading, "frames_per_step": (P or 0) * args.loading,
        "precision": args.precision, "frame_pool": args.pool,
        "end_of_input": "early retire + slot compaction (extension)" if args.early_retire else "reference (all slots iterate until the last frame stops)",
        "l2": "working set (messages) >> L2 (126 MB): no flush needed between steps",
    }


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import ldpc_decoder_b200 as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    gen, inp, syn, bits, prm = make_workload(args)
    code = L.Code.from_generated(gen)
    prec = L.FP16 if args.precision == "fp16" else L.FP32
    dec = L.Decoder(code, prec, args.log2_frames, device=local_rank)
    P = dec.parallel_factor()
    n_frames = P * args.loading  # frames per step per GPU (weak scaling: every rank decodes its own queue shard)
    W = gen.n_vars // 32

    # resident inputs for the `value` leg
    d_in = torch.from_numpy(inp.view(np.int16) if args.precision == "fp16" else inp).cuda()
    d_syn = torch.from_numpy(syn.view(np.int32)).cuda()
    d_res = torch.zeros((n_frames, W), dtype=torch.int32, device="cuda")
    h_res = np.zeros((n_frames, W), dtype=np.uint32)
    torch.cuda.synchronize()

    base_flags = L.FLAG_EARLY_RETIRE if args.early_retire else 0

    def step_device(flags=0):
        return dec.decode(L.AWGN, prm, d_in, d_syn, n_frames=n_frames, num_iter_max=args.imax,
                          num_iter_check_parity=args.check_every, results=d_res, pool_frames=args.pool,
                          flags=flags | base_flags)

    def step_host():
        return dec.decode(L.AWGN, prm, inp, syn, n_frames=n_frames, num_iter_max=args.imax,
                          num_iter_check_parity=args.check_every, results=h_res, pool_frames=args.pool, flags=base_flags)

    for _ in range(args.warmup):
        step_device()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    reps = []
    for _ in range(args.steps):
        _, iters, rep = step_device(L.FLAG_TIME_KERNELS)
        reps.append(rep)
    torch.cuda.synchronize()
    elapsed = max_over_ranks(time.perf_counter() - t0)
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # the same region on the device clock: CUDA events recorded by decode() on its own stream (first kernel ->
    # last result of every step), max over ranks.  `value` uses the larger (host-bracketed) figure.
    elapsed_events = max_over_ranks(sum(r["device_s"] for r in reps))
    total_frames = sum_over_ranks(float(n_frames * args.steps))
    value = total_frames * gen.n_vars / MBIT / elapsed
    launches = sum(r["cn_launches"] + r["vn_launches"] + r["other_launches"] for r in reps)

    # sanity: the decode is real work -- decisions of converged frames satisfy their syndromes
    res_host = d_res.cpu().numpy().view(np.uint32)
    conv = iters < args.imax
    sel = np.nonzero(conv)[0][:64]
    ok = bool(np.array_equal(gen.syndrome_packed(res_host[sel]), syn[sel % args.pool])) if len(sel) else False
    err = np.unpackbits((res_host[sel] ^ bits[sel % args.pool]).view(np.uint8), axis=1).sum(axis=1) if len(sel) else np.zeros(0)

    # roofline of the dominant kernel pair (one flood iteration = CN launch + VN launch over all P slots)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    s = 2 if args.precision == "fp16" else 4
    E, N, M = gen.n_edges, gen.n_vars, gen.n_checks
    cn_bytes = (2 * E * s + 4 * ((M + 31) // 32)) * P  # per CN launch over all P slots
    vn_bytes = ((2 * E + N) * s) * P                     # per VN launch over all P slots
    cn_s = sum(r["cn_s"] for r in reps)
    vn_s = sum(r["vn_s"] for r in reps)
    # the first check-node pass of every check period also builds the start messages of the frames loaded at that
    # check (fused refill): it is timed apart (cn_refill_s) and reported as service time, not as a plain pass
    n_cn_all = sum(r["cn_launches"] for r in reps)
    n_cn = n_cn_all - sum(r["cn_refill_launches"] for r in reps)
    n_vn = sum(r["vn_launches"] for r in reps)
    # launches over fewer than P slots (early-retire tail) count for the slots they process
    slot_frac = sum(r["slot_iterations"] for r in reps) / float(n_cn_all * P)
    cn_gbs = cn_bytes * n_cn * slot_frac / cn_s / 1e9
    vn_gbs = vn_bytes * n_vn * slot_frac / vn_s / 1e9
    pair_gbs = (cn_bytes * n_cn + vn_bytes * n_vn) * slot_frac / (cn_s + vn_s) / 1e9
    dev_s = sum(r["device_s"] for r in reps)
    traffic = traffic_detail = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{args.precision}:{P}")
        if tr and args.log2_n == 20:
            traffic = tr["cn"] + tr["vn"]  # DRAM bytes of one CN + one VN launch (ncu dram__bytes_read+write)
            traffic_detail = {"cn": tr["cn"], "vn": tr["vn"], "algorithmic_pair": cn_bytes + vn_bytes,
                              "source": tr["source"]}
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": "cn_update_kernel + vn_update_kernel (one flood iteration)",
        "achieved": pair_gbs, "peak": peak, "unit": "GB/s", "frac": pair_gbs / peak, "peak_source": peak_kind,
        "traffic": traffic, "traffic_detail": traffic_detail,
        "cn": {"gbs": cn_gbs, "frac": cn_gbs / peak, "avg_ms": 1e3 * cn_s / n_cn, "bytes_per_launch": cn_bytes},
        "vn": {"gbs": vn_gbs, "frac": vn_gbs / peak, "avg_ms": 1e3 * vn_s / n_vn, "bytes_per_launch": vn_bytes},
        "slot_fraction": slot_frac,
        "kernel_share_of_step": (cn_s + vn_s) / dev_s if dev_s else None,
        "timed": {"cn_s": cn_s, "vn_s": vn_s, "device_s": dev_s, "n_cn": int(n_cn), "n_vn": int(n_vn),
                  "global_iters": [int(r["global_iters"]) for r in reps]},
        "service_s_per_step": {k: sum(r[k] for r in reps) / len(reps)
                               for k in ("check_s", "retire_s", "refill_s", "cn_refill_s")},
        "cn_refill_launches": int(n_cn_all - n_cn),
        "frames_refilled_per_step": sum(r["frames_refilled"] for r in reps) / len(reps),
    }

    # e2e: same call with host buffers
    e2e = None
    if not args.skip_e2e:
        step_host()  # warm-up of the host path
        barrier()
        t0 = time.perf_counter()
        hreps = []
        for _ in range(args.steps):
            _, _, rep = step_host()
            hreps.append(rep)
        torch.cuda.synchronize()
        e_elapsed = max_over_ranks(time.perf_counter() - t0)
        barrier()
        e2e = {"value": total_frames * gen.n_vars / MBIT / e_elapsed, "unit": "Mb/s",
               "h2d_bytes_per_step": int(hreps[-1]["h2d_bytes"]), "d2h_bytes_per_step": int(hreps[-1]["d2h_bytes"]),
               "ms_per_step": 1e3 * e_elapsed / args.steps}

    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        cpu, _ = cpu_sample(args, gen)

    if rank == 0:
        line = {
            "metric": "decoded Mb/s", "value": value, "unit": "Mb/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16" if args.precision == "fp16" else "f32",
            "data": "synthetic", "config": workload_config(args, gen, P),
            "timing": {"host_bracketed_s": elapsed, "cuda_event_s": elapsed_events,
                       "note": "value = frames / host_bracketed_s (barrier + synchronize on both sides, max over ranks); "
                               "cuda_event_s = sum over steps of the CUDA-event time on the decoder's stream, max over ranks"},
            "avg_iter": float(np.mean([r["avg_iter"] for r in reps])),
            "decoding_throughput_mbps": gen.n_vars / MBIT / (np.mean([r["avg_iter"] for r in reps]) * (cn_s / n_cn + vn_s / n_vn) / P),
            # the same throughput as a fraction of the HBM roofline: frames/s <= peak / (B_it * avg_iter)
            "roofline_mbps": peak * 1e9 / (dec.bytes_per_frame_iteration() * float(np.mean([r["avg_iter"] for r in reps]))) * world * gen.n_vars / MBIT,
            "frames_checked": int(len(sel)), "syndromes_satisfied": ok, "bit_errors_checked": int(err.sum()),
            "frac_converged": float(conv.mean()),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        line["value_over_roofline"] = line["value"] / line["roofline_mbps"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
