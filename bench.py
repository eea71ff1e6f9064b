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
    ap.add_argument("--pool", type=int, default=4096,
                    help="distinct synthetic frames (recycled): 4096 fp16 frames = 8.6 GB, so the strided gather of the "
                         "host-fed leg walks an 8 KB row stride as it would with all-distinct frames")
    ap.add_argument("--small-pool", type=int, default=0,
                    help="also time the e2e leg on a pool of this many frames (round-1 bench: 128; 256-byte rows)")
    ap.add_argument("--e2e-steps", type=int, default=5, help="timed steps of the e2e leg (at most --steps)")
    ap.add_argument("--time-budget", type=float, default=760.0,
                    help="seconds this process may take in total (the driver's scaling run allows 870 s per N): legs "
                         "beside the headline and the e2e figure are skipped, and named in `skipped_legs`, when the "
                         "time left does not cover their estimated cost")
    ap.add_argument("--skip-extra", action="store_true",
                    help="skip the legs beside the headline: reference CUDA build, other BASELINE configs, shared queue")
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


def make_code(args):
    from ldpc_decoder_b200 import alist

    return alist.make_readme_shaped_code(args.log2_n, seed=1)


def synth_pool(dec, gen, L, channel, noise, n, fp16, start=0):
    """n distinct frames of the reference harness' generator (ChaCha8 streams keyed by the frame index,
    src/main.cpp:450-538), synthesized on the device in the decode API's layout."""
    import torch

    N, M = gen.n_vars, gen.n_checks
    d_in = torch.empty((N, n), dtype=torch.float16 if fp16 else torch.float32, device="cuda")
    d_syn = torch.empty((n, (M + 31) // 32), dtype=torch.int32, device="cuda")
    d_ref = torch.empty((n, N // 32), dtype=torch.int32, device="cuda")
    dec.synthesize_frames(channel, noise, start, n, d_in, d_syn, d_ref)
    return d_in, d_syn, d_ref


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
                      "0.824 gives converging frames with ~67 average iterations (the README run on the real code: 90.7); "
                      "decoded Mb/s scales with 1/avg_iter, see value_at_readme_avg_iter",
        "num_iter_max": args.imax, "check_every": args.check_every,
        "frames_in_flight": P, "loading_factor": args.loading, "frames_per_step": (P or 0) * args.loading,
        "precision": args.precision, "frame_pool": args.pool,
        "frames": "distinct frames per pool entry, generated on the device with the reference harness' ChaCha8 streams "
                  "(frame f of a step reads pool entry f % pool); host copy of the same pool for the e2e leg",
        "end_of_input": "early retire + slot compaction (extension)" if args.early_retire else "reference (all slots iterate until the last frame stops)",
        "l2": "working set (messages) >> L2 (126 MB): no flush needed between steps",
    }


def load_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback 6650 GB/s (B200_PROFILING.md)"


def roofline_block(reps, gen, P, s, peak, peak_kind, traffic_key=None):
    """Roofline of the dominant kernel pair (one flood iteration = CN launch + VN launch over all P slots):
    algorithmic bytes per launch / CUDA-event time of the launches inside the timed region."""
    E, N, M = gen.n_edges, gen.n_vars, gen.n_checks
    cn_bytes = (2 * E * s + 4 * ((M + 31) // 32)) * P  # per CN launch over all P slots
    vn_bytes = ((2 * E + N) * s) * P                     # per VN launch over all P slots
    cn_s = sum(r["cn_s"] for r in reps)
    vn_s = sum(r["vn_s"] for r in reps)
    # the first check-node pass of a check period that loads frames also builds their start messages (fused
    # refill): it is timed apart (cn_refill_s) and reported as service time, not as a plain pass
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
    if traffic_key:
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(traffic_key)
            if tr:
                traffic = tr["cn"] + tr["vn"]  # DRAM bytes of one CN + one VN launch (ncu dram__bytes_read+write)
                traffic_detail = {"cn": tr["cn"], "vn": tr["vn"], "algorithmic_pair": cn_bytes + vn_bytes,
                                  "source": tr["source"]}
        except Exception:
            pass
    return {
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


def other_configs_leg(L, peak, peak_kind):
    """BASELINE.json configs[1] and configs[4] (the headline line is configs[2]): one warm-up + one timed decode each,
    device-resident frames generated on the device, reference scheduling, each with its own roofline block; for
    configs[1] also the host-fed call with the three wire formats (float per bit as the reference, half, packed bits)."""
    import torch

    from ldpc_decoder_b200 import alist, harness

    cases = [
        ("configs[1]: rate-0.9 code on a BSC, fp32 messages, 2^10 frames in flight (stand-in: (3,32)-regular, N = 2^20, "
         "p = 0.004, loading factor 4)", lambda: alist.make_regular_code(1 << 20, 3, 32, seed=4), L.FP32, 10, 4, "bsc", 0.004, 100, 256),
        ("configs[4]: harness-generated regular (3,6) code, N = 2^22, AWGN sigma = 0.875 (threshold 0.881), up to 500 "
         "iterations, fp16 messages, 2^11 frames in flight, loading factor 4",
         lambda: alist.make_regular_code(1 << 22, 3, 6, seed=6), L.FP16, 11, 4, "awgn", 0.875, 500, 256),
        ("configs[4], fp32 messages, 2^10 frames in flight",
         lambda: alist.make_regular_code(1 << 22, 3, 6, seed=6), L.FP32, 10, 4, "awgn", 0.875, 500, 256),
    ]
    out = []
    for name, make, prec, log2p, load, channel, noise, imax, pool in cases:
        gen = make()
        fp16 = prec == L.FP16
        dec = L.Decoder(L.Code.from_generated(gen), prec, log2p)
        P = dec.parallel_factor()
        n = P * load
        N = gen.n_vars
        ch = L.BSC if channel == "bsc" else L.AWGN
        nz = float(np.float16(noise)) if (fp16 and channel == "awgn") else noise
        d_in, d_syn, d_ref = synth_pool(dec, gen, L, ch, nz, pool, fp16)
        d_res = torch.empty((n, N // 32), dtype=torch.int32, device="cuda")
        prm = harness.channel_param(channel, nz)
        kw = dict(n_frames=n, num_iter_max=imax, results=d_res, pool_frames=pool)
        dec.decode(ch, prm, d_in, d_syn, **kw)  # warm-up
        torch.cuda.synchronize()
        t = time.perf_counter()
        _, iters, rep = dec.decode(ch, prm, d_in, d_syn, flags=L.FLAG_TIME_KERNELS, **kw)
        torch.cuda.synchronize()
        el = time.perf_counter() - t
        conv = iters < imax
        idx = np.arange(n) % pool
        ok = bool((d_res.cpu().numpy()[conv] == d_ref.cpu().numpy()[idx[conv]]).all())
        row = {"workload": name, "frames_in_flight": P, "frames": n, "value": n * N / MBIT / el, "unit": "Mb/s",
               "ms": 1e3 * el, "avg_iter": rep["avg_iter"], "max_iter": rep["max_iter"],
               "frac_converged": float(conv.mean()), "converged_frames_equal_transmitted_words": ok,
               "code": {"n_vars": N, "n_checks": gen.n_checks, "n_edges": gen.n_edges},
               "roofline": roofline_block([rep], gen, P, 2 if fp16 else 4, peak, peak_kind)}
        if channel == "bsc":
            # host-fed call, three wire formats: bytes from the decoder's own counters
            inp = d_in.cpu().numpy()
            syn = d_syn.cpu().numpy().view(np.uint32)
            n_tx = N - gen.n_erased
            hard = (inp[:n_tx] > 0).T
            pad = np.zeros((hard.shape[0], (n_tx + 31) // 32 * 32), dtype=np.uint8)
            pad[:, :n_tx] = hard
            packed = np.packbits(pad, axis=1, bitorder="little").view(np.uint32)
            h_res = np.zeros((n, N // 32), dtype=np.uint32)
            wire = {}
            for label, arr, extra in (("float32 per bit (reference API)", inp, {}),
                                      ("float16 per bit, fp32 compute", inp.astype(np.float16), {"flags": L.FLAG_HALF_INPUT}),
                                      ("1 bit per bit (packed)", packed, {"layout": L.LAYOUT_PACKED_BITS})):
                dec.decode(ch, prm, arr, syn, n_frames=n, num_iter_max=imax, results=h_res, pool_frames=pool, **extra)
                t = time.perf_counter()
                _, it2, r2 = dec.decode(ch, prm, arr, syn, n_frames=n, num_iter_max=imax, results=h_res, pool_frames=pool, **extra)
                el2 = time.perf_counter() - t
                wire[label] = {"e2e": n * N / MBIT / el2, "unit": "Mb/s", "h2d_bytes_per_step": int(r2["h2d_bytes"]),
                               "d2h_bytes_per_step": int(r2["d2h_bytes"]), "iterations_identical": bool(np.array_equal(it2, iters))}
            row["host_fed_wire_formats"] = wire
        out.append(row)
        del dec, d_in, d_syn, d_ref, d_res
        torch.cuda.empty_cache()
    return out


def ref_cuda_build_leg(L, gen, args, sigma, prm, d_in, d_syn, pool):
    """The UNMODIFIED reference CUDA backend rebuilt for sm_100 (oracle/_ref, compiled from /root/reference by
    oracle/Makefile) next to this decoder on the same box, same host arrays, at the largest number of frames in flight
    the reference's 32-bit indices allow for this code (2^10).  A reported baseline, outside every timed region of
    the main arm; its decisions are also the parity check of this run."""
    import tempfile

    from ldpc_decoder_b200 import alist
    try:
        from oracle import ref_binding as rb
    except Exception as e:  # pragma: no cover
        return {"unavailable": f"oracle/ref_binding import failed: {e}"}
    prec = args.precision
    if not rb.available(prec):
        return {"unavailable": "oracle/_ref/libldpc_ref_%s.so not built (needs /root/reference at build time)" % prec}
    log2p, load = 10, 2
    if (gen.n_edges << log2p) >= 2 ** 32:
        return {"unavailable": "P * E >= 2^32: beyond the reference's 32-bit indices"}
    n = (1 << log2p) * load
    cols = np.arange(n) % pool
    inp = np.ascontiguousarray(d_in[:, : min(pool, n)].cpu().numpy()[:, cols[: n] % min(pool, n)])
    syn = d_syn[: min(pool, n)].cpu().numpy().view(np.uint32)[cols % min(pool, n)]
    path = os.path.join(tempfile.mkdtemp(), "bench_code.alist")
    alist.write_alist(gen, path)
    ref = rb.RefDecoder(path, "awgn", sigma, log2p, prec)
    ref.decode(inp[:, : 1 << log2p], syn[: 1 << log2p], args.imax, args.check_every)  # warm-up (one batch)
    t = time.perf_counter()
    rres, rrep = ref.decode(inp, syn, args.imax, args.check_every)
    ref_s = time.perf_counter() - t
    ref.close()
    dec = L.Decoder(L.Code(path=path), L.FP16 if prec == "fp16" else L.FP32, log2p)
    dec.decode(L.AWGN, prm, inp[:, : 1 << log2p], syn[: 1 << log2p], num_iter_max=args.imax, num_iter_check_parity=args.check_every)
    t = time.perf_counter()
    res, iters, rep = dec.decode(L.AWGN, prm, inp, syn, num_iter_max=args.imax, num_iter_check_parity=args.check_every)
    new_s = time.perf_counter() - t
    del dec
    conv = iters < args.imax
    N = gen.n_vars
    return {
        "what": "reference CUDA backend (flood.cu + ldpc_decoder_gpu.cu, unmodified, nvcc -arch=sm_100) vs this decoder: "
                "decode() with the same HOST arrays (transfers included), same B200",
        "frames": n, "frames_in_flight": 1 << log2p, "precision": prec,
        "reference_mbps": n * N / MBIT / ref_s, "b200_mbps": n * N / MBIT / new_s, "speedup": ref_s / new_s,
        "decisions_identical": bool(np.array_equal(res, rres)) if prec == "fp16" else None,
        "decisions_identical_on_converged_frames": bool(np.array_equal(res[conv], rres[conv])),
        "iteration_stats": {"reference": rrep, "b200": {k: rep[k] for k in ("max_iter", "min_iter", "avg_iter")}},
    }


def main():
    t_start = time.perf_counter()
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    skipped = []

    def time_left_for(leg, cost_s):
        """Optional legs run only when the budget still covers their estimated cost (every rank decides alike: the
        estimate is compared with rank 0's clock)."""
        left = args.time_budget - (time.perf_counter() - t_start)
        ok = left > cost_s
        if world > 1:
            t = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device="cuda")
            dist.broadcast(t, src=0)
            ok = bool(t.item() > 0.5)
        if not ok:
            skipped.append({"leg": leg, "estimated_s": cost_s, "time_left_s": left})
        return ok

    import torch
    import torch.distributed as dist

    import ldpc_decoder_b200 as L
    from ldpc_decoder_b200 import harness

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

    gen = make_code(args)
    code = L.Code.from_generated(gen)
    fp16 = args.precision == "fp16"
    prec = L.FP16 if fp16 else L.FP32
    dec = L.Decoder(code, prec, args.log2_frames, device=local_rank)
    P = dec.parallel_factor()
    n_frames = P * args.loading  # frames per step per GPU (weak scaling: every rank decodes its own queue shard)
    W = gen.n_vars // 32
    # the reference's fp16 program parses the noise level into the llr type (main.cpp:57,163)
    sigma = float(np.float16(args.sigma)) if fp16 else args.sigma
    prm = harness.channel_param("awgn", sigma)

    # the frame pool: distinct frames per rank (absolute frame indices rank*pool ..), resident in HBM for the `value`
    # leg, copied to host memory for the e2e leg
    d_in, d_syn, d_ref = synth_pool(dec, gen, L, L.AWGN, sigma, args.pool, fp16, start=rank * args.pool)
    d_res = torch.zeros((n_frames, W), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()

    base_flags = L.FLAG_EARLY_RETIRE if args.early_retire else 0
    common = dict(num_iter_max=args.imax, num_iter_check_parity=args.check_every)

    def step_device(flags=0):
        return dec.decode(L.AWGN, prm, d_in, d_syn, n_frames=n_frames, results=d_res, pool_frames=args.pool,
                          flags=flags | base_flags, **common)

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

    # sanity: the decode is real work -- converged frames equal the transmitted words and satisfy their syndromes
    res_host = d_res.cpu().numpy().view(np.uint32)
    ref_bits = d_ref.cpu().numpy().view(np.uint32)
    syn_host = d_syn.cpu().numpy().view(np.uint32)
    conv = iters < args.imax
    idx = np.arange(n_frames) % args.pool
    words_ok = bool(np.array_equal(res_host[conv], ref_bits[idx[conv]]))
    sel = np.nonzero(conv)[0][:64]
    ok = bool(np.array_equal(gen.syndrome_packed(res_host[sel]), syn_host[idx[sel]])) if len(sel) else False

    peak, peak_kind = load_peak()
    s = 2 if fp16 else 4
    roofline = roofline_block(reps, gen, P, s, peak, peak_kind, f"{args.precision}:{P}" if args.log2_n == 20 else None)
    avg_iter = float(np.mean([r["avg_iter"] for r in reps]))

    # e2e: the same call with HOST buffers (pageable numpy arrays, as a caller of the reference API has them): the
    # strided gather of prepare_vectors, H2D of channel values + syndromes, D2H of packed decisions all inside
    e2e = e2e_small = None
    step_s = elapsed / args.steps
    if not args.skip_e2e:
        inp_host = d_in.cpu().numpy()
        h_res = np.zeros((n_frames, W), dtype=np.uint32)

        def e2e_leg(inp, syn, pool, steps):
            def step_host():
                return dec.decode(L.AWGN, prm, inp, syn, n_frames=n_frames, results=h_res, pool_frames=pool,
                                  flags=base_flags, **common)
            step_host()  # warm-up of the host path
            barrier()
            t0 = time.perf_counter()
            hreps = []
            for _ in range(steps):
                _, _, r = step_host()
                hreps.append(r)
            torch.cuda.synchronize()
            e_elapsed = max_over_ranks(time.perf_counter() - t0)
            barrier()
            frames = sum_over_ranks(float(n_frames * steps))
            return {"value": frames * gen.n_vars / MBIT / e_elapsed, "unit": "Mb/s", "avg_iter": float(hreps[-1]["avg_iter"]),
                    "h2d_bytes_per_step": int(hreps[-1]["h2d_bytes"]), "d2h_bytes_per_step": int(hreps[-1]["d2h_bytes"]),
                    "ms_per_step": 1e3 * e_elapsed / steps, "steps": steps, "host_frame_pool": pool,
                    "host_stage_s_per_step": float(np.mean([r["host_stage_s"] for r in hreps])),
                    "first_batch_stage_s": float(np.mean([r["first_batch_stage_s"] for r in hreps])),
                    "device_s_per_step": float(np.mean([r["device_s"] for r in hreps])),
                    "input_row_stride_bytes": pool * s}

        e2e = e2e_leg(inp_host, syn_host, args.pool, max(1, min(args.steps, args.e2e_steps)))
        if args.small_pool and args.small_pool < args.pool and time_left_for("e2e_small_pool", 3.2 * step_s):
            sp = args.small_pool
            e2e_small = e2e_leg(np.ascontiguousarray(inp_host[:, :sp]), np.ascontiguousarray(syn_host[:sp]), sp,
                                min(args.steps, 2))
        del inp_host

    # one queue of world x n_frames frames shared by all ranks (one process per GPU, atomic cursor in POSIX shared
    # memory): dynamic load balance, no data-path collective; and the strong-scaling point (n_frames frames in total)
    shared = None
    if world > 1 and not args.skip_extra and time_left_for("shared_queue", 3.5 * step_s + 20):
        shared = {}
        name = "ldpc_b200_bench_%s" % os.environ.get("MASTER_PORT", "0")
        for label, n_total in (("weak", n_frames * world), ("strong", n_frames)):
            q = None
            if rank == 0:
                q = L.FrameQueue(n_total, 256, shm_name=name + label, create=True)
            barrier()
            if rank != 0:
                q = L.FrameQueue(shm_name=name + label, create=False)
            d_res_q = torch.empty((n_total, W), dtype=torch.int32, device="cuda")
            it_q = np.zeros(n_total, dtype=np.uint32)
            secs, mine = [], 0
            for i in range(2):  # one warm-up, one timed
                barrier()
                if rank == 0:
                    q.reset()
                barrier()
                t0 = time.perf_counter()
                _, r = dec.decode_from_queue(q, L.AWGN, prm, d_in, d_syn, d_res_q, iter_counts=it_q,
                                             pool_frames=args.pool, flags=base_flags, **common)
                torch.cuda.synchronize()
                secs.append(max_over_ranks(time.perf_counter() - t0))
                mine = r["frames_decoded"]
            counts = torch.zeros(world, dtype=torch.float64, device="cuda")
            counts[rank] = mine
            dist.all_reduce(counts)
            shared[label] = {"frames": n_total, "value": n_total * gen.n_vars / MBIT / secs[-1], "unit": "Mb/s",
                             "s": secs[-1], "frames_per_rank": [int(x) for x in counts.tolist()],
                             "scaling": label, "block_frames": 256}
            barrier()
            q.close()
            del d_res_q
            torch.cuda.empty_cache()

    # EXTENSION, not the reference's end-of-input behaviour: frames retire when they stop once the
    # input is exhausted and the survivors are compacted -- reported beside the headline, never as the headline
    early = None
    if not args.skip_extra and not args.early_retire and time_left_for("extension_early_retire", 2.1 * step_s):
        def step_early():
            return dec.decode(L.AWGN, prm, d_in, d_syn, n_frames=n_frames, results=d_res, pool_frames=args.pool,
                              flags=L.FLAG_EARLY_RETIRE, **common)
        step_early()
        barrier()
        t0 = time.perf_counter()
        _, it_e, rep_e = step_early()
        torch.cuda.synchronize()
        e_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        early = {"flag": "LDPC_B200_FLAG_EARLY_RETIRE", "value": sum_over_ranks(float(n_frames)) * gen.n_vars / MBIT / e_s,
                 "unit": "Mb/s", "ms_per_step": 1e3 * e_s, "steps": 1,
                 "slot_iterations_vs_reference_schedule": rep_e["slot_iterations"] / float(reps[-1]["slot_iterations"]),
                 "iteration_counts_identical": bool(np.array_equal(it_e, iters))}

    cpu = ref_cuda = others = None
    if rank == 0 and world == 1:
        if not args.skip_cpu:
            cpu, _ = cpu_sample(args, gen)  # ~10-25 s: part of the contract, not optional
        if not args.skip_extra:
            if time_left_for("ref_cuda_build", 45):
                try:
                    ref_cuda = ref_cuda_build_leg(L, gen, args, sigma, prm, d_in, d_syn, args.pool)
                except Exception as e:  # a baseline leg must never take the headline line down
                    ref_cuda = {"unavailable": repr(e)}
            del dec, d_in, d_syn, d_ref, d_res
            torch.cuda.empty_cache()
            if time_left_for("other_configs", 75):
                try:
                    others = other_configs_leg(L, peak, peak_kind)
                except Exception as e:
                    others = {"unavailable": repr(e)}

    if rank == 0:
        line = {
            "metric": "decoded Mb/s", "value": value, "unit": "Mb/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f16" if fp16 else "f32",
            "data": "synthetic", "config": workload_config(args, gen, P),
            "timing": {"host_bracketed_s": elapsed, "cuda_event_s": elapsed_events,
                       "note": "value = frames / host_bracketed_s (barrier + synchronize on both sides, max over ranks); "
                               "cuda_event_s = sum over steps of the CUDA-event time on the decoder's stream, max over ranks"},
            "avg_iter": avg_iter,
            # decoded Mb/s is proportional to 1/avg_iter: the same kernels at the README run's 90.7 iterations per frame
            "value_at_readme_avg_iter": value * avg_iter / 90.7148,
            "decoding_throughput_mbps": gen.n_vars / MBIT / (avg_iter * (roofline["cn"]["avg_ms"] + roofline["vn"]["avg_ms"]) * 1e-3 / P),
            # the same throughput as a fraction of the HBM roofline: frames/s <= peak / (B_it * avg_iter)
            "roofline_mbps": peak * 1e9 / (((4 * gen.n_edges + gen.n_vars) * s + 4 * ((gen.n_checks + 31) // 32)) * avg_iter) * world * gen.n_vars / MBIT,
            "frames_checked": int(conv.sum()), "converged_frames_equal_transmitted_words": words_ok,
            "syndromes_satisfied": ok, "frac_converged": float(conv.mean()),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_small_pool": e2e_small,
            "ref_cuda_build": ref_cuda, "other_configs": others, "shared_queue": shared, "extension_early_retire": early,
            "gpu_launches": int(launches), "clocks": clocks, "skipped_legs": skipped,
            "wall_s": time.perf_counter() - t_start,
        }
        line["value_over_roofline"] = line["value"] / line["roofline_mbps"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
