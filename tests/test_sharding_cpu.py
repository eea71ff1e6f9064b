"""world_size-2 gloo test of the multi-GPU host logic (frame-queue sharding + host-side merge).  The per-rank
decoder is the CPU oracle here (test infrastructure); on a GPU box the same function wraps Decoder.decode."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from ldpc_decoder_b200 import sharding

from conftest import ROOT


def test_shard_bounds_cover_queue():
    for n, w in [(0, 2), (1, 2), (7, 2), (8, 3), (100, 8), (5, 8)]:
        b = [sharding.shard_bounds(n, w, r) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        sizes = [hi - lo for lo, hi in b]
        assert max(sizes) - min(sizes) <= 1


def test_merge_reports():
    m = sharding.merge_reports([{"max_iter": 31, "min_iter": 10, "avg_iter": 20.0},
                                {"max_iter": 21, "min_iter": 11, "avg_iter": 12.0}, {}], [3, 1, 0])
    assert m == {"max_iter": 31, "min_iter": 10, "avg_iter": 18.0, "n_frames": 4}


WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch.distributed as dist
from ldpc_decoder_b200 import alist, harness, sharding
from oracle import binding as ob

dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
gen = alist.make_regular_code(384, 3, 6, seed=21)
text = alist.alist_text(gen)
inp, syn, bits = harness.make_frames(gen, 13, "awgn", 0.74, seed=5)
prm = harness.channel_param("awgn", 0.74)
code = ob.OracleCode(text=text)

def decode_fn(i, s):
    return ob.OracleDecoder(code, ob.FP32, 2, n_threads=1).decode(ob.AWGN, prm, 40, 5, i, s)

res, iters, rep = sharding.decode_sharded(decode_fn, inp, syn)
# single-process answer, shard by shard (frames are independent; only the first-batch +1 offset depends on the shard)
exp_res, exp_it = [], []
for r in range(2):
    lo, hi = sharding.shard_bounds(13, 2, r)
    a, b, _ = decode_fn(inp[:, lo:hi].copy(), syn[lo:hi].copy())
    exp_res.append(a); exp_it.append(b)
assert np.array_equal(res, np.concatenate(exp_res)) and np.array_equal(iters, np.concatenate(exp_it))
assert rep["n_frames"] == 13 and rep["max_iter"] == iters.max() and rep["min_iter"] == iters.min()
assert abs(rep["avg_iter"] - iters.mean()) < 1e-4
assert harness.count_bit_errors(res, bits).sum() == 0
dist.barrier()
dist.destroy_process_group()
print("rank", sys.argv[1], "ok")
"""


def test_two_rank_gloo_sharded_decode(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=port))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True, env=env) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o


QUEUE_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np
import torch
import torch.distributed as dist
import ldpc_decoder_b200 as L
from ldpc_decoder_b200 import alist, harness
from oracle import binding as ob

rank = int(sys.argv[1])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=2)
# one queue of frames in POSIX shared memory, two processes (what bench.py --gpus N does with one process per GPU):
# rank 0 creates it, rank 1 opens it; every process takes blocks until the cursor has passed the end
name = "ldpc_b200_gloo_test_{port}"
n_total, block = 53, 4
q = L.FrameQueue(n_total, block, shm_name=name, create=True) if rank == 0 else None
dist.barrier()
if rank != 0:
    q = L.FrameQueue(shm_name=name, create=False)
assert q.size == n_total
gen = alist.make_regular_code(384, 3, 6, seed=21)
inp, syn, bits = harness.make_frames(gen, n_total, "awgn", 0.74, seed=5)
prm = harness.channel_param("awgn", 0.74)
code = ob.OracleCode(text=alist.alist_text(gen))
res = np.zeros((n_total, gen.n_vars // 32), dtype=np.uint32)
mine = np.zeros(n_total, dtype=np.int64)
while True:
    first, cnt = q.take()
    if cnt == 0:
        break
    assert first % block == 0 and cnt == min(block, n_total - first)
    # the CPU oracle stands in for the per-GPU decoder (test infrastructure): decode this block's frames
    r, it, _ = ob.OracleDecoder(code, ob.FP32, 2, n_threads=1).decode(ob.AWGN, prm, 40, 5, inp[:, first:first + cnt].copy(),
                                                                       syn[first:first + cnt].copy())
    res[first:first + cnt] = r
    mine[first:first + cnt] = 1
taken = torch.from_numpy(mine.copy())
dist.all_reduce(taken)
assert taken.min().item() == 1 and taken.max().item() == 1, "every frame decoded by exactly one process"
merged = torch.from_numpy(res.astype(np.int64))
dist.all_reduce(merged)  # disjoint rows: the sum is the union (the host-side merge; no collective on the data path)
assert harness.count_bit_errors(merged.numpy().astype(np.uint32), bits).sum() == 0
dist.barrier()
q.close()
dist.destroy_process_group()
print("rank", rank, "ok", int(mine.sum()))
"""


def test_two_processes_share_one_frame_queue(tmp_path):
    """The shared frame queue across PROCESSES (one per GPU on a GPU box): POSIX shared-memory cursor, rank 0 creates,
    rank 1 opens, blocks are disjoint and cover the queue; world_size-2 gloo, the CPU oracle as the per-rank decoder."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "queue_worker.py"
    script.write_text(QUEUE_WORKER.format(root=ROOT, port=port))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True, env=env) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert f"rank {r} ok" in o
