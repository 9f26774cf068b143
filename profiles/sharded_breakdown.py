"""Stage-by-stage CUDA-event timing of the candidate-exchange path on N ranks (torchrun)."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qst_b200  # noqa: E402
from qst_b200 import _lib, scoring, sharded  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
N, D, K = 1_000_000, 768, 100
Q = 10_000 * world
g = torch.Generator(device=dev).manual_seed(14)
full = torch.cat([torch.randn(125_000, D, generator=g, device=dev) for _ in range(8)])
queries = torch.randn(Q, D, generator=g, device=dev)
n0, n1 = sharded.shard_bounds(N, world, rank)
corp = sharded.ShardedCorpus(full[n0:n1], N, "cos_sim", full_master=full)
lib = _lib.load()
st = _lib.stream_ptr(dev)
G, q_own = world, Q // world
names = ["prep", "K2", "select", "all_to_all", "finalize_lists", "rescan", "all_gather"]
acc = {n: 0.0 for n in names}
iters = 8
for it in range(iters + 2):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    dist.barrier()
    torch.cuda.synchronize()
    ev[0].record()
    pq = scoring.prepare_rows(queries, scoring.QUERY_PREP["cos_sim"])
    ev[1].record()
    plan = scoring.make_plan(Q, corp.index.n, D, K, 0, "cos_sim")
    m = sharded.candidates_per_shard(plan.kprime, G)
    ws = scoring._workspace(plan.ws_bytes, dev, "select")
    _lib.check(lib.qst_score_select(C.byref(plan), pq.bf16.data_ptr(), corp.index.rows.bf16.data_ptr(), ws.data_ptr(), st))
    ev[2].record()
    lists = torch.empty((Q, m + 1, 2), dtype=torch.int32, device=dev)
    _lib.check(lib.qst_select_candidates(C.byref(plan), ws.data_ptr(), m, corp.start, lists.data_ptr(), st))
    ev[3].record()
    recv = sharded.exchange_candidate_lists(lists)
    ev[4].record()
    own = slice(rank * q_own, (rank + 1) * q_own)
    vals = torch.empty((q_own, K), dtype=torch.float32, device=dev)
    idx = torch.empty((q_own, K), dtype=torch.int64, device=dev)
    margin = torch.empty(q_own, dtype=torch.float32, device=dev)
    scratch = scoring._workspace(lib.qst_finalize_lists_scratch_bytes(q_own, G), dev, "lists")
    mst = corp.master
    _lib.check(lib.qst_finalize_lists(q_own, G, m, K, plan.kprime, 0, D, recv.data_ptr(), pq.f32[own].data_ptr(),
                                      pq.inv_norm[own].data_ptr(), pq.err[own].data_ptr(), mst.f32.data_ptr(),
                                      mst.inv_norm.data_ptr(), mst.stats.data_ptr(), vals.data_ptr(), idx.data_ptr(),
                                      margin.data_ptr(), scratch.data_ptr(), st))
    ev[5].record()
    rs = scoring._workspace(lib.qst_exact_rescan_workspace_bytes(q_own, K), dev, "rescan")
    _lib.check(lib.qst_exact_rescan(q_own, N, D, K, 0, pq.f32[own].data_ptr(), pq.inv_norm[own].data_ptr(),
                                    mst.f32.data_ptr(), mst.inv_norm.data_ptr(), 0, vals.data_ptr(), idx.data_ptr(),
                                    margin.data_ptr(), rs.data_ptr(), st))
    ev[6].record()
    gv, gi = sharded.all_gather_topk(vals, idx)
    ev[7].record()
    torch.cuda.synchronize()
    if it >= 2:
        for i, n in enumerate(names):
            acc[n] += ev[i].elapsed_time(ev[i + 1])
if rank == 0:
    tot = sum(acc.values()) / iters
    print(f"world={world} Q={Q} m={m} total {tot:.3f} ms: " + "  ".join(f"{n} {acc[n] / iters:.3f}" for n in names))

# the public call, back to back without host syncs (what bench.py times)
import time
for _ in range(3):
    corp.topk(queries, K)
dist.barrier(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
a.record()
for _ in range(10):
    out = corp.topk(queries, K)
b.record()
t_host = (time.perf_counter() - t0) / 10 * 1e3
torch.cuda.synchronize()
print(f"[rank {rank}] corp.topk back-to-back: {a.elapsed_time(b) / 10:.3f} ms/step on the device, host enqueue {t_host:.3f} ms/step", flush=True)
# same again, but exactly as bench.py brackets it
def barrier():
    dist.barrier()
    torch.cuda.synchronize()
for rep in range(2):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        out = corp.topk(queries, K)
    e1.record()
    barrier()
    print(f"[rank {rank}] bench-style bracket rep {rep}: {e0.elapsed_time(e1) / 10:.3f} ms/step", flush=True)
dist.destroy_process_group()
