"""BASELINE.json configs 4 and 5 on G GPUs (launch under torchrun, one rank per GPU).

  config 4: Q = 100 000 queries x N = 10 000 000 corpus rows x 768, top-100, corpus sharded over G
  config 5: Q = 1 000 000 x N = 1 000 000 x 768, top-100 + IR metrics (8 relevant docs / query)

Every rank owns Q/G queries (`ShardedCorpus.topk_owned`).  Checks are size-independent properties:
every certificate margin > 0 after the re-scan (exactness proven on the device), rankings sorted,
ids in range and distinct, and a brute-force fp32 comparison (torch matmul on the same GPU -- test
infrastructure, not the product path) for a sample of each rank's queries.  Timing: CUDA events, max
over ranks.  Sizes can be scaled down with QST_C4_N / QST_C4_Q / QST_C5_N / QST_C5_Q for a dry run.
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qst_b200  # noqa: E402
from qst_b200 import metrics, sharded  # noqa: E402

D, K = 768, 100
SLAB = 125_000


def log(rank, *a):
    if rank == 0:
        print(*a, file=sys.stderr, flush=True)


def make_corpus(n, dev, seed):
    gen = torch.Generator(device=dev).manual_seed(seed)
    full = torch.empty((n, D), dtype=torch.float32, device=dev)
    for s in range(0, n, SLAB):
        e = min(n, s + SLAB)
        full[s:e] = torch.randn(e - s, D, generator=gen, device=dev, dtype=torch.float32)
    return full


def max_over_ranks(ms, dev):
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def brute_force_topk(q, full, k, chunk=500_000):
    """torch fp32 cos_sim + topk (the parity definition of BASELINE.json), chunked over the corpus."""
    qn = torch.nn.functional.normalize(q, p=2, dim=1)
    best_v = torch.full((q.shape[0], 0), 0.0, device=q.device)
    best_i = torch.zeros((q.shape[0], 0), dtype=torch.long, device=q.device)
    for s in range(0, full.shape[0], chunk):
        cn = torch.nn.functional.normalize(full[s:s + chunk], p=2, dim=1)
        sc = qn @ cn.T
        v, i = torch.topk(sc, min(k, sc.shape[1]), dim=1)
        best_v = torch.cat([best_v, v], 1)
        best_i = torch.cat([best_i, i + s], 1)
        v, o = torch.topk(best_v, min(k, best_v.shape[1]), dim=1)
        best_v, best_i = v, torch.gather(best_i, 1, o)
    return best_v, best_i


def check_rankings(vals, idx, margin, n_total):
    assert bool((margin > 0).all()), f"{int((margin <= 0).sum())} queries without certificate"
    assert bool((vals[:, 1:] <= vals[:, :-1]).all()), "rankings not sorted"
    assert int(idx.min()) >= 0 and int(idx.max()) < n_total, "ids out of range"
    srt = torch.sort(idx, dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all()), "duplicate ids in a ranking"


def compare_sample(own_q, vals, idx, full, n_sample=32):
    sel = torch.linspace(0, own_q.shape[0] - 1, n_sample, device=own_q.device).long()
    bv, bi = brute_force_topk(own_q[sel], full, K)
    differ = idx[sel] != bi
    # where ids differ it must be a swap between scores that agree within float rounding (the two
    # implementations sum in different orders): the VALUES at those positions still line up
    tie_ok = ((vals[sel] - bv).abs() <= 2e-6) | ~differ
    same_ids = int((~differ).all(dim=1).sum())
    up_to_ties = int(tie_ok.all(dim=1).sum())
    dv = float((vals[sel] - bv).abs().max())
    return (same_ids, up_to_ties), n_sample, dv


def config4(rank, world, dev):
    n = int(os.environ.get("QST_C4_N", 10_000_000))
    q_total = int(os.environ.get("QST_C4_Q", 100_000))
    q_own = q_total // world
    full = make_corpus(n, dev, 14 + 400)
    n0, n1 = sharded.shard_bounds(n, world, rank)
    t0 = time.time()
    corp = sharded.ShardedCorpus(full[n0:n1], n, "cos_sim", full_master=full)
    torch.cuda.synchronize()
    log(rank, f"[cfg4] corpus {n} rows: shard prep {time.time() - t0:.2f} s, "
              f"mem {torch.cuda.memory_allocated(dev) / 2**30:.1f} GiB")
    qgen = torch.Generator(device=dev).manual_seed(14 + 500 + rank)
    own_q = torch.randn(q_own, D, generator=qgen, device=dev)
    corp.enable_stage_timing()
    for _ in range(2):
        out = corp.topk_owned(own_q, K)
    torch.cuda.synchronize()
    dist.barrier()
    steps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = corp.topk_owned(own_q, K)
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1) / steps, dev)
    vals, idx, margin = out
    check_rankings(vals, idx, margin, n)
    same, ns, dv = compare_sample(own_q, vals, idx, full)
    stat = torch.tensor([same[0], ns, same[1]], dtype=torch.int64, device=dev)
    dist.all_reduce(stat)
    stage = corp.stage_ms()
    res = {"config": 4, "Q": q_own * world, "N": n, "D": D, "k": K, "gpus": world, "ms_per_step": ms,
           "queries_per_s": q_own * world / ms * 1e3,
           "tflops_per_gpu": 2.0 * q_own * world * (n1 - n0) * D / (stage.get("K2", ms) * 1e-3) / 1e12,
           "stage_ms_rank0": stage, "all_margins_positive": True,
           "brute_force_sample": {"queries": int(stat[1]), "identical_rankings": int(stat[0]),
                                  "identical_up_to_ties_within_2e-6": int(stat[2]),
                                  "max_abs_value_diff_rank0": dv}}
    del corp, full, out, vals, idx, margin
    torch.cuda.empty_cache()
    return res


def config5(rank, world, dev):
    n = int(os.environ.get("QST_C5_N", 1_000_000))
    q_total = int(os.environ.get("QST_C5_Q", 1_000_000))
    q_own = q_total // world
    batch = min(q_own, int(os.environ.get("QST_C5_BATCH", 12_500)))
    full = make_corpus(n, dev, 14 + 100)
    n0, n1 = sharded.shard_bounds(n, world, rank)
    corp = sharded.ShardedCorpus(full[n0:n1], n, "cos_sim", full_master=full)
    # query g (global id) is a noisy copy of corpus row rel(g); its 8 relevant docs are that row and 7
    # rows spread over the corpus (which it will mostly not retrieve): recall@100 ~ 1/8, mrr ~ 1
    gid = torch.arange(rank * q_own, (rank + 1) * q_own, device=dev, dtype=torch.long)
    rel0 = (gid * 7919) % n
    qgen = torch.Generator(device=dev).manual_seed(14 + 600 + rank)
    own_q = full[rel0] + 0.5 * torch.randn(q_own, D, generator=qgen, device=dev)
    rel = (rel0[:, None] + torch.arange(8, device=dev)[None, :] * (n // 8)) % n
    rel = torch.sort(rel, dim=1).values
    rowptr = torch.arange(0, 8 * q_own + 1, 8, device=dev, dtype=torch.long)
    cols = rel.reshape(-1).contiguous()
    ks = [1, 10, 100]

    def run():
        ranked = torch.empty((q_own, K), dtype=torch.long, device=dev)
        bad = 0
        for s in range(0, q_own, batch):
            v, i, m = corp.topk_owned(own_q[s:s + batch], K)
            ranked[s:s + batch] = i
            bad = bad + (m <= 0).sum()
        per_q = metrics.per_query_metrics(ranked, rowptr, cols, ks)
        return ranked, per_q, bad

    assert q_own % batch == 0, "every rank must run the same number of equally sized batches"
    run()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ranked, per_q, bad = run()
    e1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1), dev)
    assert int(bad) == 0, f"{int(bad)} queries without certificate"
    # cross-query reductions the way the reference does them (host, float64), over ALL queries:
    # gather the per-query values on rank 0
    gathered = [torch.empty_like(per_q) for _ in range(world)] if rank == 0 else None
    dist.gather(per_q, gathered, dst=0)
    res = None
    if rank == 0:
        allq = torch.cat(gathered, dim=2).cpu().numpy()
        red = metrics.reduce_like_reference(allq, ks, accuracy_at_k=[1], precision_recall_at_k=[1, 10, 100],
                                            mrr_at_k=[10], ndcg_at_k=[10], map_at_k=[100])
        res = {"config": 5, "Q": q_own * world, "N": n, "D": D, "k": K, "gpus": world, "batch_per_gpu": batch,
               "ms_total": ms, "queries_per_s": q_own * world / ms * 1e3,
               "metrics": {"mrr@10": red["mrr@k"][10], "ndcg@10": red["ndcg@k"][10],
                           "recall@1": red["recall@k"][1], "recall@10": red["recall@k"][10],
                           "recall@100": red["recall@k"][100], "map@100": red["map@k"][100]},
               "first_hit_is_planted_row": float((ranked[:, 0] == rel0).float().mean())}
    # brute-force sample on this rank's first batch
    v, i, _ = corp.topk_owned(own_q[:batch], K)
    same, ns, dv = compare_sample(own_q[:batch], v, i, full)
    stat = torch.tensor([same[0], ns, same[1]], dtype=torch.int64, device=dev)
    dist.all_reduce(stat)
    if rank == 0:
        res["brute_force_sample"] = {"queries": int(stat[1]), "identical_rankings": int(stat[0]),
                                  "identical_up_to_ties_within_2e-6": int(stat[2]),
                                     "max_abs_value_diff_rank0": dv}
    return res


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    which = sys.argv[1] if len(sys.argv) > 1 else "45"
    out = []
    if "4" in which:
        r = config4(rank, world, dev)
        log(rank, json.dumps(r))
        out.append(r)
    if "5" in which:
        r = config5(rank, world, dev)
        log(rank, json.dumps(r))
        out.append(r)
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
