#!/usr/bin/env python
"""Benchmark of the retrieval-scoring hot path (BASELINE.json metric: queries/sec on a 1M x 768
corpus, top-100), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path over one batch of synthetic queries: K1 on the queries, K2
tensor-core scoring + streaming top-k', K3 fp32 rescoring/ordering (+ the exact re-scan kernels, which
find nothing to do when every query is certified).

Workload (config 3 of BASELINE.json): 10 000 queries x 1 000 000 corpus rows x 768-d, fp32 masters ->
bf16 tensor-core operands, k = 100.  For N > 1 the 1M-row corpus is sharded over the N GPUs -- bf16
operand AND fp32 master, every rank keeps its rows only -- and the query batch
grows to N x 10 000 (every rank owns 10 000), so the per-GPU work is fixed (weak scaling) and `value`
is all queries of all ranks / max-over-ranks device time.  Exchanges per step: all-gather of the fp32
queries, ONE all-to-all of candidate lists to the owner of each query, requests to the shards and exact
scores back (two small all-to-alls); thresholds are shared between shards through peer memory.

  value      inputs resident in HBM, CUDA-event time on the launching stream
  e2e        the same step through the host-buffer entry: pinned host queries are copied in and the
             ranking is copied back inside the timed region, every step: double-buffered
             (`qst_b200.HostTopkPipeline` / `qst_b200.sharded.ShardedHostPipeline`) and as one synchronous
             call per step; `value` is the faster of the two (`mode` says which), both are kept
  roofline   dominant kernel (K2): 2*Q*N*D FLOP per launch / its mean duration inside the steps, against
             the measured sustained bf16 peak of MEASURED_PEAKS.json
  parity_sample  >= 256 queries of the timed batch re-computed with plain torch fp32
             (F.normalize -> mm -> topk, every rank over its own rows, merged on rank 0) and compared
             with the rankings the timed path produced; a mismatch that is not a tie within 2e-6 makes
             the run exit non-zero
  cpu_baseline  the CPU oracle (restated sentence-transformers 2.2.2 path: cos_sim -> per-chunk
             torch.topk -> .tolist() -> per-hit dict lists -> sorted) on 1000 queries against the FULL
             1M-row corpus (20 chunks of 50 000), all host threads: queries/s as measured, no scaling
  secondary  loss_config2 (N = 1); f2_config1 (N = 1: the whole evaluator call, k up to 900, three score
             functions); replicated_master, config4 (100k x 10M) for N >= 2; config5 (1M x 1M + metrics)
             for N = 8
`--impl reference` times only the CPU path and prints it in the same format, with the same `metric`, `unit` and
`config.workload` (what differs between the arms is in `config.implementation`).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION in this
# image) off it
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

Q_PER_GPU = 10_000
N_CORPUS = 1_000_000
DIM = 768
TOPK = 100
NO_PREFETCH = bool(os.environ.get("QST_BENCH_NO_PREFETCH"))   # A/B switch: sharded steps without topk_owned(prefetch=)
CPU_SAMPLE_Q = 1000
SLAB = 125_000            # the synthetic corpus is generated slab by slab, slab i from seed CORPUS_SEED + i
CORPUS_SEED = 14 + 1000
PARITY_SAMPLE = 256
# the workload both arms (`--impl ours` / `--impl reference`) name in config.workload: BASELINE.json configs[2],
# weak scaling (every GPU brings Q_PER_GPU queries; the corpus is sharded over the GPUs)
WORKLOAD = (f"config 3: cos_sim + exact top-{TOPK}, {Q_PER_GPU} queries per GPU x {N_CORPUS} corpus rows x {DIM}-d "
            f"fp32 embeddings (synthetic N(0,1), seed {CORPUS_SEED})")
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            d = json.load(f)
        d["_source"] = "measured"
        return d
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback"
    return d


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index: int, period: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------
# CPU baseline (oracle) -- the only place bench.py touches oracle/
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(steps: int, warmup: int, budget_s: float = 150.0):
    """Times the restated reference CPU path on Q = 1000 queries against the FULL 1M x 768 corpus (20
    chunks of 50 000 rows, as corpus_chunk_size=50000 of ir_evauation_script.py:161 cuts it).  The
    metric is queries/s on this corpus, so nothing is extrapolated: value = 1000 / seconds per step.
    `steps`/`warmup` are honoured as long as the whole run fits `budget_s` seconds."""
    import torch
    from oracle import ir_oracle
    import qst_b200  # synthetic generators only (no GPU work here)

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    q = qst_b200.synth.gaussian_embeddings(CPU_SAMPLE_Q, DIM, 20)
    gen = torch.Generator().manual_seed(21)
    table = torch.empty((CPU_SAMPLE_Q + N_CORPUS, DIM), dtype=torch.float32)
    table[:CPU_SAMPLE_Q] = q
    for s0 in range(0, N_CORPUS, SLAB):
        table[CPU_SAMPLE_Q + s0:CPU_SAMPLE_Q + s0 + SLAB] = torch.randn(SLAB, DIM, generator=gen)
    queries = {str(i): str(i) for i in range(CPU_SAMPLE_Q)}
    corpus = {str(i): str(CPU_SAMPLE_Q + i) for i in range(N_CORPUS)}
    relevant = {str(i): {str(i)} for i in range(CPU_SAMPLE_Q)}
    ev = ir_oracle.InformationRetrievalEvaluatorOracle(
        queries, corpus, relevant, corpus_chunk_size=50000, mrr_at_k=[10], ndcg_at_k=[10],
        accuracy_at_k=[1], precision_recall_at_k=[1], map_at_k=[TOPK],
        score_functions={"cos_sim": ir_oracle.cos_sim}, write_csv=False)
    model = ir_oracle.PrecomputedEmbeddingModel(table)

    def one():
        hits = ev.collect_hits(model)                    # cos_sim -> topk -> tolist -> dict lists
        return ir_oracle.ranked_ids(hits["cos_sim"], TOPK)  # per-query stable sort

    t0 = time.perf_counter()
    one()                                                # first warm-up step, also the time estimate
    est = time.perf_counter() - t0
    warm_done = 1
    warmup = max(warmup, 1)
    steps_run = max(1, min(steps, int((budget_s - est * warmup) / max(est, 1e-3))))
    while warm_done < warmup and est * (warm_done + 1 + steps_run) <= budget_s:
        one()
        warm_done += 1
    t0 = time.perf_counter()
    for _ in range(steps_run):
        one()
    dt = (time.perf_counter() - t0) / steps_run
    return {"value": CPU_SAMPLE_Q / dt, "unit": "queries/s", "cores": cores, "kind": "port",
            "sample": f"Q={CPU_SAMPLE_Q} queries x the full N={N_CORPUS} x D={DIM} corpus, k={TOPK}, 20 chunks of 50000; "
                      f"{dt * 1e3:.0f} ms/step measured over {steps_run} steps after {warm_done} warm-up; not extrapolated "
                      f"(queries/s scales linearly in the number of queries)",
            "ms_per_sample_step": dt * 1e3, "steps": steps_run, "warmup": warm_done}


def loss_config2(dev, peaks):
    """BASELINE.json config 2 (secondary line): GammaQuadrupletLoss fwd+bwd, 4096 x 768 fp32.
    Inputs rotate through 12 independent sets (604 MB > L2) so every launch reads HBM.  The kernels
    are replayed from a CUDA graph (12 launches per replay) so the number is device time, not the
    Python launch overhead; the per-call wall time of the Python API is reported next to it."""
    import ctypes as C
    import torch
    import qst_b200
    from qst_b200 import _lib, quad_loss
    B, D, sets = 4096, 768, 12
    lib = _lib.load()
    g = torch.Generator(device=dev).manual_seed(14 + 300)
    data = [[torch.randn(B, D, generator=g, device=dev) for _ in range(4)] for _ in range(sets)]
    grads = [[torch.empty(B, D, device=dev) for _ in range(4)] for _ in range(sets)]
    loss = torch.empty(sets, device=dev)
    saved = torch.empty(B, _lib.QST_QUAD_SAVED_PER_ROW, device=dev)
    gout = torch.ones(1, device=dev)
    ws = torch.zeros(lib.qst_quadruplet_workspace_bytes(), dtype=torch.uint8, device=dev)
    prm = quad_loss._params(0.6, 1.0, 0.5, 0.5, 2.0, False)
    red = _lib.QST_RED_MEAN

    def launch_fused(st):
        for i in range(sets):
            x, gr = data[i], grads[i]
            _lib.check(lib.qst_quadruplet_fwd_bwd(x[0].data_ptr(), x[1].data_ptr(), x[2].data_ptr(), x[3].data_ptr(),
                                                  _lib.QST_F32, B, D, C.byref(prm), red, 1.0,
                                                  loss[i:].data_ptr(), gr[0].data_ptr(), gr[1].data_ptr(),
                                                  gr[2].data_ptr(), gr[3].data_ptr(), ws.data_ptr(), st))

    def launch_split(st):
        for i in range(sets):
            x, gr = data[i], grads[i]
            _lib.check(lib.qst_quadruplet_fwd(x[0].data_ptr(), x[1].data_ptr(), x[2].data_ptr(), x[3].data_ptr(),
                                              _lib.QST_F32, B, D, C.byref(prm), red, loss[i:].data_ptr(),
                                              saved.data_ptr(), ws.data_ptr(), st))
            _lib.check(lib.qst_quadruplet_bwd(x[0].data_ptr(), x[1].data_ptr(), x[2].data_ptr(), x[3].data_ptr(),
                                              _lib.QST_F32, B, D, C.byref(prm), red, saved.data_ptr(),
                                              gout.data_ptr(), gr[0].data_ptr(), gr[1].data_ptr(),
                                              gr[2].data_ptr(), gr[3].data_ptr(), st))

    out = {}
    for name, fn, launches in (("fused_fwd_bwd", launch_fused, sets), ("fwd_then_bwd", launch_split, 2 * sets)):
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            fn(side.cuda_stream)                      # warm-up outside capture
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                fn(torch.cuda.current_stream(dev).cuda_stream)
        torch.cuda.current_stream(dev).wait_stream(side)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        reps = 20
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            graph.replay()
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / (reps * sets)
        out[name] = {"us_per_step": us, "gbs_algorithmic": 8 * B * D * 4 / (us * 1e-6) / 1e9,
                     "kernel_launches_per_step": launches // sets}
    # the Python API as a user calls it: forward + backward through autograd, gradients reset to None
    # between steps as optimizer.zero_grad() does; wall time per step.  Next to it what torch's autograd
    # engine costs for a custom Function that launches NOTHING (the floor of any drop-in loss module),
    # the same step without autograd (loss_and_grads: one launch), and the reference's formulation
    # (three F.triplet_margin_loss calls, models/losses/losses.py:35-69) on the same GPU.
    import torch.nn.functional as F
    xs = [x.requires_grad_(True) for x in data[0]]
    mod = qst_b200.GammaQuadrupletLoss(gamma=0.6, margin_pos_neg=1.0, margin_pos_part=0.5, margin_part_neg=0.5)

    class _Floor(torch.autograd.Function):
        @staticmethod
        def forward(ctx, a, p, q, n, buf, l0):
            ctx.buf = buf
            return l0

        @staticmethod
        def backward(ctx, go):
            b = ctx.buf.unbind(0)
            return b[0], b[1], b[2], b[3], None, None

    fbuf, fl0 = torch.zeros(4, B, D, device=dev), torch.zeros((), device=dev)

    def torch_ops():
        a, p, q, n = xs
        return (F.triplet_margin_loss(a, p, n, margin=1.0) + 0.6 * F.triplet_margin_loss(a, q, n, margin=0.5)
                + 0.4 * F.triplet_margin_loss(a, p, q, margin=0.5))

    def wall(fn, iters, backward=True):
        for _ in range(10):
            for x in xs:
                x.grad = None
            r = fn()
            if backward:
                r.backward()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            for x in xs:
                x.grad = None
            r = fn()
            if backward:
                r.backward()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / iters * 1e6

    out["python_api_autograd_us_per_step"] = wall(lambda: mod(*xs), 200)
    out["torch_autograd_floor_us_per_step"] = wall(lambda: _Floor.apply(*xs, fbuf, fl0), 200)
    out["python_api_no_autograd_us_per_step"] = wall(lambda: mod.loss_and_grads(*xs)[0], 200, backward=False)
    out["torch_ops_reference_formulation_us_per_step"] = wall(torch_ops, 50)
    hbm = float(peaks.get("hbm_gbs", FALLBACK_PEAKS["hbm_gbs"]))
    out["algorithmic_bytes"] = 8 * B * D * 4
    out["hbm_peak_gbs"] = hbm
    out["fused_frac_of_hbm_peak"] = out["fused_fwd_bwd"]["gbs_algorithmic"] / hbm
    out["note"] = "device time from CUDA-graph replays over 12 rotating input sets (604 MB > L2)"
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base = cpu_reference_run(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "queries/sec, 1Mx768 corpus top-100", "value": base["value"],
        "unit": "queries/s", "n_gpus": args.gpus, "steps": base["steps"], "warmup": base["warmup"],
        "ms_per_step": base["ms_per_sample_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "implementation": f"the reference's path restated (oracle port) on the host CPUs; a step = "
                                     f"{CPU_SAMPLE_Q} queries against the full corpus, see cpu_baseline.sample"},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def corpus_rows(n0, n1, dev, seed=CORPUS_SEED):
    """Rows [n0, n1) of the synthetic fp32 corpus.  Slab i (SLAB rows) is drawn from seed + i, so the
    corpus is the same whatever the number of ranks, and a rank only generates the slabs it keeps."""
    import torch
    out = torch.empty((n1 - n0, DIM), dtype=torch.float32, device=dev)
    for sl in range(n0 // SLAB, -(-n1 // SLAB)):
        g = torch.Generator(device=dev).manual_seed(seed + sl)
        rows = torch.randn(SLAB, DIM, generator=g, device=dev, dtype=torch.float32)
        lo, hi = max(n0, sl * SLAB), min(n1, (sl + 1) * SLAB)
        out[lo - n0:hi - n0] = rows[lo - sl * SLAB:hi - sl * SLAB]
    return out


def brute_force_local(q, rows, start, k, chunk=250_000):
    """Plain torch fp32 cos_sim + topk of `q` against `rows` (global id = start + row): the arithmetic of
    the reference's path (F.normalize -> mm -> topk), used only to CHECK the timed path."""
    import torch
    F = torch.nn.functional
    qn = F.normalize(q, p=2, dim=1)
    best_v = torch.empty((q.shape[0], 0), device=q.device)
    best_i = torch.empty((q.shape[0], 0), dtype=torch.long, device=q.device)
    for s0 in range(0, rows.shape[0], chunk):
        sc = qn @ F.normalize(rows[s0:s0 + chunk], p=2, dim=1).T
        v, i = torch.topk(sc, min(k, sc.shape[1]), dim=1)
        best_v, best_i = torch.cat([best_v, v], 1), torch.cat([best_i, i + (start + s0)], 1)
        v, o = torch.topk(best_v, min(k, best_v.shape[1]), dim=1)
        best_v, best_i = v, torch.gather(best_i, 1, o)
    return best_v, best_i


def parity_sample(comm, own_queries, vals, idx, rows_f32, start, k, n_sample=PARITY_SAMPLE):
    """>= n_sample queries of the timed batch against a brute-force fp32 torch ranking (every rank
    scores the sample against ITS rows, the per-rank lists are merged).  Same result on every rank."""
    import torch
    G = comm.world
    per = -(-n_sample // G)
    sel = torch.linspace(0, own_queries.shape[0] - 1, per, device=own_queries.device).long()
    q_s = comm.all_gather(own_queries[sel].float().contiguous())              # [G*per, D]
    ours_v, ours_i = comm.all_gather(vals[sel].contiguous()), comm.all_gather(idx[sel].contiguous())
    lv, li = brute_force_local(q_s, rows_f32, start, k)
    gv = comm.all_gather(lv).view(G, G * per, -1).permute(1, 0, 2).reshape(G * per, -1)
    gi = comm.all_gather(li).view(G, G * per, -1).permute(1, 0, 2).reshape(G * per, -1)
    bv, o = torch.topk(gv, k, dim=1)
    bi = torch.gather(gi, 1, o)
    differ = ours_i != bi
    n = G * per
    identical = int((~differ).all(dim=1).sum())
    # Scores we REPORT must agree with the reference's at every rank (same document or a tie partner);
    # 5e-6: two fp32 summation orders of a 768-term dot product (observed: up to 9e-7 at |score| ~ 1).
    value_bad = (ours_v - bv).abs() > 5e-6
    ok_pos = ~differ
    settled = 0
    if bool(differ.any()):
        # A different document at a rank is legal only as a swap inside a tie.  Neither side's fp32 value is
        # trusted for that: both documents are scored in float64 by the rank that holds them, and the swap
        # counts as a tie when the two exact scores agree within 2e-6 (fp32 rounding of a 768-term dot
        # product, the reason two fp32 implementations may order such a pair differently).
        bad_q, bad_j = differ.nonzero(as_tuple=True)
        qd = torch.nn.functional.normalize(q_s[bad_q].double(), p=2, dim=1)

        def exact64(ids):
            loc = ids - start
            mine = (loc >= 0) & (loc < rows_f32.shape[0])
            sc = torch.full((ids.numel(),), float("-inf"), dtype=torch.float64, device=ids.device)
            if bool(mine.any()):
                rows = torch.nn.functional.normalize(rows_f32[loc[mine]].double(), p=2, dim=1)
                sc[mine] = (qd[mine] * rows).sum(1)
            # every other rank holds -inf for this row: the max over ranks is the owner's value
            return comm.all_reduce_max(sc.float()).double() if G > 1 else sc

        s_ours, s_ref = exact64(ours_i[bad_q, bad_j]), exact64(bi[bad_q, bad_j])
        tie = (s_ours - s_ref).abs() <= 2e-6
        ok_pos = ok_pos.clone()
        ok_pos[bad_q[tie], bad_j[tie]] = True
        settled = int(tie.sum())
    ok_rows = ok_pos.all(dim=1) & ~value_bad.any(dim=1)
    up_to_ties = int(ok_rows.sum())
    return {"queries": n, "identical": identical, "ties_within_1e-6": up_to_ties - identical,
            "mismatch": n - up_to_ties, "max_abs_score_diff": float((ours_v - bv).abs().max()),
            "tie_tolerance": 2e-6, "value_tolerance": 5e-6, "swapped_positions_checked_in_float64": settled,
            "reference": "torch fp32 F.normalize -> mm -> topk on the GPU, per shard, merged"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import ctypes as C
    import qst_b200
    from qst_b200 import comm as qcomm, scoring, sharded, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    cm = qcomm.default_comm()
    steps, warmup = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    corp_box = []      # [ShardedCorpus or None] once it exists (timed() finishes its deferred checks)

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.qst_launch_count()
        e0.record()
        for _ in range(n):
            out = fn()
        if corp_box and corp_box[0] is not None:
            corp_box[0].finish_exact()          # the last step's deferred certificate check
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / n, out, lib.qst_launch_count() - l0

    # ---- synthetic inputs: this rank's rows of the 1M corpus, this rank's 10 000 queries ----------
    Q = Q_PER_GPU * world
    n0, n1 = sharded.shard_bounds(N_CORPUS, world, rank)
    shard = corpus_rows(n0, n1, dev)
    qgen = torch.Generator(device=dev).manual_seed(14 + 200 + rank)
    queries = torch.randn(Q_PER_GPU, DIM, generator=qgen, device=dev, dtype=torch.float32)
    queries_host = queries.cpu().pin_memory()
    if world > 1:
        corp = sharded.ShardedCorpus(shard, N_CORPUS, "cos_sim", comm=cm)      # fp32 master sharded too
        index = corp.index
    else:
        corp = None
        index = qst_b200.CorpusIndex(shard, "cos_sim", idx_offset=0)
    rows_f32 = index.rows.f32
    corp_box.append(corp)
    del shard
    torch.cuda.synchronize()

    def step_device():
        if corp is not None:
            # exact="deferred": the certificate check of a step (one host read) is made when the NEXT step
            # is submitted, the last one inside timed() -- every step is checked within the timed region
            # prefetch=: the NEXT batch (here the same 10 000 rows again) is prepared and distributed
            # underneath this step's exchanges, every step, so no step waits for an all-gather before K2
            return corp.topk_owned(queries, TOPK, exact="deferred", prefetch=None if NO_PREFETCH else queries)
        r = scoring.topk(queries, index, TOPK)
        return r.values, r.indices, r.margin

    if corp is not None:
        corp.enable_stage_timing()
    for _ in range(warmup):
        out = step_device()
    sampler = ClockSampler(local_rank, float(os.environ.get("QST_BENCH_CLOCK_PERIOD", "0.02")))
    sampler.start()
    ms_step, out, launches = timed(step_device, steps)
    clocks = sampler.stop()
    vals, idx, margin = out
    unc = (~(margin > 0)).sum().to(torch.int64)
    if world > 1:
        dist.all_reduce(unc)
    uncertified = int(unc)
    parity = parity_sample(cm, queries, vals, idx, rows_f32, n0, TOPK)

    # ---- the dominant kernel alone, on the same stream, inside the same kind of step -----------
    plan = scoring.make_plan(Q, index.n, DIM, TOPK, 0, "cos_sim")
    stage = None
    if corp is not None:
        # sharded run: K2 is launched inside ShardedCorpus.topk_owned (peer-hint variant); its CUDA-event
        # bracket is recorded there for every timed step
        stage = corp.stage_ms()
        k2 = stage["K2"]
    else:
        pq = scoring.prepare_rows(queries, True)
        ws = scoring._workspace(plan.ws_bytes, dev, "select")
        st = _lib.stream_ptr(dev)
        k2_ms = []
        for i in range(warmup + steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            step_device()                                   # keep the device in the steady state of a step
            a.record()
            _lib.check(lib.qst_score_select(C.byref(plan), pq.bf16.data_ptr(), index.rows.bf16.data_ptr(),
                                            ws.data_ptr(), st))
            b.record()
            torch.cuda.synchronize()
            if i >= warmup:
                k2_ms.append(a.elapsed_time(b))
        k2 = sum(k2_ms) / len(k2_ms)

    # ---- end to end through the host-buffer entry --------------------------------------------
    host_out = {}

    def step_host():
        if corp is None:
            return scoring.topk_host(queries_host, index, TOPK)
        qd = queries_host.to(dev, non_blocking=True)
        v, i, _ = corp.topk_owned(qd, TOPK)
        if not host_out:   # pinned result buffers are allocated once (cudaHostAlloc costs ms)
            host_out["v"] = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
            host_out["i"] = torch.empty(i.shape, dtype=i.dtype, pin_memory=True)
        host_out["v"].copy_(v, non_blocking=True)
        host_out["i"].copy_(i, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return host_out["v"], host_out["i"]

    for _ in range(2):
        host_check = step_host()
    host_check = (host_check[0].clone(), host_check[1].clone())
    ms_e2e_serial, _, _ = timed(step_host, steps)
    pipe = scoring.HostTopkPipeline(index, TOPK) if corp is None else sharded.ShardedHostPipeline(corp, TOPK)
    for _ in range(3):
        pipe.submit(queries_host)
    pipe.drain()
    barrier()
    # a stream of batches through the double-buffered entry: every step still copies its own queries in
    # from pinned memory and its own ranking back out inside the timed region, but the copies of
    # neighbouring steps overlap the kernels (two slots, each with its own stream, workspace and pinned
    # result buffers)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream(dev)
    e0.record(cur)
    for st_ in pipe.streams:
        st_.wait_event(e0)
    for _ in range(steps):
        # sharded: the next batch (the same pinned rows again) is copied in and announced one step ahead
        ticket = pipe.submit(queries_host, queries_host) if (corp is not None and not NO_PREFETCH) else pipe.submit(queries_host)
    for st_ in pipe.streams:
        cur.wait_stream(st_)
    e1.record(cur)
    barrier()
    pv, pi = pipe.result(ticket)
    pipe_same = bool(torch.equal(pi, host_check[1]) and torch.equal(pv, host_check[0]))
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / steps
    e2e_mode = "double-buffered stream of host-buffer calls (%s)" % type(pipe).__name__
    del pipe

    # ---- secondary measurements (never allowed to break the headline line) -----------------------
    secondary = {}
    if world > 1 and not os.environ.get("QST_BENCH_SKIP_SECONDARY"):
        for name, fn in (("replicated_master", lambda: replicated_master_block(cm, dev, queries, n0, n1, rank, timed)),
                         ("config4", lambda: config4_block(cm, dev, rank, world, barrier, max_over_ranks))):
            try:
                secondary[name] = fn()
            except Exception as e:  # noqa: BLE001
                secondary[name] = {"error": repr(e)}
            torch.cuda.empty_cache()
        if world == 8 or os.environ.get("QST_BENCH_CONFIG5"):
            try:
                secondary["config5"] = config5_block(cm, corp, dev, rank, world, barrier, max_over_ranks)
            except Exception as e:  # noqa: BLE001
                secondary["config5"] = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    flop = 2.0 * Q * index.n * DIM
    achieved = flop / (k2 * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops_sustained", FALLBACK_PEAKS["bf16_tflops_sustained"]))
    traffic = None
    prof = os.path.join(ROOT, "profiles", "score_select_ncu_summary.json")
    if os.path.isfile(prof):
        try:
            with open(prof) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    cpu = cpu_reference_run(3, 1, budget_s=40.0) if world == 1 and not args.no_cpu_baseline else None
    h2d = Q * DIM * 4
    d2h = Q * TOPK * (4 + 8)
    kernel_name = "score_select_qs_kernel" if plan.qs else "score_select_kernel"
    line = {
        "metric": "queries/sec, 1Mx768 corpus top-100",
        "value": Q / (ms_step * 1e-3), "unit": "queries/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "implementation": f"{Q} queries x {N_CORPUS} corpus x {DIM}-d, fp32 masters, "
                                     f"bf16 tensor-core first pass k'={plan.kprime}, fp32 rescore; corpus sharded over {world} "
                                     f"GPU(s): bf16 operand AND fp32 master of a rank's own rows only, exact scores of "
                                     f"requested rows returned to the owner of a query; {Q_PER_GPU} queries per GPU per step",
                   "l2": "inputs larger than L2 (bf16 corpus shard %.2f GB + fp32 masters)" % (index.n * DIM * 2 / 1e9),
                   "plan": {"m_tiles": plan.m_tiles, "n_tiles": plan.n_tiles, "stripes": plan.stripes,
                            "units": plan.units, "grid": plan.grid, "query_stationary": int(plan.qs)},
                   "uncertified_queries_after_first_pass_and_rescan": uncertified},
        # both host-buffer entries are timed with their copies inside the timed region; the headline is the
        # faster of the two (normally the double-buffered stream; at N = 2 the pipeline has been seen to
        # fall behind the plain call on some boxes), both numbers are kept
        "e2e": {"value": Q / (min(ms_e2e, ms_e2e_serial) * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": min(ms_e2e, ms_e2e_serial),
                "mode": e2e_mode if ms_e2e <= ms_e2e_serial else "one synchronous host-buffer call per step",
                "pipelined_value": Q / (ms_e2e * 1e-3), "pipelined_ms_per_step": ms_e2e,
                "serial_call_value": Q / (ms_e2e_serial * 1e-3), "serial_call_ms_per_step": ms_e2e_serial,
                "pipelined_ranking_identical_to_serial": pipe_same},
        # counted by the library (qst_launch_count) around the timed region, on rank 0
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak, "traffic": traffic, "kernel": kernel_name,
                     "kernel_ms": k2, "peak_source": peaks["_source"] + " sustained bf16 (kernel timed inside the step loop)",
                     "frac_of_burst_peak": achieved / float(peaks.get("bf16_tflops", FALLBACK_PEAKS["bf16_tflops"]))},
        "parity_sample": parity,
        "clocks": clocks,
    }
    if stage is not None:
        line["stage_ms_rank0"] = stage
    if cpu is not None:
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    line.update(secondary)
    if world == 1:
        for name, fn in (("loss_config2", lambda: loss_config2(dev, peaks)), ("f2_config1", lambda: f2_config1(dev)),
                         ("small_q", lambda: small_q_block(dev, index, peaks))):
            try:
                line[name] = fn()
            except Exception as e:  # secondary measurement must never break the headline line
                line[name] = {"error": repr(e)}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    bad = parity["mismatch"] + sum(v.get("parity_sample", {}).get("mismatch", 0) for v in secondary.values()
                                   if isinstance(v, dict))
    if bad:
        sys.stderr.write(f"bench.py: {bad} sampled queries disagree with the brute-force fp32 ranking\n")
        sys.exit(1)


def replicated_master_block(cm, dev, queries, n0, n1, rank, timed):
    """The r01 partition for comparison: bf16 operand sharded, fp32 master (3 GB) on every rank."""
    import torch
    from qst_b200 import sharded
    full = corpus_rows(0, N_CORPUS, dev)
    corp = sharded.ShardedCorpus(full[n0:n1], N_CORPUS, "cos_sim", comm=cm, full_master=full)
    for _ in range(3):
        corp.topk_owned(queries, TOPK)
    ms, out, _ = timed(lambda: corp.topk_owned(queries, TOPK), 5)
    par = parity_sample(cm, queries, out[0], out[1], full[n0:n1], n0, TOPK)
    Q = Q_PER_GPU * cm.world
    corp.close()
    return {"value": Q / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms, "steps": 5,
            "per_gpu_fp32_master_gb": N_CORPUS * DIM * 4 / 1e9, "parity_sample": par}


def config4_block(cm, dev, rank, world, barrier, max_over_ranks):
    """BASELINE.json config 4: 100 000 queries x 10 000 000 corpus x 768 over `world` GPUs, sharded master."""
    import torch
    from qst_b200 import sharded
    n_total = int(os.environ.get("QST_C4_N", 10_000_000))
    q_total = int(os.environ.get("QST_C4_Q", 100_000))
    q_own = q_total // world
    n0, n1 = sharded.shard_bounds(n_total, world, rank)
    shard = corpus_rows(n0, n1, dev, seed=CORPUS_SEED + 5000)
    corp = sharded.ShardedCorpus(shard, n_total, "cos_sim", comm=cm)
    rows_f32 = corp.index.rows.f32
    del shard
    qgen = torch.Generator(device=dev).manual_seed(14 + 500 + rank)
    own_q = torch.randn(q_own, DIM, generator=qgen, device=dev)
    corp.enable_stage_timing()
    nxt = None if NO_PREFETCH else own_q
    for _ in range(2):
        out = corp.topk_owned(own_q, TOPK, exact="deferred", prefetch=nxt)
    corp.finish_exact()
    barrier()
    steps = 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = corp.topk_owned(own_q, TOPK, exact="deferred", prefetch=nxt)
    corp.finish_exact()            # certificate check of the last step, inside the timed region
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    vals, idx, margin = out
    unc = (~(margin > 0)).sum().to(torch.float64)
    unc = float(cm.all_reduce_max(unc.view(1)))
    par = parity_sample(cm, own_q, vals, idx, rows_f32, n0, TOPK)
    stage = corp.stage_ms()
    mem = torch.cuda.max_memory_allocated(dev) / 2 ** 30
    corp.close()           # the peer-mapped buffers (hints, query gather) are cudaMalloc'ed outside torch's pool
    return {"workload": f"{q_own * world} queries x {n_total} corpus x {DIM}, top-{TOPK}, {world} GPUs, fp32 master sharded",
            "value": q_own * world / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms, "steps": steps,
            "k2_tflops_per_gpu": 2.0 * q_own * world * (n1 - n0) * DIM / (stage["K2"] * 1e-3) / 1e12,
            "stage_ms_rank0": stage, "max_uncertified_on_a_rank": unc, "peak_mem_gib_rank0": mem, "parity_sample": par}


def config5_block(cm, corp, dev, rank, world, barrier, max_over_ranks):
    """BASELINE.json config 5: 1 000 000 queries x the 1M corpus of the headline run, top-100 + the IR metric
    suite (MRR@10, NDCG@10, Recall@{1,10,100}, MAP@100) on the device, 8 relevant documents per query."""
    import torch
    import torch.distributed as dist
    from qst_b200 import metrics
    n = N_CORPUS
    q_total = int(os.environ.get("QST_C5_Q", 1_000_000))
    q_own = q_total // world
    batch = min(q_own, 12_500)
    q_own = (q_own // batch) * batch
    rows = corp.index.rows.f32
    n0 = corp.start
    # query g is a noisy copy of a corpus row of THIS rank's shard; its 8 relevant docs are that row and 7
    # rows spread over the corpus (which it will mostly not retrieve): recall@100 ~ 1/8, mrr ~ 1
    gid = torch.arange(q_own, device=dev, dtype=torch.long)
    loc = (gid * 7919) % rows.shape[0]
    qgen = torch.Generator(device=dev).manual_seed(14 + 600 + rank)
    own_q = rows[loc] + 0.5 * torch.randn(q_own, DIM, generator=qgen, device=dev)
    rel0 = loc + n0
    rel = torch.sort((rel0[:, None] + torch.arange(8, device=dev)[None, :] * (n // 8)) % n, dim=1).values
    rowptr = torch.arange(0, 8 * q_own + 1, 8, device=dev, dtype=torch.long)
    cols = rel.reshape(-1).contiguous()
    ks = [1, 10, 100]

    def run():
        ranked = torch.empty((q_own, TOPK), dtype=torch.long, device=dev)
        bad = 0
        for s0 in range(0, q_own, batch):
            nxt = own_q[s0 + batch:s0 + 2 * batch] if (s0 + batch < q_own and not NO_PREFETCH) else None
            v, i, m = corp.topk_owned(own_q[s0:s0 + batch], TOPK, prefetch=nxt)
            ranked[s0:s0 + batch] = i
            bad = bad + (~(m > 0)).sum()
        return ranked, metrics.per_query_metrics(ranked, rowptr, cols, ks), bad

    run()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ranked, per_q, bad = run()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    gathered = [torch.empty_like(per_q) for _ in range(world)] if rank == 0 else None
    dist.gather(per_q, gathered, dst=0)
    v, i, _ = corp.topk_owned(own_q[:batch], TOPK)
    par = parity_sample(cm, own_q[:batch], v, i, rows, n0, TOPK)
    res = {"workload": f"{q_own * world} queries x {n} corpus x {DIM}, top-{TOPK} + IR metrics on the device, {world} GPUs",
           "value": q_own * world / (ms * 1e-3), "unit": "queries/s", "ms_total": ms, "batch_per_gpu": batch,
           "uncertified_on_rank0": int(bad), "first_hit_is_planted_row_rank0": float((ranked[:, 0] == rel0).float().mean()),
           "parity_sample": par}
    if rank == 0:
        allq = torch.cat(gathered, dim=2).cpu().numpy()
        red = metrics.reduce_like_reference(allq, ks, accuracy_at_k=[1], precision_recall_at_k=[1, 10, 100],
                                            mrr_at_k=[10], ndcg_at_k=[10], map_at_k=[100])
        res["metrics"] = {"mrr@10": red["mrr@k"][10], "ndcg@10": red["ndcg@k"][10], "recall@1": red["recall@k"][1],
                          "recall@10": red["recall@k"][10], "recall@100": red["recall@k"][100], "map@100": red["map@k"][100]}
    return res


def f2_config1(dev):
    """SURVEY.md section 8 f2 at speed: the whole evaluator call at BASELINE config 1 (1000 queries x 10 000 corpus
    x 384) with the script defaults of ir_evauation_script.py:163-173 (k-lists up to 900, three score
    functions), embeddings precomputed, against the CPU oracle evaluator: wall time and value equality."""
    import torch
    import qst_b200
    from oracle import ir_oracle
    q, c, queries, corpus, relevant = qst_b200.synth.ir_eval_set(1000, 10_000, 384)
    table = torch.cat([q, c])
    kw = dict(qst_b200.synth.SCRIPT_DEFAULT_K_LISTS, write_csv=False)
    ev = qst_b200.InformationRetrievalEvaluator(queries, corpus, relevant, score_functions={
        "cos_sim": qst_b200.cos_sim, "dot_score": qst_b200.dot_score, "euclid_score": qst_b200.euclidean_score}, **kw)
    model = qst_b200.synth.TableModel(table.to(dev))
    got = ev.compute_metrices(model)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        got = ev.compute_metrices(model)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    ref = ir_oracle.InformationRetrievalEvaluatorOracle(queries, corpus, relevant, score_functions={
        "cos_sim": ir_oracle.cos_sim, "dot_score": ir_oracle.dot_score, "euclid_score": ir_oracle.euclidean_score}, **kw)
    t0 = time.perf_counter()
    want = ref.compute_metrices(ir_oracle.PrecomputedEmbeddingModel(table))
    cpu_s = time.perf_counter() - t0
    n = same = 0
    for fn in want:
        for metric in want[fn]:
            for k, v in want[fn][metric].items():
                n += 1
                same += float(got[fn][metric][k]) == float(v)
    return {"workload": "InformationRetrievalEvaluator.compute_metrices, 1000 x 10000 x 384, k-lists up to 900, "
                        "cos_sim + dot_score + euclidean_score", "gpu_ms_per_call": ms, "cpu_oracle_s_per_call": cpu_s,
            "metric_values": n, "metric_values_equal": same}


def small_q_block(dev, index, peaks):
    """K2 alone for small query batches against the 1M x 768 corpus: one corpus pass per call, so the floor is
    max(HBM streaming of the bf16 corpus, tensor time of the padded tile rows)."""
    import ctypes as C
    import torch
    from qst_b200 import scoring, _lib
    lib = _lib.load()
    st = _lib.stream_ptr(dev)
    hbm = float(peaks.get("hbm_gbs", FALLBACK_PEAKS["hbm_gbs"])) * 1e9
    tens = float(peaks.get("bf16_tflops", FALLBACK_PEAKS["bf16_tflops"])) * 1e12
    out = {}
    g = torch.Generator(device=dev).manual_seed(77)
    for Qs in (1, 32, 128, 512):
        qs = torch.randn(Qs, DIM, generator=g, device=dev)
        pq = scoring.prepare_rows(qs, True)
        plan = scoring.make_plan(Qs, index.n, DIM, TOPK, 0, "cos_sim")
        ws = scoring._workspace(plan.ws_bytes, dev, "select")
        ts = []
        for i in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.qst_score_select(C.byref(plan), pq.bf16.data_ptr(), index.rows.bf16.data_ptr(), ws.data_ptr(), st))
            b.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(a.elapsed_time(b))
        ms = sorted(ts)[len(ts) // 2]
        rows = plan.m_tiles * plan.rows_per_unit
        floor_hbm = index.n * plan.D_pad * 2 / hbm * 1e3
        floor_tensor = 2.0 * rows * index.n * plan.D_pad / tens * 1e3
        out[str(Qs)] = {"k2_ms": ms, "floor_ms": max(floor_hbm, floor_tensor),
                        "bound": "hbm" if floor_hbm >= floor_tensor else "tensor",
                        "frac_of_floor": max(floor_hbm, floor_tensor) / ms}
    return out


_REAL_STDOUT = None


def _claim_stdout():
    """stdout must carry exactly one JSON line, but libraries loaded later (NCCL's version banner on
    some boxes) write to file descriptor 1 directly: keep a private duplicate of the real stdout for
    the JSON line and point fd 1 at stderr for everything else."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
