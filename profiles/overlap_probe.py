"""Does K3 (HBM-gather-bound) of batch i hide under K2 (tensor-bound, persistent) of batch i+1 when K2 leaves a
few SMs free?  Two streams, config-3 shape; QST_K2_GROUPS = CTA pairs K2 may use (74 = all SMs)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qst_b200  # noqa: E402
from qst_b200 import _lib, scoring  # noqa: E402

Q, N, D, K = 10_000, 1_000_000, 768, 100
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(14)
corpus = torch.cat([torch.randn(N // 8, D, generator=g, device=dev) for _ in range(8)])
queries = torch.randn(Q, D, generator=g, device=dev)
index = qst_b200.CorpusIndex(corpus, "cos_sim")
del corpus
pq = scoring.prepare_rows(queries, True)
lib = _lib.load()
c = index.rows


def run(groups, overlap, iters=30):
    if groups:
        os.environ["QST_K2_GROUPS"] = str(groups)
    else:
        os.environ.pop("QST_K2_GROUPS", None)
    plan = scoring.make_plan(Q, N, D, K, 0, "cos_sim")
    ws = [torch.empty(plan.ws_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    vals = [torch.empty((Q, K), device=dev) for _ in range(2)]
    idx = [torch.empty((Q, K), dtype=torch.int64, device=dev) for _ in range(2)]
    mar = [torch.empty(Q, device=dev) for _ in range(2)]
    sa = torch.cuda.Stream(device=dev)
    sb = torch.cuda.Stream(device=dev, priority=-1) if overlap else sa
    k2_done = [None, None]
    k3_done = [None, None]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sa.wait_event(e0)
    sb.wait_event(e0)
    for i in range(iters + 10):
        s = i % 2
        if i == 10:
            for st in (sa, sb):
                ev = torch.cuda.Event(); ev.record(st)
                torch.cuda.current_stream().wait_event(ev)
            e0.record()
            sa.wait_event(e0); sb.wait_event(e0)
        with torch.cuda.stream(sa):
            if k3_done[s] is not None:
                sa.wait_event(k3_done[s])          # the slot's workspace is free again
            _lib.check(lib.qst_score_select(C.byref(plan), pq.bf16.data_ptr(), c.bf16.data_ptr(), ws[s].data_ptr(), sa.cuda_stream))
            k2_done[s] = torch.cuda.Event(); k2_done[s].record(sa)
        with torch.cuda.stream(sb):
            sb.wait_event(k2_done[s])
            _lib.check(lib.qst_finalize_topk(C.byref(plan), ws[s].data_ptr(), pq.f32.data_ptr(), pq.inv_norm.data_ptr(),
                                             pq.err.data_ptr(), c.f32.data_ptr(), c.inv_norm.data_ptr(), c.stats.data_ptr(), 0,
                                             vals[s].data_ptr(), idx[s].data_ptr(), mar[s].data_ptr(), sb.cuda_stream))
            k3_done[s] = torch.cuda.Event(); k3_done[s].record(sb)
    for st in (sa, sb):
        ev = torch.cuda.Event(); ev.record(st)
        torch.cuda.current_stream().wait_event(ev)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"K2 on {plan.grid} pairs, {'two streams' if overlap else 'one stream '}: {ms:.3f} ms per batch (K2 + K3) -> {Q / ms:.1f}k q/s", flush=True)


for rep in range(2):
    run(0, False)
    run(0, True)
    run(72, True)
    run(70, True)
    run(68, True)
    run(64, True)
