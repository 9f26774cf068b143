"""Small query batches against the resident 1M x 768 corpus ('small-Q' row): the pass
is corpus-streaming bound, N*D*2 bytes of bf16 per pass.  Reports K2 alone and the whole topk()."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qst_b200  # noqa: E402
from qst_b200 import _lib, scoring  # noqa: E402

N, D, K = 1_000_000, 768, 100
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(14)
corpus = torch.cat([torch.randn(125_000, D, generator=g, device=dev) for _ in range(N // 125_000)])
index = qst_b200.CorpusIndex(corpus, "cos_sim")
del corpus
lib = _lib.load()
st = _lib.stream_ptr(dev)
peak = 6548.2
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2: evict the corpus between timed passes
for Q in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "1,32,128,256,512,1024").split(",")]:
    queries = torch.randn(Q, D, generator=g, device=dev)
    pq = scoring.prepare_rows(queries, True)
    plan = scoring.make_plan(Q, N, D, K, 0, "cos_sim")
    ws = scoring._workspace(plan.ws_bytes, dev, "select")
    t_k2, t_all = [], []
    for i in range(6):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _lib.check(lib.qst_score_select(C.byref(plan), pq.bf16.data_ptr(), index.rows.bf16.data_ptr(), ws.data_ptr(), st))
        b.record()
        torch.cuda.synchronize()
        t_k2.append(a.elapsed_time(b))
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = qst_b200.topk(queries, index, K)
        b.record()
        torch.cuda.synchronize()
        t_all.append(a.elapsed_time(b))
    k2, al = min(t_k2[1:]), min(t_all[1:])
    gbs = N * D * 2 / (k2 * 1e-3) / 1e9
    print(f"Q={Q:5d} stripes={plan.stripes:3d} units={plan.units:4d} grid={plan.grid:3d}: K2 {k2:.3f} ms = {gbs:6.0f} GB/s "
          f"({100 * gbs / peak:4.1f}% of measured HBM copy peak), topk() {al:.3f} ms, uncertified {int((r.margin <= 0).sum())}")
