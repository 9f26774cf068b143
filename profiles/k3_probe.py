"""K3 (finalize) alone at config 3: time per launch with and without the certificate inputs (without
q_err / c_stats there is no eps, hence no early stop: every one of the k' candidates is rescored)."""
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
corpus = torch.cat([torch.randn(125_000, D, generator=g, device=dev) for _ in range(N // 125_000)])
queries = torch.randn(Q, D, generator=g, device=dev)
index = qst_b200.CorpusIndex(corpus, "cos_sim")
del corpus
pq = scoring.prepare_rows(queries, True)
lib = _lib.load()
st = _lib.stream_ptr(dev)
for kp in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "0,320").split(",")]:
    plan = scoring.make_plan(Q, N, D, K, kp, "cos_sim")
    ws = scoring._workspace(plan.ws_bytes, dev, "select")
    vals = torch.empty((Q, K), dtype=torch.float32, device=dev)
    idx = torch.empty((Q, K), dtype=torch.int64, device=dev)
    margin = torch.empty(Q, dtype=torch.float32, device=dev)
    c = index.rows
    _lib.check(lib.qst_score_select(C.byref(plan), pq.bf16.data_ptr(), c.bf16.data_ptr(), ws.data_ptr(), st))
    ref = None
    for name, with_eps in (("early stop", True), ("all k' rescored", False)):
        ts = []
        for i in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.qst_finalize_topk(C.byref(plan), ws.data_ptr(), pq.f32.data_ptr(), pq.inv_norm.data_ptr(),
                                             pq.err.data_ptr() if with_eps else None, c.f32.data_ptr(),
                                             c.inv_norm.data_ptr(), c.stats.data_ptr() if with_eps else None, 0,
                                             vals.data_ptr(), idx.data_ptr(), margin.data_ptr(), st))
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        same = ""
        if ref is None:
            ref = (vals.clone(), idx.clone())
            same = f", uncertified {int((margin <= 0).sum())}, min margin {float(margin.min()):.5f}"
        else:
            same = f", rankings identical to early stop: {bool((idx == ref[1]).all() and (vals == ref[0]).all())}"
        print(f"k'={plan.kprime} {name}: K3 min {min(ts):.3f} ms median {sorted(ts)[2]:.3f} ms{same}")
