"""ncu target: ONE launch of the default K2 (query-stationary pair kernel) at config-3 scale."""
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
plan = scoring.make_plan(Q, N, D, K, 0, "cos_sim")
ws = scoring._workspace(plan.ws_bytes, dev, "select")
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    _lib.check(lib.qst_score_select(C.byref(plan), pq.bf16.data_ptr(), index.rows.bf16.data_ptr(), ws.data_ptr(),
                                    _lib.stream_ptr(dev)))
torch.cuda.synchronize()
print("qs" if plan.qs else "classic", plan.ctas, plan.stripes, plan.units)
