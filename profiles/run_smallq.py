"""ncu target for the small-Q regime: a few K2 launches of 128 queries against the resident 1M x 768
corpus (HBM-bound: one corpus pass per launch).  Prints CUDA-event times of the plain run."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qst_b200  # noqa: E402
from qst_b200 import _lib, scoring  # noqa: E402

N, D, K, Q = 1_000_000, 768, 100, int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(14)
corpus = torch.cat([torch.randn(125_000, D, generator=g, device=dev) for _ in range(N // 125_000)])
index = qst_b200.CorpusIndex(corpus, "cos_sim")
del corpus
queries = torch.randn(Q, D, generator=g, device=dev)
pq = scoring.prepare_rows(queries, True)
plan = scoring.make_plan(Q, N, D, K, 0, "cos_sim")
ws = scoring._workspace(plan.ws_bytes, dev, "select")
lib, st = _lib.load(), _lib.stream_ptr(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(4):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    _lib.check(lib.qst_score_select(C.byref(plan), pq.bf16.data_ptr(), index.rows.bf16.data_ptr(), ws.data_ptr(), st))
    b.record()
    torch.cuda.synchronize()
    print(f"launch {i}: {a.elapsed_time(b):.3f} ms  ({N * D * 2 / a.elapsed_time(b) / 1e6:.0f} GB/s)  ctas={plan.ctas} stripes={plan.stripes}")
