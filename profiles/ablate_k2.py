"""Ablation timings of the K2 kernel alone (QST_SCORE_DEBUG bits; results are invalid when set)."""
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
for ctas in (sys.argv[1] if len(sys.argv) > 1 else "2").split(","):
    os.environ["QST_SCORE_CTAS"] = ctas
    plan = scoring.make_plan(Q, N, D, K, 0, "cos_sim")
    ws = scoring._workspace(plan.ws_bytes, dev, "select")
    for dbg in (sys.argv[2] if len(sys.argv) > 2 else "0,16,2,1").split(","):
        os.environ["QST_SCORE_DEBUG"] = dbg
        ts = []
        for i in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(lib.qst_score_select(C.byref(plan), pq.bf16.data_ptr(), index.rows.bf16.data_ptr(),
                                            ws.data_ptr(), st))
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print(f"ctas={ctas} debug={dbg:>2s} stripes={plan.stripes} units={plan.units}: "
              f"min {min(ts):.3f} ms  median {sorted(ts)[2]:.3f} ms  -> {2 * Q * N * D / min(ts) / 1e9:.0f} TFLOP/s")
