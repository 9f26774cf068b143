"""Small driver for ncu captures: config 3 (10k queries x 1M corpus x 768, k=100), a few steps of
the device-resident hot path.  Prints CUDA-event times so the plain run can be compared with the
launch list (never quote numbers printed while running under ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qst_b200  # noqa: E402

Q, N, D, K = 10_000, int(os.environ.get("QST_PROF_N", 1_000_000)), 768, 100
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(14)
corpus = torch.cat([torch.randn(125_000, D, generator=g, device=dev) for _ in range(N // 125_000)])
queries = torch.randn(Q, D, generator=g, device=dev)
index = qst_b200.CorpusIndex(corpus, "cos_sim")
del corpus
for i in range(steps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    r = qst_b200.topk(queries, index, K)
    b.record()
    torch.cuda.synchronize()
    print(f"step {i}: {a.elapsed_time(b):.3f} ms, uncertified {int((r.margin <= 0).sum())}, k'={r.plan.kprime}")
