"""Distribution of the certificate margin before any re-scan (config-3 data, 20k queries)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qst_b200  # noqa: E402

dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(14 + 100)
full = torch.cat([torch.randn(125_000, 768, generator=gen, device=dev) for _ in range(8)])
qgen = torch.Generator(device=dev).manual_seed(14 + 200)
queries = torch.randn(20_000, 768, generator=qgen, device=dev)
for n_rows in (1_000_000, 500_000):
    index = qst_b200.CorpusIndex(full[:n_rows], "cos_sim")
    for ku in os.environ.get("KUNITS", "16,24,32").split(","):
        os.environ["QST_KUNIT"] = ku
        r = qst_b200.topk(queries, index, 100, exact=False)
        m = r.margin
        fin = m[torch.isfinite(m)]
        print(f"N={n_rows} kunit={r.plan.kunit} stripes={r.plan.stripes}: flagged {int((m <= 0).sum())} of {m.numel()}, "
              f"margin min {float(fin.min()):.5f} p0.1% {float(fin.kthvalue(max(1, fin.numel() // 1000)).values):.5f} "
              f"median {float(fin.median()):.5f}")
