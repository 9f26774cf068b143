"""Stress loop for a flaky mismatch seen once in test_topk_matches_oracle[euclid_score-1000-10000-384-10]."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qst_b200
from oracle import ir_oracle
dev = torch.device("cuda:0")
Q, N, D, k = 1000, 10000, 384, 10
g = torch.Generator().manual_seed(14 + Q)
q = torch.randn(Q, D, generator=g)
c = torch.randn(N, D, generator=g) * (1.0 + torch.rand(N, 1, generator=g))
want_val, want_idx = ir_oracle.topk_dense(q, c, k, "euclid_score")
qd, cd = q.to(dev), c.to(dev)
bad = 0
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
for it in range(iters):
    index = qst_b200.CorpusIndex(cd, "euclid_score")
    raw = qst_b200.topk(qd, index, k, exact=False)
    res = qst_b200.topk(qd, index, k, exact=True)
    for name, r in (("raw", raw), ("exact", res)):
        dv = (r.values.cpu() - want_val).abs()
        rows = (dv > 2e-6).any(dim=1).nonzero().flatten()
        if rows.numel():
            bad += 1
            flagged = (raw.margin <= 0).cpu()
            print(f"iter {it} {name}: {rows.numel()} rows off, max {float(dv.max()):.3e}; flagged(raw) total {int(flagged.sum())}, "
                  f"off rows flagged: {int(flagged[rows].sum())}; idx equal on off rows: "
                  f"{bool((r.indices.cpu()[rows] == want_idx[rows]).all())}; first rows {rows[:8].tolist()}", flush=True)
            # recompute exactly on device for one off row
            r0 = int(rows[0])
            ids = r.indices[r0]
            d2 = ((qd[r0][None, :] - cd[ids]) ** 2).sum(1)
            print("   torch-gpu scores:", (1 / (1 + d2.sqrt())).cpu().tolist()[:4], " ours:", r.values[r0].cpu().tolist()[:4],
                  " want:", want_val[r0].tolist()[:4], flush=True)
print("iterations", iters, "bad", bad, "flagged last", int((raw.margin <= 0).sum()))

# force the exact re-scan path: no head-room in the first pass
for kp in (16, 32):
    index = qst_b200.CorpusIndex(cd, "euclid_score")
    raw = qst_b200.topk(qd, index, k, kprime=kp, exact=False)
    res = qst_b200.topk(qd, index, k, kprime=kp, exact=True)
    dv = (res.values.cpu() - want_val).abs()
    rows = (dv > 2e-6).any(dim=1).nonzero().flatten()
    flagged = (raw.margin <= 0).cpu()
    print(f"kprime={kp}: flagged {int(flagged.sum())}, rows off after rescan {rows.numel()}, max err {float(dv.max()):.3e}, "
          f"off rows flagged {int(flagged[rows].sum()) if rows.numel() else 0}, idx mismatches {int((res.indices.cpu() != want_idx).sum())}, "
          f"still uncertified {int((res.margin <= 0).sum())}")
    if rows.numel():
        r0 = int(rows[0]); ids = res.indices[r0]
        d2 = ((qd[r0][None, :] - cd[ids]) ** 2).sum(1)
        print("   torch-gpu:", (1 / (1 + d2.sqrt())).cpu().tolist()[:4], " ours:", res.values[r0].cpu().tolist()[:4], " want:", want_val[r0].tolist()[:4])
