"""Per-tile trace of K2 in the small-Q regime (QST_SCORE_DEBUG bit 32): where a cold unit spends its time."""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ["QST_SCORE_DEBUG"] = "32"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qst_b200  # noqa: E402
from qst_b200 import _lib, scoring  # noqa: E402

N, D, K = 1_000_000, 768, 100
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(14)
corpus = torch.cat([torch.randn(125_000, D, generator=g, device=dev) for _ in range(N // 125_000)])
index = qst_b200.CorpusIndex(corpus, "cos_sim")
del corpus
lib = _lib.load()
st = _lib.stream_ptr(dev)
queries = torch.randn(Q, D, generator=g, device=dev)
pq = scoring.prepare_rows(queries, True)
plan = scoring.make_plan(Q, N, D, K, 0, "cos_sim")
ws = scoring._workspace(plan.ws_bytes, dev, "select")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for i in range(3):
    flush.zero_()
    _lib.check(lib.qst_score_select(C.byref(plan), pq.bf16.data_ptr(), index.rows.bf16.data_ptr(), ws.data_ptr(), st))
    torch.cuda.synchronize()
T = 64
n = 160 * T
ns = np.zeros(n, dtype=np.int64)
cnt = np.zeros(n, dtype=np.int32)
thr = np.zeros(n, dtype=np.float32)
_lib.check(lib.qst_debug_read_trace(ns.ctypes.data, cnt.ctypes.data, thr.ctypes.data, n))
ns, cnt, thr = ns.reshape(160, T), cnt.reshape(160, T), thr.reshape(160, T)
tiles = plan.tiles_per_stripe
print(f"Q={Q} ctas={plan.ctas} stripes={plan.stripes} tiles/stripe={tiles} kunit={plan.kunit} cap={plan.cap} grid={plan.grid}")
t_start = ns[:plan.grid * plan.ctas, 0].min()
for cta in (0, 1, plan.grid * plan.ctas // 2, plan.grid * plan.ctas - 1):
    row = ns[cta]
    if row[0] == 0:
        continue
    d = np.diff(row[:tiles + 1]) / 1e3
    print(f"cta {cta}: start +{(row[0] - t_start) / 1e3:.1f} us, per-tile us:", " ".join(f"{x:.1f}" for x in d),
          f"| final compaction {(row[T - 1] - row[tiles]) / 1e3:.1f} us, total {(row[T - 1] - row[0]) / 1e3:.1f} us")
    if tiles < 40 and row[41]:
        print(f"   first tile: accumulator ready +{(row[40] - row[0]) / 1e3:.1f} us, cold pass 1 {(row[41] - row[40]) / 1e3:.1f} us, "
              f"pass 2 {(row[1] - row[41]) / 1e3:.1f} us")
    print("   cnt:", " ".join(str(int(x)) for x in cnt[cta, 1:tiles + 1]))
    print("   thr:", " ".join(f"{x:.3f}" for x in thr[cta, 1:tiles + 1]))
act = [c for c in range(plan.grid * plan.ctas) if ns[c, 0] != 0]
d_all = np.stack([np.diff(ns[c, :tiles + 1]) for c in act]) / 1e3
print("mean per-tile us over CTAs:", " ".join(f"{x:.1f}" for x in d_all.mean(0)))
print("span: first start -> last end", (ns[act, T - 1].max() - t_start) / 1e3, "us")
