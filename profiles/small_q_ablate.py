import ctypes as C, os, sys, torch
sys.path.insert(0, "/root/repo")
import qst_b200
from qst_b200 import _lib, scoring
N, D, K = 1_000_000, 768, 100
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(14)
corpus = torch.cat([torch.randn(125_000, D, generator=g, device=dev) for _ in range(8)])
index = qst_b200.CorpusIndex(corpus, "cos_sim"); del corpus
lib = _lib.load(); st = _lib.stream_ptr(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for Q in (1, 128):
    queries = torch.randn(Q, D, generator=g, device=dev)
    pq = scoring.prepare_rows(queries, True)
    for ctas in ("2", "1"):
        os.environ["QST_SCORE_CTAS"] = ctas
        plan = scoring.make_plan(Q, N, D, K, 0, "cos_sim")
        ws = scoring._workspace(plan.ws_bytes, dev, "select")
        for dbg in ("0", "16", "2", "1", "4"):
            os.environ["QST_SCORE_DEBUG"] = dbg
            ts = []
            for i in range(5):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                _lib.check(lib.qst_score_select(C.byref(plan), pq.bf16.data_ptr(), index.rows.bf16.data_ptr(), ws.data_ptr(), st))
                b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
            print(f"Q={Q} ctas={ctas} stripes={plan.stripes} units={plan.units} grid={plan.grid} debug={dbg}: {min(ts[1:]):.3f} ms")
