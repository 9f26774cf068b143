"""BASELINE.json config 1 (the reference's own CPU-runnable case) through the whole evaluator call:
1000 queries x 10000 corpus x 384, the three score functions of `--score_functions all` and the
script-default k-lists (ir_evauation_script.py:163-173), embeddings precomputed.  Times the drop-in
evaluator on the GPU and the oracle evaluator (the restated sentence-transformers 2.2.2 path) on
the host CPU, and compares every metric for equality."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import qst_b200  # noqa: E402
from oracle import ir_oracle  # noqa: E402  (the checker and the timed CPU baseline, as in bench.py)

dev = torch.device("cuda:0")
K10 = [5, 10, 20, 30, 40, 50, 100, 200, 500, 900]              # ir_evauation_script.py:163-166
K12 = [1, 3, 5, 10, 20, 30, 40, 50, 100, 200, 500, 900]        # ir_evauation_script.py:167-173
kw = dict(mrr_at_k=K10, ndcg_at_k=K10, accuracy_at_k=K12, precision_recall_at_k=K12, map_at_k=K12, write_csv=False)
q, c, queries, corpus, relevant = qst_b200.synth.ir_eval_set(1000, 10000, 384)
table = torch.cat([q, c])
ours = qst_b200.InformationRetrievalEvaluator(queries, corpus, relevant, score_functions={
    "cos_sim": qst_b200.cos_sim, "dot_score": qst_b200.dot_score, "euclid_score": qst_b200.euclidean_score}, **kw)
ref = ir_oracle.InformationRetrievalEvaluatorOracle(queries, corpus, relevant, score_functions={
    "cos_sim": ir_oracle.cos_sim, "dot_score": ir_oracle.dot_score, "euclid_score": ir_oracle.euclidean_score}, **kw)
model = qst_b200.synth.TableModel(table.to(dev))
ref_model = ir_oracle.PrecomputedEmbeddingModel(table)

ours.compute_metrices(model)                       # warm-up (library load, workspaces)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    got = ours.compute_metrices(model)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
t0 = time.perf_counter()
want = ref.compute_metrices(ref_model)
t_ref = time.perf_counter() - t0

n, bad = 0, []
for fn in want:
    for metric in want[fn]:
        for k, v in want[fn][metric].items():
            n += 1
            if float(got[fn][metric][k]) != float(v):
                bad.append((fn, metric, k, float(got[fn][metric][k]), float(v)))
print(f"drop-in evaluator (GPU, 3 score functions, max k = 900): min {min(ts) * 1e3:.1f} ms, median {sorted(ts)[2] * 1e3:.1f} ms")
print(f"oracle evaluator on {torch.get_num_threads()} CPU threads: {t_ref:.2f} s  ->  {t_ref / min(ts):.0f}x")
print(f"metric values compared for equality: {n}, different: {len(bad)} {bad[:3]}")
print("uncertified queries:", {fn: int((m <= 0).sum()) for fn, m in ours.last_margins.items()})
