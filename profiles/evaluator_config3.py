"""The whole evaluator call at config-3 scale: 10 000 queries x 1 000 000 corpus x 768, cos_sim,
default k-lists (max k = 100), corpus embeddings precomputed and resident on the device (scored in
one pass), query embeddings looked up by the model stand-in.  Wall time per `compute_metrices` call."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qst_b200  # noqa: E402

Q, N, D = 10_000, 1_000_000, 768
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(14)
corpus_emb = torch.cat([torch.randn(125_000, D, generator=g, device=dev) for _ in range(N // 125_000)])
rel0 = (torch.arange(Q, device=dev) * 97) % N
query_emb = corpus_emb[rel0] + 0.5 * torch.randn(Q, D, generator=g, device=dev)
t0 = time.perf_counter()
queries = {f"q{i}": str(i) for i in range(Q)}
corpus = {f"d{i}": "" for i in range(N)}
rel_host = rel0.cpu().tolist()
relevant = {f"q{i}": {f"d{rel_host[i]}", f"d{(rel_host[i] + N // 2) % N}"} for i in range(Q)}
ev = qst_b200.InformationRetrievalEvaluator(queries, corpus, relevant, score_functions={"cos_sim": qst_b200.cos_sim},
                                            write_csv=False)
print(f"evaluator construction (host dicts for 1M documents): {time.perf_counter() - t0:.2f} s")
model = qst_b200.synth.TableModel(query_emb)          # queries are rows 0..Q-1 of the table
ev.compute_metrices(model, corpus_embeddings=corpus_emb)
torch.cuda.synchronize()
ts = []
for _ in range(4):
    t0 = time.perf_counter()
    scores = ev.compute_metrices(model, corpus_embeddings=corpus_emb)
    torch.cuda.synchronize()
    ts.append(time.perf_counter() - t0)
print(f"compute_metrices: min {min(ts) * 1e3:.1f} ms, median {sorted(ts)[1] * 1e3:.1f} ms "
      f"(includes building the bf16 corpus operand: the evaluator keeps no index between calls)")
print({m: {k: round(float(v), 4) for k, v in d.items()} for m, d in scores["cos_sim"].items()})
print("uncertified:", int((ev.last_margins["cos_sim"] <= 0).sum()))
