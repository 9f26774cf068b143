"""The IR oracle has no reference-side golden vectors ("parity unpinned", see oracle/__init__.py);
these tests check the restatement against values worked out by hand from the published
sentence-transformers 2.2.2 definitions."""
import math
import os

import numpy as np
import torch

from oracle import ir_oracle as io


def _unit(deg, scale=1.0):
    r = math.radians(deg)
    return [scale * math.cos(r), scale * math.sin(r)]


def _tiny_case():
    # corpus angles 0,10,30,90,180,45 degrees with different norms (cos_sim must ignore them)
    corpus_emb = torch.tensor([_unit(0, 2.0), _unit(10, 0.5), _unit(30, 3.0), _unit(90, 1.0),
                               _unit(180, 0.1), _unit(45, 7.0)], dtype=torch.float32)
    query_emb = torch.tensor([_unit(2, 1.5), _unit(85, 0.2), _unit(300, 1.0), _unit(170, 4.0)],
                             dtype=torch.float32)
    table = torch.cat([query_emb, corpus_emb])
    queries = {f"q{i}": str(i) for i in range(4)}
    corpus = {f"d{i}": str(4 + i) for i in range(6)}
    relevant = {"q0": {"d1", "d5"}, "q1": {"d3"}, "q2": set(), "q3": {"d0", "d2", "d4"}}
    return io.PrecomputedEmbeddingModel(table), queries, corpus, relevant


def test_hand_computed_metrics_cos_sim():
    model, queries, corpus, relevant = _tiny_case()
    ev = io.InformationRetrievalEvaluatorOracle(
        queries, corpus, relevant, corpus_chunk_size=4, mrr_at_k=[10], ndcg_at_k=[10],
        accuracy_at_k=[1, 3], precision_recall_at_k=[3], map_at_k=[100],
        score_functions={"cos_sim": io.cos_sim}, write_csv=False)
    assert ev.queries_ids == ["q0", "q1", "q3"]          # q2 has no relevant docs -> dropped
    hits = ev.collect_hits(model)
    assert io.ranked_ids(hits["cos_sim"], 6) == [
        ["d0", "d1", "d2", "d5", "d3", "d4"],
        ["d3", "d5", "d2", "d1", "d0", "d4"],
        ["d4", "d3", "d5", "d2", "d1", "d0"]]
    m = ev.compute_metrics(hits["cos_sim"])
    l2 = math.log2
    ndcg = [(1 / l2(3) + 1 / l2(5)) / (1 / l2(2) + 1 / l2(3)),
            1.0,
            (1 / l2(2) + 1 / l2(5) + 1 / l2(7)) / (1 / l2(2) + 1 / l2(3) + 1 / l2(4))]
    assert m["accuracy@k"] == {1: 2 / 3, 3: 1.0}
    assert m["precision@k"][3] == np.mean([1 / 3, 1 / 3, 1 / 3])
    assert m["recall@k"][3] == np.mean([1 / 2, 1.0, 1 / 3])
    assert m["mrr@k"][10] == (0.5 + 1.0 + 1.0) / 3
    assert abs(m["ndcg@k"][10] - np.mean(ndcg)) < 1e-15
    assert m["map@k"][100] == np.mean([(1 / 2 + 2 / 4) / 2, 1.0, (1 / 1 + 2 / 4 + 3 / 6) / 3])
    # __call__ returns map@max(map_at_k) of the best score function
    assert ev(model) == m["map@k"][100]


def test_chunked_topk_keeps_global_topk():
    model, queries, corpus, relevant = _tiny_case()
    ev = io.InformationRetrievalEvaluatorOracle(
        queries, corpus, relevant, corpus_chunk_size=4, mrr_at_k=[2], ndcg_at_k=[2],
        accuracy_at_k=[1, 2], precision_recall_at_k=[2], map_at_k=[2],
        score_functions={"cos_sim": io.cos_sim}, write_csv=False)
    hits = ev.collect_hits(model)
    # two chunks (4 + 2 docs), top-2 of each kept -> 4 hits per query, global top-2 survives
    assert all(len(h) == 4 for h in hits["cos_sim"])
    assert io.ranked_ids(hits["cos_sim"], 2) == [["d0", "d1"], ["d3", "d5"], ["d4", "d3"]]


def test_dot_and_euclid_differ_from_cos():
    model, queries, corpus, relevant = _tiny_case()
    q = model.table[:4]
    c = model.table[4:]
    torch.testing.assert_close(io.dot_score(q, c), q @ c.T)
    torch.testing.assert_close(io.euclidean_score(q, c), 1 / (1 + torch.cdist(q, c)))
    cs = io.cos_sim(q, c)
    torch.testing.assert_close(cs, (q / q.norm(dim=1, keepdim=True)) @ (c / c.norm(dim=1, keepdim=True)).T)
    # 1-D inputs are promoted to [1, D]
    assert io.cos_sim(q[0], c).shape == (1, 6)
    # zero vector: eps=1e-12 clamp, no NaN
    z = torch.zeros(1, 2)
    assert torch.isfinite(io.cos_sim(z, c)).all()


def test_topk_dense_equals_list_path():
    g = torch.Generator().manual_seed(14)
    q = torch.randn(17, 32, generator=g)
    c = torch.randn(301, 32, generator=g)
    table = torch.cat([q, c])
    model = io.PrecomputedEmbeddingModel(table)
    queries = {str(i): str(i) for i in range(17)}
    corpus = {str(i): str(17 + i) for i in range(301)}
    relevant = {str(i): {str((7 * i) % 301)} for i in range(17)}
    ev = io.InformationRetrievalEvaluatorOracle(queries, corpus, relevant, corpus_chunk_size=100,
                                                mrr_at_k=[10], ndcg_at_k=[10], accuracy_at_k=[1],
                                                precision_recall_at_k=[1], map_at_k=[20],
                                                score_functions={"cos_sim": io.cos_sim}, write_csv=False)
    hits = ev.collect_hits(model)
    want = io.ranked_ids(hits["cos_sim"], 20)
    _, idx = io.topk_dense(q, c, 20, "cos_sim", corpus_chunk_size=100)
    assert [[str(j) for j in row] for row in idx.tolist()] == want


def test_csv_layout(tmp_path):
    model, queries, corpus, relevant = _tiny_case()
    ev = io.InformationRetrievalEvaluatorOracle(
        queries, corpus, relevant, mrr_at_k=[10], ndcg_at_k=[10], accuracy_at_k=[1],
        precision_recall_at_k=[1], map_at_k=[100], name="x",
        score_functions={"dot_score": io.dot_score, "cos_sim": io.cos_sim})
    ev(model, output_path=str(tmp_path), epoch=1, steps=2)
    ev(model, output_path=str(tmp_path), epoch=1, steps=3)
    lines = open(os.path.join(tmp_path, "Information-Retrieval_evaluation_x_results.csv")).read().splitlines()
    assert lines[0].split(",")[:4] == ["epoch", "steps", "cos_sim-Accuracy@1", "cos_sim-Precision@1"]
    assert "dot_score-MAP@100" == lines[0].split(",")[-1]
    assert len(lines) == 3 and lines[1].startswith("1,2,") and lines[2].startswith("1,3,")


def test_metrics_against_independent_implementations():
    """The IR half of the oracle has no reference fixture to pin it (sentence-transformers 2.2.2 is not
    available), so its metric loops are cross-checked against independent code: scikit-learn's
    ``ndcg_score`` (same log2 discount; binary gains; ideal ranking over the whole corpus) and
    closed-form numpy expressions for MRR / MAP / precision / recall / accuracy on full rankings
    without ties."""
    from sklearn.metrics import ndcg_score
    rng = np.random.default_rng(14)
    n_q, n_c, ks = 40, 60, [1, 3, 10, 25]
    queries = {f"q{i}": str(i) for i in range(n_q)}
    corpus = {f"d{j}": str(j) for j in range(n_c)}
    scores = rng.permuted(np.tile(np.linspace(0.0, 1.0, n_c), (n_q, 1)), axis=1)      # no ties
    relevant = {f"q{i}": {f"d{j}" for j in rng.choice(n_c, size=int(rng.integers(1, 9)), replace=False)}
                for i in range(n_q)}
    ev = io.InformationRetrievalEvaluatorOracle(queries, corpus, relevant, mrr_at_k=ks, ndcg_at_k=ks,
                                                accuracy_at_k=ks, precision_recall_at_k=ks, map_at_k=ks,
                                                write_csv=False)
    hits = [[{"corpus_id": f"d{j}", "score": float(scores[i, j])} for j in range(n_c)] for i in range(n_q)]
    got = ev.compute_metrics(hits)
    y_true = np.array([[1.0 if f"d{j}" in relevant[f"q{i}"] else 0.0 for j in range(n_c)] for i in range(n_q)])
    order = np.argsort(-scores, axis=1)
    rel_sorted = np.take_along_axis(y_true, order, axis=1)                              # [q, rank]
    n_rel = y_true.sum(1)
    for k in ks:
        assert math.isclose(got["ndcg@k"][k], ndcg_score(y_true, scores, k=k), rel_tol=1e-12)
        top = rel_sorted[:, :k]
        assert math.isclose(got["accuracy@k"][k], float((top.sum(1) > 0).mean()), rel_tol=1e-12)
        assert math.isclose(got["precision@k"][k], float((top.sum(1) / k).mean()), rel_tol=1e-12)
        assert math.isclose(got["recall@k"][k], float((top.sum(1) / n_rel).mean()), rel_tol=1e-12)
        first = np.where(top.any(1), top.argmax(1) + 1, np.inf)
        assert math.isclose(got["mrr@k"][k], float((1.0 / first).mean()), rel_tol=1e-12)
        prec_at_hit = np.cumsum(top, axis=1) / np.arange(1, top.shape[1] + 1)
        ap = (prec_at_hit * top).sum(1) / np.minimum(k, n_rel)
        assert math.isclose(got["map@k"][k], float(ap.mean()), rel_tol=1e-12)


def test_scores_and_topk_against_independent_implementations():
    """Score functions and the top-k of the oracle against scikit-learn / numpy in float64."""
    from sklearn.metrics.pairwise import cosine_similarity, euclidean_distances
    g = torch.Generator().manual_seed(14)
    q = torch.randn(50, 96, generator=g)
    c = torch.randn(400, 96, generator=g) * (0.5 + torch.rand(400, 1, generator=g))
    q64, c64 = q.double().numpy(), c.double().numpy()
    truth = {"cos_sim": cosine_similarity(q64, c64), "dot_score": q64 @ c64.T,
             "euclid_score": 1.0 / (1.0 + euclidean_distances(q64, c64))}
    fns = {"cos_sim": io.cos_sim, "dot_score": io.dot_score, "euclid_score": io.euclidean_score}
    for name, want in truth.items():
        got = fns[name](q, c).double().numpy()
        scale = np.abs(want).max()
        assert np.abs(got - want).max() <= 2e-6 * max(1.0, scale), name
        vals, idx = io.topk_dense(q, c, 10, name, corpus_chunk_size=150)
        order = np.argsort(-want, axis=1, kind="stable")[:, :10]
        gap = np.take_along_axis(want, np.argsort(-want, axis=1)[:, :11], axis=1)
        clear = (gap[:, :-1] - gap[:, 1:]).min(1) > 1e-5 * max(1.0, scale)       # rows without near-ties
        assert clear.sum() > 25
        assert np.array_equal(idx.numpy()[clear], order[clear]), name
