"""Parity of the scoring / top-k / metrics kernels (K1-K4, K6) with the CPU oracle, through the
C ABI.  Bars (BASELINE.json north_star): top-k indices identical to torch fp32 cos_sim+topk except
ties within 1e-6; IR metric values bit-identical given identical rankings."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TIE = 1e-6


def _dev():
    return torch.device("cuda:0")


@pytest.fixture(params=[(2, 1), (2, 0), (1, 0)], ids=["cta_pair_qs", "cta_pair", "single_cta"])
def ctas(request, monkeypatch):
    """The three K2 kernels: cta_group::2 pairs with the query block resident in TMEM (default for
    D_pad <= 768), pairs that re-stage the query tile per corpus tile, and the single-CTA tile."""
    n, qs = request.param
    monkeypatch.setenv("QST_SCORE_CTAS", str(n))
    monkeypatch.setenv("QST_SCORE_QS", str(qs))
    return n


def _fp64_scores(q, c, idx, score):
    """Float64 scores of the listed documents, straight from the definitions (diagnostics)."""
    qq, cc = q.double(), c.double()[idx.clamp_min(0)]          # [Q, k, D]
    if score == "cos_sim":
        return torch.nn.functional.cosine_similarity(qq[:, None, :], cc, dim=2, eps=1e-12)
    if score == "dot_score":
        return (qq[:, None, :] * cc).sum(2)
    return 1 / (1 + (qq[:, None, :] - cc).norm(dim=2))


def assert_euclid_dense_close(got, q, c, oracle):
    """Dense euclid_score matrix: within 2e-6 of the float64 definition, and as close to the CPU oracle
    (torch.cdist's matmul formulation, whose accuracy depends on the host CPU's matmul path: a few GPU
    boxes' hosts are off by up to 3e-5) as the oracle is to float64."""
    truth = 1 / (1 + torch.cdist(q.double(), c.double(), compute_mode="donot_use_mm_for_euclid_dist"))
    oracle_off = float((oracle.double() - truth).abs().max())
    if oracle_off > 2e-6:
        print(f"CPU ORACLE DISAGREES WITH FLOAT64 by {oracle_off:.3e} (euclid_score, dense)")
    torch.testing.assert_close(got.double(), truth, rtol=0, atol=2e-6)
    torch.testing.assert_close(got, oracle, rtol=0, atol=max(2e-6, 2 * oracle_off))


ARBITRATIONS = []   # every time the float64 arbitration below let a run continue (euclid_score only)


def assert_same_ranking(got_idx, got_val, want_idx, want_val, what="", truth=None, scale=1.0):
    """Index-for-index equality, except inside groups of reference scores within TIE.
    ``scale``: scores are compared after division by it (dot_score: ties within 1e-6 RELATIVE to the
    largest score of the batch, the scale at which fp32 dot products of unnormalised rows agree).
    ``truth`` = (q, c, "euclid_score") enables the float64 arbitration of a VALUE mismatch, for
    euclid_score only: the reference computes it through torch.cdist's matmul formulation, which on the
    host is less accurate than summing (q-c)^2 directly; everywhere else a value mismatch fails."""
    got_idx, got_val = got_idx.cpu(), got_val.cpu() / scale
    want_val = want_val / scale
    if truth is not None and truth[2] == "euclid_score" and not torch.allclose(got_val, want_val, rtol=0, atol=2e-6):
        # If OUR values are off, fail with the evidence; if ours agree with float64 and the oracle's do
        # not (it depends on the host CPU's matmul path: a few of the GPU boxes' hosts are off by up to
        # 3e-5, enough to reorder neighbours), say so loudly, record it, and go on with the ranking the
        # DEFINITION gives in float64 for those rows.
        q, c, score = truth
        rows = ((got_val - want_val).abs() > 2e-6).any(dim=1).nonzero().flatten()
        t_got = _fp64_scores(q[rows], c, got_idx[rows], score)
        t_want = _fp64_scores(q[rows], c, want_idx[rows], score)
        ours_off = float((got_val[rows].double() - t_got).abs().max())
        oracle_off = float((want_val[rows].double() - t_want).abs().max())
        report = (f"{what}: {rows.numel()} rows differ (first {rows[:10].tolist()}); max |ours - fp64| = "
                  f"{ours_off:.3e}, max |oracle - fp64| = {oracle_off:.3e}")
        if ours_off > 2e-6 or oracle_off <= 2e-6:
            raise AssertionError(report)
        import warnings
        warnings.warn("CPU ORACLE DISAGREES WITH FLOAT64 (ours agrees): " + report)
        print("CPU ORACLE DISAGREES WITH FLOAT64 (ours agrees): " + report)
        ARBITRATIONS.append(report)
        d64 = torch.cdist(q[rows].double(), c.double(), compute_mode="donot_use_mm_for_euclid_dist")
        v64, i64 = (1 / (1 + d64)).topk(want_idx.shape[1], dim=1)
        want_val, want_idx = want_val.clone(), want_idx.clone()
        want_val[rows], want_idx[rows] = v64.float(), i64
    torch.testing.assert_close(got_val, want_val, rtol=0, atol=2e-6, msg=lambda m: f"{what} scores: {m}")
    mism = (got_idx != want_idx)
    if not mism.any():
        return 0
    n_tie = 0
    for q, j in mism.nonzero().tolist():
        # the document we emitted at rank j must have a reference score within TIE of the
        # reference's rank-j score (a tie swap), and must be part of the reference's top-k or tie
        # with its last element
        s_ref = float(want_val[q, j])
        where = (want_idx[q] == got_idx[q, j]).nonzero()
        if where.numel():
            s_doc = float(want_val[q, where[0, 0]])
        else:
            s_doc = float(got_val[q, j])
            assert abs(s_doc - float(want_val[q, -1])) <= 2 * TIE, f"{what}: q={q} rank={j} not in reference top-k"
        assert abs(s_doc - s_ref) <= 2 * TIE, f"{what}: q={q} rank={j}: {s_doc} vs {s_ref} is not a tie"
        n_tie += 1
    return n_tie


@pytest.mark.parametrize("Q,N,D", [(100, 700, 384), (128, 256, 64), (1, 1, 8), (257, 513, 100), (300, 5000, 768)])
def test_tensorcore_scores_match_bf16_matmul(Q, N, D, ctas):
    """K1 + K2 (dense debug epilogue): TMA/UMMA descriptors, swizzle, TMEM layout, ragged tiles."""
    import qst_b200
    from qst_b200 import scoring
    g = torch.Generator().manual_seed(14)
    q = torch.randn(Q, D, generator=g)
    c = torch.randn(N, D, generator=g)
    pq = scoring.prepare_rows(q.to(_dev()), True)
    pc = scoring.prepare_rows(c.to(_dev()), True)
    # K1 against torch
    qn = torch.nn.functional.normalize(q, p=2, dim=1)
    assert pq.bf16.shape[1] % 64 == 0
    torch.testing.assert_close(pq.bf16[:, :D].float().cpu(), qn.bfloat16().float(), rtol=0, atol=2 ** -8)
    assert float(pq.bf16[:, D:].float().abs().max() if pq.bf16.shape[1] > D else 0.0) == 0.0
    torch.testing.assert_close(pq.inv_norm.cpu(), 1 / q.norm(dim=1).clamp_min(1e-12), rtol=1e-6, atol=0)
    torch.testing.assert_close(pq.sq_norm.cpu(), (q * q).sum(1), rtol=1e-5, atol=0)
    err = (pq.bf16[:, :D].float() - (q.to(_dev()) * pq.inv_norm[:, None])).norm(dim=1)
    torch.testing.assert_close(pq.err, err, rtol=1e-3, atol=1e-7)
    # K2 against an fp32 matmul of the very same bf16 operands (products exact, fp32 accumulate)
    got = scoring.dense_tensorcore_scores(pq.bf16, pc.bf16)
    want = pq.bf16.float() @ pc.bf16.float().T
    torch.testing.assert_close(got, want, rtol=0, atol=2e-5)


def _oracle_topk(q, c, k, score="cos_sim", chunk=50000):
    from oracle import ir_oracle
    return ir_oracle.topk_dense(q, c, k, score, corpus_chunk_size=chunk)


@pytest.mark.parametrize("Q,N,D,k", [(1000, 10000, 384, 10), (64, 3000, 768, 100), (200, 20000, 96, 100),
                                     (5, 40, 16, 10), (130, 257, 33, 7)])
@pytest.mark.parametrize("score", ["cos_sim", "dot_score", "euclid_score"])
def test_topk_matches_oracle(Q, N, D, k, score, ctas):
    import qst_b200
    g = torch.Generator().manual_seed(14 + Q)
    q = torch.randn(Q, D, generator=g)
    c = torch.randn(N, D, generator=g) * (1.0 + torch.rand(N, 1, generator=g))   # varied norms
    want_val, want_idx = _oracle_topk(q, c, k, score)
    index = qst_b200.CorpusIndex(c.to(_dev()), score)
    res = qst_b200.topk(q.to(_dev()), index, k)
    assert res.plan.ctas == ctas
    # dot_score: unnormalised rows, so "ties within 1e-6" is meant relative to the batch's largest score
    scale = float(want_val.abs().max()) if score == "dot_score" else 1.0
    assert_same_ranking(res.indices, res.values, want_idx, want_val, f"{score} {Q}x{N}x{D} k={k}",
                        truth=(q, c, score), scale=scale)
    assert bool((res.margin > 0).all()), "every query must end certified (after the exact re-scan if needed)"


def test_topk_large_k_and_small_corpus():
    """k up to the script default 900 (ir_evauation_script.py:163-173); corpus smaller than k."""
    import qst_b200
    g = torch.Generator().manual_seed(3)
    q = torch.randn(20, 64, generator=g)
    c = torch.randn(9000, 64, generator=g)
    want_val, want_idx = _oracle_topk(q, c, 900)
    res = qst_b200.topk(q.to(_dev()), qst_b200.CorpusIndex(c.to(_dev())), 900)
    assert_same_ranking(res.indices, res.values, want_idx, want_val, "k=900")
    # N < k: the reference returns min(max_k, len(chunk)) hits; we pad with -1 / -inf
    res = qst_b200.topk(q.to(_dev()), qst_b200.CorpusIndex(c[:6].to(_dev())), 10)
    want_val, want_idx = _oracle_topk(q, c[:6], 10)
    assert_same_ranking(res.indices[:, :6], res.values[:, :6], want_idx, want_val, "N<k")
    assert bool((res.indices[:, 6:] == -1).all()) and bool(torch.isinf(res.values[:, 6:]).all())


def test_topk_clustered_near_ties_uses_certificate():
    """corpus = centroid + 0.3*noise: near-ties much denser than the bf16 error -> the certificate
    must flag and the exact re-scan must restore the exact ranking."""
    import qst_b200
    q = qst_b200.synth.clustered_embeddings(64, 128, 5, n_centroids=8, noise=0.02)
    c = qst_b200.synth.clustered_embeddings(20000, 128, 6, n_centroids=8, noise=0.02)
    want_val, want_idx = _oracle_topk(q, c, 100)
    index = qst_b200.CorpusIndex(c.to(_dev()))
    raw = qst_b200.topk(q.to(_dev()), index, 100, exact=False)
    res = qst_b200.topk(q.to(_dev()), index, 100, exact=True)
    assert_same_ranking(res.indices, res.values, want_idx, want_val, "clustered")
    assert bool((res.margin > 0).all())
    # informational: how many queries the first pass could not certify
    print("uncertified after bf16 pass:", int((raw.margin <= 0).sum()), "of", raw.margin.numel())


def test_exact_rescan_repairs_a_deliberately_starved_first_pass():
    """kprime == k leaves no head-room: wherever the bf16 order differs from the fp32 order the
    certificate must fire, and the re-scan must repair the list."""
    import qst_b200
    g = torch.Generator().manual_seed(9)
    q = torch.randn(50, 256, generator=g)
    c = torch.randn(50000, 256, generator=g)
    want_val, want_idx = _oracle_topk(q, c, 32)
    index = qst_b200.CorpusIndex(c.to(_dev()))
    res = qst_b200.topk(q.to(_dev()), index, 32, kprime=32, exact=True)
    assert_same_ranking(res.indices, res.values, want_idx, want_val, "starved")


def test_chunk_merge_and_evaluator_metrics_bit_identical():
    """BASELINE.json config 1 through the drop-in evaluator vs the oracle evaluator."""
    import qst_b200
    from oracle import ir_oracle
    for use_part in (True, False):
        q, c, queries, corpus, relevant = qst_b200.synth.ir_eval_set(1000, 10000, 384, use_part_pos=use_part)
        table = torch.cat([q, c])
        kw = dict(corpus_chunk_size=3000, mrr_at_k=[10], ndcg_at_k=[10], accuracy_at_k=[1, 3, 5, 10],
                  precision_recall_at_k=[1, 3, 5, 10], map_at_k=[10], write_csv=False)
        ev = qst_b200.InformationRetrievalEvaluator(queries, corpus, relevant, score_functions={
            "cos_sim": qst_b200.cos_sim, "dot_score": qst_b200.dot_score}, **kw)
        ref = ir_oracle.InformationRetrievalEvaluatorOracle(queries, corpus, relevant, score_functions={
            "cos_sim": ir_oracle.cos_sim, "dot_score": ir_oracle.dot_score}, **kw)
        model = qst_b200.synth.TableModel(table.to(_dev()))
        ranked = ev.rank(model)
        ref_model = ir_oracle.PrecomputedEmbeddingModel(table)
        hits = ref.collect_hits(ref_model)
        identical = True
        for fn in ("cos_sim", "dot_score"):
            want_ids = ir_oracle.ranked_ids(hits[fn], 10)
            got_rows = ranked[fn].indices.cpu().tolist()
            got_ids = [[f"d{j}" for j in row] for row in got_rows]
            if got_ids != want_ids:      # only legal difference: swaps of scores tied within 1e-6
                identical = False
                want_val, want_idx = _oracle_topk(q, c, 10, fn, chunk=3000)
                assert_same_ranking(ranked[fn].indices, ranked[fn].values, want_idx, want_val, fn,
                                    scale=float(want_val.abs().max()) if fn == "dot_score" else 1.0)
            # metrics are bit-identical GIVEN the ranking: feed our ranking to the reference loops
            own_hits = [[{"corpus_id": f"d{j}", "score": float(-r)} for r, j in enumerate(row)] for row in got_rows]
            want_m = ref.compute_metrics(own_hits)
            got_m = ev.compute_metrics_from_ranking(ranked[fn].indices)
            for metric in want_m:
                for k, v in want_m[metric].items():
                    assert float(got_m[metric][k]) == float(v), (fn, metric, k)
        if identical:                    # then the whole evaluator call is bit-identical too
            got = ev.compute_metrices(model)
            want = ref.compute_metrices(ref_model)
            for fn in want:
                for metric in want[fn]:
                    for k, v in want[fn][metric].items():
                        assert float(got[fn][metric][k]) == float(v), (fn, metric, k)
            assert ev(model) == ref(ref_model)
        print("rankings identical to the oracle:", identical)


def test_metrics_kernel_bit_identical_on_random_rankings():
    """K4 alone: random rankings with hits, misses, relevant ids outside the corpus, k > list length."""
    import qst_b200
    from qst_b200 import metrics
    from oracle import ir_oracle
    rng = np.random.default_rng(14)
    n_q, n_c, K = 257, 500, 100
    queries = {f"q{i}": str(i) for i in range(n_q)}
    corpus = {f"d{i}": str(i) for i in range(n_c)}
    relevant = {}
    for i in range(n_q):
        rel = {f"d{j}" for j in rng.choice(n_c, size=rng.integers(1, 12), replace=False)}
        if i % 7 == 0:
            rel.add("not-in-corpus")            # counts in the denominators, can never be hit
        relevant[f"q{i}"] = rel
    k_lists = dict(mrr_at_k=[5, 10, 200], ndcg_at_k=[5, 10, 100, 200], accuracy_at_k=[1, 3, 5, 10],
                   precision_recall_at_k=[1, 3, 5, 10, 100, 200], map_at_k=[1, 10, 100, 200])
    ref = ir_oracle.InformationRetrievalEvaluatorOracle(queries, corpus, relevant, write_csv=False,
                                                        score_functions={"cos_sim": ir_oracle.cos_sim}, **k_lists)
    ev = qst_b200.InformationRetrievalEvaluator(queries, corpus, relevant, write_csv=False,
                                                score_functions={"cos_sim": qst_b200.cos_sim}, **k_lists)
    ranked = np.stack([rng.permutation(n_c)[:K] for _ in range(n_q)])
    hits = [[{"corpus_id": f"d{j}", "score": float(K - r)} for r, j in enumerate(row)] for row in ranked]
    want = ref.compute_metrics(hits)
    got = ev.compute_metrics_from_ranking(torch.from_numpy(ranked).to(_dev()))
    for metric in want:
        for k, v in want[metric].items():
            assert float(got[metric][k]) == float(v), (metric, k, got[metric][k], v)


def test_merge_kernel_against_sort():
    from qst_b200 import sharded
    g = torch.Generator().manual_seed(1)
    G, Q, k = 5, 300, 37
    vals = torch.randn(G, Q, k, generator=g)
    vals[2, :, 5:9] = vals[3, :, 5:9]                       # cross-list score ties -> lower id first
    vals = vals.sort(dim=2, descending=True).values
    idx = torch.stack([torch.randperm(100000, generator=g)[:Q * k].view(Q, k) + 100000 * s for s in range(G)])
    idx[4, :, 30:] = -1                                      # a short list (shard smaller than k)
    vals[4, :, 30:] = float("-inf")
    mv, mi = sharded.merge_topk(vals.to(_dev()), idx.to(_dev()))
    flat_v = vals.permute(1, 0, 2).reshape(Q, G * k)
    flat_i = idx.permute(1, 0, 2).reshape(Q, G * k)
    key_i = torch.where(flat_i < 0, torch.full_like(flat_i, 2 ** 62), flat_i)
    order = torch.argsort(key_i, dim=1, stable=True)
    flat_v, flat_i = flat_v.gather(1, order), flat_i.gather(1, order)
    order = torch.argsort(flat_v, dim=1, descending=True, stable=True)[:, :k]
    assert torch.equal(mv.cpu(), flat_v.gather(1, order))
    assert torch.equal(mi.cpu(), flat_i.gather(1, order))


def test_single_process_shard_emulation_matches_global():
    """G shards on one GPU, merged on device == the unsharded answer."""
    import qst_b200
    from qst_b200 import sharded
    g = torch.Generator().manual_seed(2)
    q = torch.randn(70, 128, generator=g)
    c = torch.randn(10007, 128, generator=g)
    want_val, want_idx = _oracle_topk(q, c, 50)
    parts_v, parts_i = [], []
    for r in range(4):
        s, e = sharded.shard_bounds(c.shape[0], 4, r)
        res = qst_b200.topk(q.to(_dev()), qst_b200.CorpusIndex(c[s:e].to(_dev()), idx_offset=s), 50)
        parts_v.append(res.values)
        parts_i.append(res.indices)
    mv, mi = sharded.merge_topk(torch.stack(parts_v), torch.stack(parts_i))
    assert_same_ranking(mi, mv, want_idx, want_val, "sharded")


def test_host_buffer_entry_and_dense_callable():
    import qst_b200
    from oracle import ir_oracle
    g = torch.Generator().manual_seed(4)
    q = torch.randn(33, 96, generator=g)
    c = torch.randn(500, 96, generator=g)
    index = qst_b200.CorpusIndex(c.to(_dev()))
    vals, idx = qst_b200.topk_host(q.pin_memory(), index, 10)
    want_val, want_idx = _oracle_topk(q, c, 10)
    assert not vals.is_cuda and vals.is_pinned()
    assert_same_ranking(idx, vals, want_idx, want_val, "host entry")
    dense = qst_b200.cos_sim(q.to(_dev()), c.to(_dev()))
    torch.testing.assert_close(dense.cpu(), ir_oracle.cos_sim(q, c), rtol=0, atol=2e-6)
    with pytest.raises(qst_b200.QstError):
        qst_b200.topk(q, index, 10)            # CPU queries: no fallback


def test_full_size_config3_properties():
    """BASELINE.json config 3 at full size (10k x 1M x 768, k=100) through size-independent
    properties: (a) a sample of queries against the CPU oracle on the full corpus, (b) planted
    self-matches come back first with score 1, (c) every list is sorted and duplicate-free,
    (d) every query ends certified, (e) a 4-way sharded run merges to the same answer."""
    import qst_b200
    from qst_b200 import sharded
    Q, N, D, k = 10_000, 1_000_000, 768, 100
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(14)
    corpus = torch.cat([torch.randn(125_000, D, generator=g, device=dev) for _ in range(N // 125_000)])
    queries = torch.randn(Q, D, generator=g, device=dev)
    planted = torch.arange(0, 64, device=dev) * 15_000 + 7            # 64 queries are corpus rows (scaled)
    queries[:64] = corpus[planted] * 3.0
    index = qst_b200.CorpusIndex(corpus, "cos_sim")
    res = qst_b200.topk(queries, index, k)
    vals, idx = res.values, res.indices
    assert bool((res.margin > 0).all())                                                   # (d)
    assert bool((vals[:, :-1] >= vals[:, 1:]).all())                                      # (c)
    assert bool((idx.sort(dim=1).values.diff(dim=1) > 0).all()) and bool((idx >= 0).all() and (idx < N).all())
    assert torch.equal(idx[:64, 0], planted) and bool((vals[:64, 0] - 1).abs().max() < 1e-5)  # (b)
    sample = torch.cat([torch.arange(0, 8), torch.arange(5000, 5008), torch.arange(Q - 8, Q)])
    want_val, want_idx = _oracle_topk(queries[sample].cpu(), corpus.cpu(), k)             # (a)
    assert_same_ranking(idx[sample.to(dev)], vals[sample.to(dev)], want_idx, want_val, "config 3 sample")
    pv, pi = [], []                                                                       # (e)
    sub = torch.arange(0, Q, 25, device=dev)
    for r in range(4):
        s, e = sharded.shard_bounds(N, 4, r)
        part = qst_b200.topk(queries[sub], qst_b200.CorpusIndex(corpus[s:e], "cos_sim", idx_offset=s), k)
        pv.append(part.values)
        pi.append(part.indices)
    mv, mi = sharded.merge_topk(torch.stack(pv), torch.stack(pi))
    assert_same_ranking(mi, mv, idx[sub].cpu(), vals[sub].cpu(), "4 shards vs 1")


def test_metrics_at_scale_match_reference_loops():
    """Config-5 style metric suite (MRR@10, NDCG@10, Recall@{1,10,100}, MAP@100) on 50k queries:
    K4 + host reductions vs the reference's Python loops fed with the same ranking."""
    import qst_b200
    from oracle import ir_oracle
    rng = np.random.default_rng(5)
    n_q, n_c, K = 50_000, 200_000, 100
    ranked = rng.integers(0, n_c, size=(n_q, K), dtype=np.int64)
    rel_first = rng.integers(0, n_c, size=(n_q, 8), dtype=np.int64)
    # make ~half of the queries have hits at random ranks
    rows = rng.choice(n_q, n_q // 2, replace=False)
    ranked[rows, rng.integers(0, K, size=rows.size)] = rel_first[rows, 0]
    ranked[rows[: n_q // 4], rng.integers(0, 10, size=n_q // 4)] = rel_first[rows[: n_q // 4], 1]
    queries = {str(i): str(i) for i in range(n_q)}
    relevant = {str(i): {str(j) for j in rel_first[i]} for i in range(n_q)}
    k_lists = dict(mrr_at_k=[10], ndcg_at_k=[10], accuracy_at_k=[1, 10], precision_recall_at_k=[1, 10, 100],
                   map_at_k=[100])
    ref = ir_oracle.InformationRetrievalEvaluatorOracle(queries, {}, relevant, write_csv=False,
                                                        score_functions={"cos_sim": ir_oracle.cos_sim}, **k_lists)
    hits = [[{"corpus_id": str(j), "score": float(K - r)} for r, j in enumerate(row)] for row in ranked.tolist()]
    want = ref.compute_metrics(hits)
    from qst_b200 import metrics
    rel_pos = [sorted(int(j) for j in rel_first[i]) for i in range(n_q)]
    rel_pos = [sorted(set(r)) for r in rel_pos]
    rowptr, cols = metrics.relevance_csr(rel_pos, _dev())
    ks = sorted({k for v in k_lists.values() for k in v})
    per_query = metrics.per_query_metrics(torch.from_numpy(ranked).to(_dev()), rowptr, cols, ks).cpu().numpy()
    got = metrics.reduce_like_reference(per_query, ks, k_lists["accuracy_at_k"], k_lists["precision_recall_at_k"],
                                        k_lists["mrr_at_k"], k_lists["ndcg_at_k"], k_lists["map_at_k"])
    for metric in want:
        for k, v in want[metric].items():
            assert float(got[metric][k]) == float(v), (metric, k, got[metric][k], v)


def test_euclidean_score_operands_and_script_default_score_set():
    """euclidean_score (models/evaluators.py:392-405) on the tensor-core path: the augmented bf16
    operands reproduce 2 q.c - ||c||^2, and the evaluator runs the script's default score set
    ('all' = cos_sim, dot_score, euclid_score, ir_evauation_script.py:70,179-183)."""
    import qst_b200
    from qst_b200 import scoring, _lib
    from oracle import ir_oracle
    g = torch.Generator().manual_seed(21)
    q = torch.randn(40, 100, generator=g) * 2.0
    c = torch.randn(900, 100, generator=g) * (0.5 + torch.rand(900, 1, generator=g))
    pq = scoring.prepare_rows(q.to(_dev()), _lib.QST_PREP_EUCLID_QUERY)
    pc = scoring.prepare_rows(c.to(_dev()), _lib.QST_PREP_EUCLID_CORPUS)
    assert pq.bf16.shape[1] == 128 and pc.bf16.shape[1] == 128          # 100 + 3 -> 128
    keys = scoring.dense_tensorcore_scores(pq.bf16, pc.bf16).cpu()
    want = 2 * (q.bfloat16().float() @ c.bfloat16().float().T) - (c * c).sum(1)[None, :]
    torch.testing.assert_close(keys, want, rtol=0, atol=2e-3 * float(want.abs().max()) / 100)
    dense = qst_b200.euclidean_score(q.to(_dev()), c.to(_dev()))
    assert_euclid_dense_close(dense.cpu(), q, c, ir_oracle.euclidean_score(q, c))
    # evaluator with the three score functions of the reference script
    qq, cc, queries, corpus, relevant = qst_b200.synth.ir_eval_set(200, 3000, 64)
    table = torch.cat([qq, cc])
    kw = dict(mrr_at_k=[10], ndcg_at_k=[10], accuracy_at_k=[1, 3, 5, 10], precision_recall_at_k=[1, 3, 5, 10],
              map_at_k=[100], write_csv=False)
    ev = qst_b200.InformationRetrievalEvaluator(queries, corpus, relevant, score_functions={
        "cos_sim": qst_b200.cos_sim, "dot_score": qst_b200.dot_score, "euclid_score": qst_b200.euclidean_score}, **kw)
    ref = ir_oracle.InformationRetrievalEvaluatorOracle(queries, corpus, relevant, score_functions={
        "cos_sim": ir_oracle.cos_sim, "dot_score": ir_oracle.dot_score, "euclid_score": ir_oracle.euclidean_score}, **kw)
    got = ev.compute_metrices(qst_b200.synth.TableModel(table.to(_dev())))
    want_m = ref.compute_metrices(ir_oracle.PrecomputedEmbeddingModel(table))
    for fn in want_m:
        for metric in want_m[fn]:
            for k, v in want_m[fn][metric].items():
                assert float(got[fn][metric][k]) == float(v), (fn, metric, k, got[fn][metric][k], v)
    assert all(bool((m > 0).all()) for m in ev.last_margins.values())


def test_candidate_exchange_emulated_on_one_gpu():
    """Sharded retrieval with candidate exchange (qst_select_candidates -> all-to-all ->
    qst_finalize_lists), the G shards emulated on one GPU: must equal the unsharded oracle."""
    import ctypes as C
    import qst_b200
    from qst_b200 import scoring, sharded, _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(8)
    Q, N, D, k, G = 150, 30011, 96, 50, 4
    q = torch.randn(Q, D, generator=g)
    c = torch.randn(N, D, generator=g) * (1 + torch.rand(N, 1, generator=g))
    dev = _dev()
    for score in ("cos_sim", "dot_score", "euclid_score"):
        want_val, want_idx = _oracle_topk(q, c, k, score)
        pq = scoring.prepare_rows(q.to(dev), scoring.QUERY_PREP[score])
        master = scoring.prepare_rows(c.to(dev), scoring.CORPUS_PREP[score], want_bf16=False)
        lists, kprime, m = [], None, None
        for r in range(G):
            s, e = sharded.shard_bounds(N, G, r)
            index = qst_b200.CorpusIndex(c[s:e].to(dev), score, idx_offset=s)
            plan = scoring.make_plan(Q, e - s, D, k, 0, score)
            kprime = plan.kprime
            m = sharded.candidates_per_shard(kprime, G)
            ws = torch.empty(plan.ws_bytes, dtype=torch.uint8, device=dev)
            _lib.check(lib.qst_score_select(C.byref(plan), pq.bf16.data_ptr(), index.rows.bf16.data_ptr(),
                                            ws.data_ptr(), _lib.stream_ptr(dev)))
            out = torch.empty((Q, m + 1, 2), dtype=torch.int32, device=dev)
            _lib.check(lib.qst_select_candidates(C.byref(plan), ws.data_ptr(), m, s, out.data_ptr(), _lib.stream_ptr(dev)))
            lists.append(out)
            # list sanity: trailer count <= m, ids inside the shard
            cnt = out[:, m, 1]
            assert int(cnt.max()) <= m and int(cnt.min()) > 0
            ids = out[:, :m, 1].long() & 0xffffffff
            valid = torch.arange(m, device=dev)[None, :] < cnt[:, None]
            assert bool(((ids >= s) & (ids < e))[valid].all())
        recv = torch.stack(lists)                          # what the owner of all queries would receive
        vals = torch.empty((Q, k), dtype=torch.float32, device=dev)
        idx = torch.empty((Q, k), dtype=torch.int64, device=dev)
        margin = torch.empty(Q, dtype=torch.float32, device=dev)
        scratch = torch.empty(lib.qst_finalize_lists_scratch_bytes(Q, G), dtype=torch.uint8, device=dev)
        cos = score == "cos_sim"
        _lib.check(lib.qst_finalize_lists(Q, G, m, k, kprime, scoring.SCORE_CODES[score], D, recv.data_ptr(),
                                          pq.f32.data_ptr(), pq.inv_norm.data_ptr() if cos else None,
                                          pq.err.data_ptr(), master.f32.data_ptr(),
                                          master.inv_norm.data_ptr() if cos else None, master.stats.data_ptr(),
                                          vals.data_ptr(), idx.data_ptr(), margin.data_ptr(), scratch.data_ptr(),
                                          _lib.stream_ptr(dev)))
        scale = float(want_val.abs().max()) if score == "dot_score" else 1.0
        assert_same_ranking(idx, vals, want_idx, want_val, f"candidate exchange {score}", scale=scale)
        assert bool((margin > 0).all()), score


@pytest.mark.parametrize("Q", [3, 100, 300])
def test_topk_heavy_ties_small_and_large_batches(Q):
    """A corpus of 40 distinct rows each repeated 250 times: every score is tied 250-fold, so units end
    with far more than kunit entries above any threshold (the selecting compaction after the final
    filter) and the ranking must break ties towards the lower corpus position."""
    import qst_b200
    g = torch.Generator().manual_seed(77)
    base = torch.randn(40, 128, generator=g)
    c = base.repeat(250, 1)                       # row i = base[i % 40]
    q = base[torch.arange(Q) % 40] + 0.05 * torch.randn(Q, 128, generator=g)
    k = 20
    res = qst_b200.topk(q.to(_dev()), qst_b200.CorpusIndex(c.to(_dev())), k)
    want_val, _ = _oracle_topk(q, c, k)
    torch.testing.assert_close(res.values.cpu(), want_val, rtol=0, atol=2e-6)
    idx = res.indices.cpu()
    # the 20 best are the 20 lowest positions of the query's own base row: i, i+40, i+80, ...
    want_idx = (torch.arange(Q) % 40)[:, None] + 40 * torch.arange(k)[None, :]
    assert torch.equal(idx, want_idx)
    assert bool((res.margin > 0).all())


def test_host_pipeline_matches_synchronous_calls():
    """HostTopkPipeline: double-buffered host-buffer calls return what topk_host returns, batch by
    batch, whatever the interleaving of their copies and kernels."""
    import qst_b200
    g = torch.Generator().manual_seed(9)
    c = torch.randn(20_000, 128, generator=g)
    index = qst_b200.CorpusIndex(c.to(_dev()))
    batches = [torch.randn(n, 128, generator=g).pin_memory() for n in (300, 300, 1, 700, 300, 300, 129)]
    want = []
    for b in batches:
        v, i = qst_b200.topk_host(b, index, 10)
        want.append((v.clone(), i.clone()))
    pipe = qst_b200.HostTopkPipeline(index, 10)
    tickets = []
    for n, b in enumerate(batches):
        tickets.append(pipe.submit(b))
        if n >= 1:                                   # collect one behind: two batches in flight
            v, i = pipe.result(tickets[n - 1])
            assert torch.equal(i, want[n - 1][1]) and torch.equal(v, want[n - 1][0]), n - 1
    v, i = pipe.result(tickets[-1])
    assert torch.equal(i, want[-1][1]) and torch.equal(v, want[-1][0])
    with pytest.raises(ValueError):
        pipe.result(tickets[0])                      # long overwritten
    pipe.drain()


def test_large_query_batches_are_tiled(monkeypatch):
    """Batches above scoring.QUERY_TILE rows are walked in tiles (bounded workspace): same result."""
    import qst_b200
    from qst_b200 import scoring
    g = torch.Generator().manual_seed(31)
    q = torch.randn(350, 64, generator=g).to(_dev())
    index = qst_b200.CorpusIndex(torch.randn(5000, 64, generator=g).to(_dev()), "euclid_score")
    whole = qst_b200.topk(q, index, 10)
    monkeypatch.setattr(scoring, "QUERY_TILE", 100)
    tiled = qst_b200.topk(q, index, 10)
    assert tiled.plan.Q == 50                         # the last tile's plan: 350 = 3 * 100 + 50
    assert torch.equal(tiled.indices, whole.indices) and torch.equal(tiled.values, whole.values)
    assert bool((tiled.margin > 0).all())


def test_evaluator_scores_device_resident_corpus_in_one_pass():
    """corpus_embeddings already on the device: one pass over the corpus instead of corpus_chunk_size
    chunks + merge -- identical metrics and rankings."""
    import qst_b200
    q, c, queries, corpus, relevant = qst_b200.synth.ir_eval_set(300, 5000, 96)
    table = torch.cat([q, c]).to(_dev())
    model = qst_b200.synth.TableModel(table)
    kw = dict(corpus_chunk_size=700, mrr_at_k=[10], ndcg_at_k=[10], accuracy_at_k=[1, 10],
              precision_recall_at_k=[1, 10], map_at_k=[100], write_csv=False)
    ev = qst_b200.InformationRetrievalEvaluator(queries, corpus, relevant, score_functions={
        "cos_sim": qst_b200.cos_sim, "euclid_score": qst_b200.euclidean_score}, **kw)
    chunked = ev.rank(model)                                             # encodes chunk by chunk, merges 8 lists
    one_pass = ev.rank(model, corpus_embeddings=table[300:])            # resident: a single K2/K3 pass
    for fn in chunked:
        assert torch.equal(chunked[fn].indices, one_pass[fn].indices), fn
        assert torch.equal(chunked[fn].values, one_pass[fn].values), fn
    assert ev.compute_metrices(model) == ev.compute_metrices(model, corpus_embeddings=table[300:])
    host = ev.rank(model, corpus_embeddings=table[300:].cpu())          # host embeddings: chunked copies, same result
    assert all(torch.equal(host[fn].indices, chunked[fn].indices) for fn in chunked)


def test_evaluator_csv_is_byte_identical_to_the_reference_writer(tmp_path):
    """The PRODUCT's CSV writer, with the script defaults of ir_evauation_script.py:163-177 (k-lists up to
    900, three score functions, write_csv=True): ``ev(model, output_path)`` called twice (header once, two
    rows) must produce the same bytes as the oracle evaluator's file whenever the rankings are
    identical, and the same header and row layout in any case."""
    import os
    import qst_b200
    from oracle import ir_oracle
    q, c, queries, corpus, relevant = qst_b200.synth.ir_eval_set(300, 4000, 96)
    table = torch.cat([q, c])
    kw = dict(qst_b200.synth.SCRIPT_DEFAULT_K_LISTS, name="val")
    ev = qst_b200.InformationRetrievalEvaluator(queries, corpus, relevant, score_functions={
        "cos_sim": qst_b200.cos_sim, "dot_score": qst_b200.dot_score, "euclid_score": qst_b200.euclidean_score}, **kw)
    ref = ir_oracle.InformationRetrievalEvaluatorOracle(queries, corpus, relevant, score_functions={
        "cos_sim": ir_oracle.cos_sim, "dot_score": ir_oracle.dot_score, "euclid_score": ir_oracle.euclidean_score}, **kw)
    assert ev.write_csv and ev.csv_file == ref.csv_file
    ours, theirs = tmp_path / "ours", tmp_path / "theirs"
    ours.mkdir(); theirs.mkdir()
    model, ref_model = qst_b200.synth.TableModel(table.to(_dev())), ir_oracle.PrecomputedEmbeddingModel(table)
    got = [ev(model, output_path=str(ours), epoch=e, steps=s_) for e, s_ in ((0, 100), (1, -1))]
    want = [ref(ref_model, output_path=str(theirs), epoch=e, steps=s_) for e, s_ in ((0, 100), (1, -1))]
    a = open(os.path.join(ours, ev.csv_file), "rb").read()
    b = open(os.path.join(theirs, ref.csv_file), "rb").read()
    la, lb = a.decode().splitlines(), b.decode().splitlines()
    assert len(la) == len(lb) == 3 and la[0] == lb[0], "one header, two rows, same columns in the same order"
    assert [len(r.split(",")) for r in la] == [len(r.split(",")) for r in lb]
    # identical rankings (the normal case on this data) -> identical bytes and identical return values
    ranked = ev.rank(model)
    hits = ref.collect_hits(ref_model)
    same = all([[f"d{j}" for j in row] for row in ranked[fn].indices.cpu().tolist()] == ir_oracle.ranked_ids(hits[fn], 900)
               for fn in ("cos_sim", "dot_score", "euclid_score"))
    if same:
        assert a == b
        assert got == want
    else:   # tie swaps: every cell still parses and agrees to 1e-12 (metrics move by a swap inside a tie only at k cut-offs)
        print("rankings differ by tie swaps; CSV compared cell by cell")
        for ra, rb in zip(la[1:], lb[1:]):
            for x, y in zip(ra.split(","), rb.split(",")):
                assert abs(float(x) - float(y)) <= 1e-3
    assert all(v == 0 for v in ev.last_uncertified.values())


def test_exact_rescan_serves_more_than_8192_flagged_queries():
    """one pass of the re-scan lists at most 8192 flagged queries; the call must keep going
    until every flagged query has been repaired (here: all 9000 of them, none certified beforehand)."""
    import ctypes as C
    from qst_b200 import _lib, scoring
    lib = _lib.load()
    dev = _dev()
    g = torch.Generator().manual_seed(5)
    Q, N, D, k = 9000, 1500, 32, 10
    q = torch.randn(Q, D, generator=g)
    c = torch.randn(N, D, generator=g)
    want_val, want_idx = _oracle_topk(q, c, k, "dot_score")
    qd, cd = q.to(dev).contiguous(), c.to(dev).contiguous()
    vals = torch.full((Q, k), float("-inf"), device=dev)      # k-th best so far = -inf: every row qualifies
    idx = torch.full((Q, k), -1, dtype=torch.int64, device=dev)
    margin = torch.full((Q,), -1.0, device=dev)
    scratch = torch.empty(lib.qst_exact_rescan_workspace_bytes(Q, k), dtype=torch.uint8, device=dev)
    _lib.check(lib.qst_exact_rescan(Q, N, D, k, _lib.QST_SCORE_DOT, qd.data_ptr(), None, cd.data_ptr(), None, 0,
                                    vals.data_ptr(), idx.data_ptr(), margin.data_ptr(), scratch.data_ptr(),
                                    _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    assert bool(torch.isinf(margin).all()) and bool((margin > 0).all()), int((~(margin > 0)).sum())
    scale = float(want_val.abs().max())
    assert_same_ranking(idx, vals, want_idx, want_val, "rescan of 9000 flagged queries", scale=scale)
    # and rows wider than the old 3200-column limit of the re-scan are accepted (query batch in smem shrinks)
    D2 = 4000
    q2, c2 = torch.randn(3, D2, generator=g), torch.randn(200, D2, generator=g)
    w2v, w2i = _oracle_topk(q2, c2, 5, "dot_score")
    v2 = torch.full((3, 5), float("-inf"), device=dev)
    i2 = torch.full((3, 5), -1, dtype=torch.int64, device=dev)
    m2 = torch.full((3,), -1.0, device=dev)
    s2 = torch.empty(lib.qst_exact_rescan_workspace_bytes(3, 5), dtype=torch.uint8, device=dev)
    _lib.check(lib.qst_exact_rescan(3, 200, D2, 5, _lib.QST_SCORE_DOT, q2.to(dev).data_ptr(), None,
                                    c2.to(dev).contiguous().data_ptr(), None, 0, v2.data_ptr(), i2.data_ptr(),
                                    m2.data_ptr(), s2.data_ptr(), _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    assert_same_ranking(i2, v2, w2i, w2v, "rescan D=4000", scale=float(w2v.abs().max()))


@pytest.mark.parametrize("score", ["cos_sim", "dot_score", "euclid_score"])
def test_script_default_max_k_on_config1_shape(score):
    """max_k = 900 (ir_evauation_script.py:163-173) at BASELINE config 1's shape
    (1000 x 10 000 x 384), every score function of `--score_functions all`, against the oracle."""
    import qst_b200
    g = torch.Generator().manual_seed(41)
    q = torch.randn(1000, 384, generator=g)
    c = torch.randn(10_000, 384, generator=g)
    want_val, want_idx = _oracle_topk(q, c, 900, score)
    index = qst_b200.CorpusIndex(c.to(_dev()), score)
    res = qst_b200.topk(q.to(_dev()), index, 900)
    scale = float(want_val.abs().max()) if score == "dot_score" else 1.0
    assert_same_ranking(res.indices, res.values, want_idx, want_val, f"k=900 {score}", truth=(q, c, score), scale=scale)
    assert bool((res.margin > 0).all())


def test_max_k_900_on_a_long_corpus_sampled_against_brute_force():
    """k = 900 where it costs something: 4000 queries x 300 000 corpus rows x 128.  Size-independent
    properties for every query (certified, sorted, ids distinct and in range) and a brute-force fp32
    comparison for a sample."""
    import qst_b200
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(43)
    Q, N, D, k = 4000, 300_000, 128, 900
    q = torch.randn(Q, D, generator=g, device=dev)
    c = torch.randn(N, D, generator=g, device=dev)
    res = qst_b200.topk(q, qst_b200.CorpusIndex(c, "cos_sim"), k)
    assert bool((res.margin > 0).all())
    assert bool((res.values[:, 1:] <= res.values[:, :-1]).all())
    assert int(res.indices.min()) >= 0 and int(res.indices.max()) < N
    srt = torch.sort(res.indices, dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())
    sel = torch.linspace(0, Q - 1, 64, device=dev).long()
    sc = torch.nn.functional.normalize(q[sel], dim=1) @ torch.nn.functional.normalize(c, dim=1).T
    bv, bi = torch.topk(sc, k, dim=1)
    differ = res.indices[sel] != bi
    assert bool((((res.values[sel] - bv).abs() <= 2e-6) | ~differ).all())
    assert float((res.values[sel] - bv).abs().max()) <= 2e-6


@pytest.mark.parametrize("score", ["cos_sim", "dot_score", "euclid_score"])
def test_score_functions_called_directly_return_the_dense_matrix(score):
    """cos_sim / dot_score / euclidean_score called directly (the reference does at
    dataset/positive_examples_selection.py:55 and dataset/quadruplet_dataset.py:229-234) return the fp32
    [Q, N] matrix -- any N (round 1 stopped at 1024 columns) -- with the values top-k reports."""
    import qst_b200
    from oracle import ir_oracle
    fn = {"cos_sim": qst_b200.cos_sim, "dot_score": qst_b200.dot_score, "euclid_score": qst_b200.euclidean_score}[score]
    ref = {"cos_sim": ir_oracle.cos_sim, "dot_score": ir_oracle.dot_score, "euclid_score": ir_oracle.euclidean_score}[score]
    g = torch.Generator().manual_seed(12)
    q = torch.randn(37, 96, generator=g)
    c = torch.randn(3001, 96, generator=g) * (1 + torch.rand(3001, 1, generator=g))
    got = fn(q.to(_dev()), c.to(_dev()))
    want = ref(q, c)
    scale = float(want.abs().max()) if score == "dot_score" else 1.0
    assert got.shape == (37, 3001)
    if score == "euclid_score":
        assert_euclid_dense_close(got.cpu(), q, c, want)
    else:
        torch.testing.assert_close(got.cpu() / scale, want / scale, rtol=0, atol=2e-6)
    # 1-D inputs are promoted like the reference does, and the values are the ones topk() emits
    one = fn(q[0].to(_dev()), c.to(_dev()))
    assert one.shape == (1, 3001) and torch.equal(one[0], got[0])
    res = qst_b200.topk(q.to(_dev()), qst_b200.CorpusIndex(c.to(_dev()), score), 10)
    assert torch.equal(res.values, torch.gather(got, 1, res.indices))
    # the negative-mining filter built on it (dataset/quadruplet_dataset.py:229-234)
    if score == "cos_sim":
        mask, scores = qst_b200.dissimilar_mask(q[0].to(_dev()), c.to(_dev()), 0.2)
        assert mask.shape == (3001,) and torch.equal(mask, scores <= 0.2)


def test_bench_parity_sample_counts_ties_and_mismatches():
    """bench.py's in-run check: identical rankings pass, a swap inside a true tie is a tie (settled in
    float64 when the two fp32 summation orders disagree by more than the tolerance), a wrong document is
    a mismatch (the bench then exits non-zero)."""
    import importlib.util, os
    import qst_b200
    from qst_b200 import comm
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    dev = _dev()
    g = torch.Generator(device=dev).manual_seed(2)
    q = torch.randn(64, 96, generator=g, device=dev)
    c = torch.randn(5000, 96, generator=g, device=dev)
    c[4001] = c[17] * 3.0                    # same direction, different norm: an exact cos_sim tie with row 17
    res = qst_b200.topk(q, qst_b200.CorpusIndex(c, "cos_sim"), 20)
    cm = comm.SingleComm()
    ok = bench.parity_sample(cm, q, res.values, res.indices, c, 0, 20, n_sample=64)
    assert ok["queries"] == 64 and ok["mismatch"] == 0 and ok["identical"] + ok["ties_within_1e-6"] == 64
    wrong = res.indices.clone()
    other = int(res.indices[5, -1]) + 1 if int(res.indices[5, -1]) + 1 < 5000 else 0
    if other not in res.indices[5].tolist():
        wrong[5, 3] = other                  # a document that is not in the top 20, reported with a top-20 score
        bad = bench.parity_sample(cm, q, res.values, wrong, c, 0, 20, n_sample=64)
        assert bad["mismatch"] == 1
    vals_off = res.values.clone()
    vals_off[7, 0] += 1e-3                   # a wrong VALUE with the right document is a mismatch too
    off = bench.parity_sample(cm, q, vals_off, res.indices, c, 0, 20, n_sample=64)
    assert off["mismatch"] == 1 and off["max_abs_score_diff"] > 5e-4


@pytest.mark.parametrize("score", ["cos_sim", "dot_score", "euclid_score"])
@pytest.mark.parametrize("D", [1, 3, 100])
def test_degenerate_rows_zero_vectors_collisions_odd_widths(score, D):
    """What the reference's arithmetic does with rows it was not written for: all-zero query and document
    rows (``F.normalize`` divides by max(norm, 1e-12) -> a zero row scores 0 against everything),
    duplicated documents (exact score collisions: the lower corpus position ranks first), rows of 1, 3
    and 100 elements (padding of the bf16 operands), a single query, a single document."""
    import qst_b200
    g = torch.Generator().manual_seed(1000 + D)
    q = torch.randn(19, D, generator=g)
    c = torch.randn(301, D, generator=g)
    q[3] = 0
    c[5] = 0
    c[7] = c[9] = c[200]
    k = 10
    for qq, cc in ((q, c), (q[:1], c), (q, c[:1]), (q[3:4], c[5:6])):
        n = min(k, cc.shape[0])
        want_val, want_idx = _oracle_topk(qq, cc, n, score)
        if score == "euclid_score":
            # distances near 0 (a query next to a document in 1-3 dimensions): torch.cdist's matmul
            # formulation cancels catastrophically on the host, so the yardstick is the definition in
            # float64 (the same arbitration assert_same_ranking applies to euclid_score)
            want_val = (1 / (1 + torch.cdist(qq.double(), cc.double(), compute_mode="donot_use_mm_for_euclid_dist")
                             )).topk(n, dim=1).values.float()
        res = qst_b200.topk(qq.to(_dev()), qst_b200.CorpusIndex(cc.to(_dev()), score), k)
        got_val, got_idx = res.values.cpu(), res.indices.cpu()
        assert torch.isfinite(got_val[:, :n]).all()
        scale = max(float(want_val.abs().max()), 1.0) if score == "dot_score" else 1.0
        torch.testing.assert_close(got_val[:, :n] / scale, want_val / scale, rtol=0, atol=2e-6)
        # ids: the documents returned must carry the returned scores (ties make the id set ambiguous)
        dense = {"cos_sim": qst_b200.cos_sim, "dot_score": qst_b200.dot_score,
                 "euclid_score": qst_b200.euclidean_score}[score](qq.to(_dev()), cc.to(_dev())).cpu()
        assert torch.equal(torch.gather(dense, 1, got_idx[:, :n]), got_val[:, :n])
        assert bool((got_idx[:, n:] == -1).all())
        # collisions and all-equal rows: among equal scores the lower position comes first
        same = got_val[:, 1:n] == got_val[:, :n - 1]
        assert bool((got_idx[:, 1:n][same] > got_idx[:, :n - 1][same]).all())
        assert bool((res.margin > 0).all())


def test_no_queries_and_non_contiguous_inputs():
    """Zero queries give empty results (the reference's loops simply do not run); strided views and
    half-precision embeddings are accepted like torch accepts them."""
    import qst_b200
    g = torch.Generator().manual_seed(5)
    c = torch.randn(500, 64, generator=g)
    index = qst_b200.CorpusIndex(c.to(_dev()))
    res = qst_b200.topk(torch.empty(0, 64, device=_dev()), index, 10)
    assert res.values.shape == (0, 10) and res.indices.shape == (0, 10)
    wide = torch.randn(40, 128, generator=g)
    q = wide[:, ::2]                                  # stride 2
    want_val, want_idx = _oracle_topk(q.contiguous(), c, 10)
    res = qst_b200.topk(q.to(_dev()), index, 10)
    assert_same_ranking(res.indices, res.values, want_idx, want_val, "strided queries")
    res = qst_b200.topk(wide.to(_dev())[:, ::2], index, 10)
    assert_same_ranking(res.indices, res.values, want_idx, want_val, "strided device view")
    # fp16 embeddings (model.half()): scored from their fp32 up-cast, like torch.mm on .float()
    qh = q.half()
    want_val, want_idx = _oracle_topk(qh.float(), c, 10)
    res = qst_b200.topk(qh.to(_dev()), index, 10)
    assert_same_ranking(res.indices, res.values, want_idx, want_val, "fp16 queries")


def test_euclid_arbitration_survives_an_inaccurate_host_oracle():
    """The yardstick logic itself: a CPU oracle whose euclid scores are off by ~2e-5 (what a few hosts'
    cdist does) reorders neighbours; the float64 arbitration must accept OUR ranking then -- and must
    still reject a result of ours that is wrong."""
    import qst_b200
    g = torch.Generator().manual_seed(8)
    q = torch.randn(50, 96, generator=g)
    c = torch.randn(20000, 96, generator=g) * (1 + torch.rand(20000, 1, generator=g))
    k = 50
    res = qst_b200.topk(q.to(_dev()), qst_b200.CorpusIndex(c.to(_dev()), "euclid_score"), k)
    dense = 1 / (1 + torch.cdist(q, c, compute_mode="donot_use_mm_for_euclid_dist"))
    noisy = dense + 2e-5 * torch.randn(dense.shape, generator=g)
    bad_val, bad_idx = noisy.topk(k, dim=1)
    ARBITRATIONS.clear()
    assert_same_ranking(res.indices, res.values, bad_idx, bad_val, "noisy oracle", truth=(q, c, "euclid_score"))
    assert len(ARBITRATIONS) == 1
    wrong = res.values.clone()
    wrong[3, 7] += 1e-4
    with pytest.raises(AssertionError):
        assert_same_ranking(res.indices, wrong, bad_idx, bad_val, "noisy oracle, wrong result", truth=(q, c, "euclid_score"))
    ARBITRATIONS.clear()
