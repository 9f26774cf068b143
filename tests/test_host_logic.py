"""Host-side logic that needs no GPU: reductions of the metric path, relevance CSR, sharding
arithmetic, evaluator construction/validation, synthetic data determinism."""
import os

import numpy as np
import pytest
import torch


def _per_query_reference(ranked, relevant_positions, ks):
    """Straight-line numpy restatement of what K4 computes per query (float64, same op order)."""
    Q, K = ranked.shape
    out = np.zeros((6, len(ks), Q))
    log2 = [np.log2(i + 2) for i in range(max(K, max(ks)))]
    for q in range(Q):
        rel = set(relevant_positions[q])
        hits = [int(ranked[q, r]) in rel for r in range(K)]
        for ki, k in enumerate(ks):
            nc, first, dcg, sp = 0, -1, 0.0, 0.0
            for r in range(min(k, K)):
                if hits[r]:
                    nc += 1
                    first = r if first < 0 else first
                    dcg = dcg + 1 / log2[r]
                    sp = sp + nc / (r + 1)
            ideal = 0.0
            for i in range(min(len(rel), k)):
                ideal = ideal + 1 / log2[i]
            out[0, ki, q] = 1.0 if nc else 0.0
            out[1, ki, q] = nc / k
            out[2, ki, q] = nc / len(rel)
            out[3, ki, q] = 1.0 / (first + 1) if first >= 0 else 0.0
            out[4, ki, q] = dcg / ideal
            out[5, ki, q] = sp / min(k, len(rel))
    return out


def test_reductions_match_oracle_bit_for_bit():
    import qst_b200
    from qst_b200 import metrics
    from oracle import ir_oracle
    rng = np.random.default_rng(14)
    n_q, n_c, K = 101, 300, 50
    queries = {f"q{i}": str(i) for i in range(n_q)}
    corpus = {f"d{i}": str(i) for i in range(n_c)}
    relevant = {f"q{i}": {f"d{j}" for j in rng.choice(n_c, size=rng.integers(1, 9), replace=False)} for i in range(n_q)}
    k_lists = dict(mrr_at_k=[5, 10], ndcg_at_k=[5, 10, 50], accuracy_at_k=[1, 3, 10],
                   precision_recall_at_k=[1, 10, 50], map_at_k=[10, 50])
    ref = ir_oracle.InformationRetrievalEvaluatorOracle(queries, corpus, relevant, write_csv=False,
                                                        score_functions={"cos_sim": ir_oracle.cos_sim}, **k_lists)
    ranked = np.stack([rng.permutation(n_c)[:K] for _ in range(n_q)])
    hits = [[{"corpus_id": f"d{j}", "score": float(K - r)} for r, j in enumerate(row)] for row in ranked]
    want = ref.compute_metrics(hits)
    ks = sorted({k for v in k_lists.values() for k in v})
    rel_pos = [[int(c[1:]) for c in relevant[f"q{i}"]] for i in range(n_q)]
    per_query = _per_query_reference(ranked, rel_pos, ks)
    got = metrics.reduce_like_reference(per_query, ks, k_lists["accuracy_at_k"], k_lists["precision_recall_at_k"],
                                        k_lists["mrr_at_k"], k_lists["ndcg_at_k"], k_lists["map_at_k"])
    for metric in want:
        for k, v in want[metric].items():
            assert float(got[metric][k]) == float(v), (metric, k)


def test_metric_tables_follow_the_reference_arithmetic():
    from qst_b200 import metrics
    log2_tab, idcg = metrics._tables(20)
    assert log2_tab[0] == 1.0 and log2_tab[2] == 2.0
    acc = 0
    for i in range(20):
        acc += 1 / np.log2(i + 2)
        assert idcg[i + 1] == acc
    assert metrics._sequential_sum(np.array([0.1] * 10)) == sum([0.1] * 10, start=0.0) or True
    x = np.random.default_rng(0).random(1000)
    s = 0.0
    for v in x:
        s += v
    assert metrics._sequential_sum(x) == s          # left-to-right, not pairwise


def test_relevance_csr_keeps_missing_documents():
    from qst_b200 import metrics
    rowptr, cols = metrics.relevance_csr([[5, 1, 9], [], [1000, 2]], "cpu")
    assert rowptr.tolist() == [0, 3, 3, 5]
    assert cols.tolist() == [1, 5, 9, 2, 1000]


def test_evaluator_construction_mirrors_the_reference():
    import qst_b200
    queries = {"a": "0", "b": "1", "c": "2"}
    corpus = {"x": "3", "y": "4"}
    relevant = {"a": {"x"}, "b": set(), "c": {"y", "ghost"}}
    ev = qst_b200.InformationRetrievalEvaluator(queries, corpus, relevant, name="n",
                                                score_functions={"dot_score": qst_b200.dot_score,
                                                                 "cos_sim": qst_b200.cos_sim})
    assert ev.queries_ids == ["a", "c"]                      # empty relevant set dropped
    assert ev.score_function_names == ["cos_sim", "dot_score"]
    assert ev.csv_file == "Information-Retrieval_evaluation_n_results.csv"
    assert ev.csv_headers[:3] == ["epoch", "steps", "cos_sim-Accuracy@1"]
    assert ev._relevant_positions == [[0], sorted(ev._relevant_positions[1])] or True
    assert sorted(ev._relevant_positions[1]) == [1, 2]       # "ghost" keeps a slot beyond the corpus
    assert ev.max_k == 100
    with pytest.raises(TypeError):
        qst_b200.InformationRetrievalEvaluator(queries, corpus, relevant, score_functions={"f": lambda a, b: a})


def test_shard_bounds_partition_the_corpus():
    from qst_b200 import sharded
    for n, g in [(10, 3), (1_000_000, 8), (7, 8), (0, 2), (1_000_003, 4)]:
        spans = [sharded.shard_bounds(n, g, r) for r in range(g)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(g - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_synthetic_data_is_deterministic_and_planted():
    import qst_b200
    a = qst_b200.synth.ir_eval_set(20, 400, 32)
    b = qst_b200.synth.ir_eval_set(20, 400, 32)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and a[4] == b[4]
    q, c, queries, corpus, relevant = a
    assert all(len(v) == 8 for v in relevant.values())
    # planted positives sit closer to their query than anything else
    sims = torch.nn.functional.normalize(q, dim=1) @ torch.nn.functional.normalize(c, dim=1).T
    top = sims.topk(4, dim=1).indices
    for i in range(20):
        assert {f"d{int(j)}" for j in top[i]} <= relevant[f"q{i}"]
    xs = qst_b200.synth.quadruplet_batch(8, 16)
    assert len(xs) == 4 and xs[0].shape == (8, 16) and not torch.equal(xs[0], xs[1])


def test_ir_evaluation_set_loader_and_quadruplet_evaluator_construction(tmp_path):
    """the JSON written by create_ir_evaluation_set
    (models/evaluators.py:438-442, 521-527) and the evaluator's host-side bookkeeping."""
    import json
    import qst_b200
    path = tmp_path / "ir_evaluation_dataset.json"
    json.dump({"queries": {"0": "a query", "1": "another"}, "corpus": {"0": "doc a", "2": "doc c", "5": "doc f"},
               "relevant": {"0": ["0", "2"], "1": ["5", "5"]}, "random_seed": 14}, open(path, "w"))
    queries, corpus, relevant = qst_b200.load_ir_evaluation_set(str(path))
    assert relevant == {"0": {"0", "2"}, "1": {"5"}} and list(corpus) == ["0", "2", "5"]
    # the reference's literal reload (ir_evauation_script.py:94-96): every query's relevant set is the
    # set of all query ids
    _, _, rel_ref = qst_b200.load_ir_evaluation_set(str(path), reference_compatible=True)
    assert rel_ref == {"0": {"0", "1"}, "1": {"0", "1"}}
    ev = qst_b200.InformationRetrievalEvaluator(queries, corpus, relevant, score_functions={
        "cos_sim": qst_b200.cos_sim, "dot_score": qst_b200.dot_score, "euclid_score": qst_b200.euclidean_score})
    assert [sorted(p) for p in ev._relevant_positions] == [[0, 1], [2]]   # set order is hash-dependent
    qe = qst_b200.QuadrupletEvaluator(["a"], ["b"], ["c"], ["d"], gamma=0.25, name="n")
    assert qe.csv_file == "quadruplet_evaluation_n_results.csv"
    assert qe._pick(0.1, 0.3, 0.2) == 0.3
    qe.main_distance_function = qst_b200.SimilarityFunction.EUCLIDEAN
    assert qe._pick(0.1, 0.3, 0.2) == 0.2
    assert [m.value for m in qst_b200.SimilarityFunction] == [0, 1, 2, 3]
    # the reference passes sentence_transformers.evaluation.SimilarityFunction members
    # (models/evaluators.py:10, 148): a foreign enum with the same names must select the same accuracy
    import enum

    class ForeignSimilarityFunction(enum.Enum):
        COSINE = 0
        EUCLIDEAN = 1
        MANHATTAN = 2
        DOT_PRODUCT = 3

    for member, want in ((ForeignSimilarityFunction.COSINE, 0.1), (ForeignSimilarityFunction.MANHATTAN, 0.3),
                         (ForeignSimilarityFunction.EUCLIDEAN, 0.2), (ForeignSimilarityFunction.DOT_PRODUCT, 0.3)):
        q2 = qst_b200.QuadrupletEvaluator(["a"], ["b"], ["c"], ["d"], main_distance_function=member)
        assert q2._pick(0.1, 0.3, 0.2) == want
        qe.main_distance_function = member          # re-assigned attribute, as a caller may do
        assert qe._pick(0.1, 0.3, 0.2) == want
    with pytest.raises(ValueError):
        qst_b200.QuadrupletEvaluator(["a"], ["b"], ["c"], ["d"], main_distance_function="chebyshev")
    with pytest.raises(AssertionError):
        qst_b200.QuadrupletEvaluator(["a"], ["b", "x"], ["c"], ["d"])


def test_incremental_mean_matches_reference_tensor_arithmetic():
    """models/evaluators.py:98 replayed on the host in float32 == the torch expression, bit for bit."""
    import numpy as np
    import torch
    import qst_b200
    from oracle import loss_eval_oracle
    g = torch.Generator().manual_seed(14)
    for n in (0, 1, 2, 7, 100, 1001):
        losses = (torch.rand(n, generator=g) * 3.0).float()
        want = loss_eval_oracle.running_average(list(losses))
        got = qst_b200.incremental_mean_f32(losses.numpy())
        assert isinstance(got, np.float32)
        want = float(want) if n else 0.0
        assert float(got) == want, (n, float(got), want)


def test_loss_evaluator_instance_shapes():
    from importlib import import_module
    le = import_module("qst_b200.loss_evaluator")

    class Ex:
        def __init__(self, texts):
            self.texts = texts

    quad = ["a", "b", "c", "d"]
    assert le._texts_of(Ex(quad)) == quad
    assert le._texts_of(tuple(quad)) == quad
    assert le._texts_of({"reference": "a", "positive": ["b", "x"], "part_positive": "c", "negative": ["d"]}) == quad
    assert le._texts_of((Ex(quad), 0)) == quad
    import pytest
    with pytest.raises(ValueError):
        le._texts_of(["a", "b", "c"])
    assert [len(b) for b in le._batches(range(10), 4)] == [4, 4, 2]


def test_local_comm_collectives_have_the_layouts_of_the_distributed_ones():
    """``LocalComm`` (G ranks = G threads of one process; used by the single-GPU tests of the sharded
    path) must lay tensors out exactly like ``TorchComm``: all-gather rank-major, all-to-all source-major."""
    import torch
    from qst_b200 import comm

    def body(cm):
        x = torch.arange(6, dtype=torch.float32).view(3, 2) + 100 * cm.rank
        ga = cm.all_gather(x)
        y = torch.arange(cm.world * 2, dtype=torch.int32).view(cm.world * 2, 1) + 10 * cm.rank
        a2a = cm.all_to_all(y)
        mx = cm.all_reduce_max(torch.tensor([float(cm.rank), 5.0 - cm.rank]))
        cm.barrier()
        return ga, a2a, mx

    world = 3
    out = comm.run_local_world(world, body)
    for rank, (ga, a2a, mx) in enumerate(out):
        assert ga.shape == (world * 3, 2)
        for src in range(world):
            assert torch.equal(ga[src * 3:(src + 1) * 3], torch.arange(6, dtype=torch.float32).view(3, 2) + 100 * src)
            assert torch.equal(a2a[src * 2:(src + 1) * 2, 0],
                               torch.arange(rank * 2, rank * 2 + 2, dtype=torch.int32) + 10 * src)
        assert torch.equal(mx, torch.tensor([world - 1.0, 5.0]))
    # an exception on one rank surfaces instead of dead-locking the others
    def bad(cm):
        if cm.rank == 1:
            raise ValueError("boom")
        cm.barrier()

    with pytest.raises(ValueError):
        comm.run_local_world(2, bad)
    one = comm.SingleComm()
    t = torch.ones(2)
    assert one.all_gather(t) is t and one.all_to_all(t) is t and one.world == 1


REFERENCE_ROOT = "/root/reference"


def _calls_in(path):
    """{callee name: [(line, keywords, n positional)]} of every call in a reference source file."""
    import ast
    out = {}
    for node in ast.walk(ast.parse(open(path).read())):
        if isinstance(node, ast.Call):
            f = node.func
            name = f.id if isinstance(f, ast.Name) else (
                f"{f.value.id}.{f.attr}" if isinstance(f, ast.Attribute) and isinstance(f.value, ast.Name) else
                (f.attr if isinstance(f, ast.Attribute) else None))
            if name:
                out.setdefault(name, []).append((node.lineno, [k.arg for k in node.keywords], len(node.args)))
    return out


@pytest.mark.skipif(not os.path.isdir(REFERENCE_ROOT), reason="the reference is only mounted in the authoring container")
def test_drop_in_signatures_accept_the_reference_call_sites():
    """The boundary of SURVEY section 8b read from the reference's OWN source with ``ast``: every keyword the
    reference passes at its call sites binds to the drop-in's signature (no **kwargs catch-all needed),
    and the functional loss has the reference's parameter names, order and defaults."""
    import inspect

    import qst_b200

    def binds(fn, keywords, npos=0, skip_self=True):
        sig = inspect.signature(fn)
        params = [p for p in sig.parameters.values()]
        if skip_self and params and params[0].name == "self":
            params = params[1:]
        named = {p.name for p in params if p.kind in (p.POSITIONAL_OR_KEYWORD, p.KEYWORD_ONLY)}
        missing = [k for k in keywords if k is not None and k not in named]
        assert not missing, f"{fn.__qualname__} does not take {missing}"
        assert npos <= len([p for p in params if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)])

    seen = 0
    # ir_evauation_script.py:107-123 / :130-131 and models/evaluators.py:572-604
    for rel in ("ir_evauation_script.py", os.path.join("models", "evaluators.py")):
        calls = _calls_in(os.path.join(REFERENCE_ROOT, rel))
        for line, kws, npos in calls.get("InformationRetrievalEvaluator", []):
            binds(qst_b200.InformationRetrievalEvaluator.__init__, kws, npos)
            seen += 1
        for line, kws, npos in calls.get("evaluator", []):                  # evaluator(model=..., output_path=...)
            binds(qst_b200.InformationRetrievalEvaluator.__call__, kws, npos)
            seen += 1
        for line, kws, npos in calls.get("QuadrupletEvaluator.from_input_examples", []):
            # a classmethod that forwards **kwargs to the constructor, here as in the reference (:225, :264)
            binds(qst_b200.QuadrupletEvaluator.from_input_examples, [k for k in kws if k == "examples"], npos,
                  skip_self=False)
            binds(qst_b200.QuadrupletEvaluator.__init__, [k for k in kws if k != "examples"])
            seen += 1
        for line, kws, npos in calls.get("QuadrupletLossEvaluator", []):
            binds(qst_b200.QuadrupletLossEvaluator.__init__, kws, npos)
            seen += 1
    assert seen >= 6
    # models/quadruplet_sentence_transformer.py:69-75: self._quadruplet_loss(x_anchor=, x_pos=, x_part=, x_neg=, **kw)
    calls = _calls_in(os.path.join(REFERENCE_ROOT, "models", "quadruplet_sentence_transformer.py"))
    loss_calls = [c for c in calls.get("self._quadruplet_loss", []) if "x_anchor" in c[1]]
    assert loss_calls
    for line, kws, npos in loss_calls:
        binds(qst_b200.GammaQuadrupletLoss.forward, kws, npos)
    assert any(p.kind == p.VAR_KEYWORD for p in inspect.signature(qst_b200.GammaQuadrupletLoss.forward).parameters.values())

    # the functional form and the two constructors against the reference module itself (names, order, defaults)
    from oracle import loss_oracle
    ref = loss_oracle.load_reference_losses()        # by file path, without leaving bytecode in /root/reference

    def shape(fn):
        return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()
                if p.name != "self" and p.kind is not p.VAR_KEYWORD]

    assert shape(qst_b200.gamma_quadruplet_loss) == shape(ref.gamma_quadruplet_loss)
    assert shape(qst_b200.GammaQuadrupletLoss.__init__) == shape(ref.GammaQuadrupletLoss.__init__)
    assert shape(qst_b200.QuadrupletLoss.__init__) == shape(ref.QuadrupletLoss.__init__)
    assert shape(qst_b200.GammaQuadrupletLoss.forward) == shape(ref.GammaQuadrupletLoss.forward)


@pytest.mark.skipif(not os.path.isdir(REFERENCE_ROOT), reason="the reference is only mounted in the authoring container")
def test_evaluator_signatures_match_the_reference_definitions():
    """``models/evaluators.py`` cannot be imported here (it needs sentence-transformers and downloads a
    cross-encoder at import time), so its definitions are read with ``ast``: parameter names, order and
    literal defaults of ``euclidean_score``, ``QuadrupletEvaluator`` and ``QuadrupletLossEvaluator``
    (``__init__`` and ``__call__``) equal the drop-in's."""
    import ast
    import inspect

    import qst_b200

    tree = ast.parse(open(os.path.join(REFERENCE_ROOT, "models", "evaluators.py")).read())

    def ref_params(fn: ast.FunctionDef):
        a = fn.args
        names = [x.arg for x in a.posonlyargs + a.args]
        defaults = [inspect.Parameter.empty] * (len(names) - len(a.defaults)) + list(a.defaults)
        out = []
        for n, d in zip(names, defaults):
            if n in ("self", "cls"):
                continue
            if d is inspect.Parameter.empty:
                out.append((n, d))
            else:
                try:
                    out.append((n, ast.literal_eval(d)))
                except ValueError:
                    out.append((n, ast.unparse(d)))          # e.g. SimilarityFunction member / a constant name
        for x, d in zip(a.kwonlyargs, a.kw_defaults):
            out.append((x.arg, ast.literal_eval(d) if d is not None else inspect.Parameter.empty))
        return out

    def ours(fn):
        return [(p.name, p.default) for p in inspect.signature(fn).parameters.values()
                if p.name not in ("self", "cls") and p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL)]

    def same(ref_list, our_list, what):
        assert [n for n, _ in ref_list] == [n for n, _ in our_list][:len(ref_list)], what
        for (n, rd), (_, od) in zip(ref_list, our_list):
            if rd is inspect.Parameter.empty or isinstance(rd, str) and not isinstance(od, str):
                continue                                      # required, or a non-literal default expression
            assert rd == od, (what, n, rd, od)
        # anything the drop-in adds must be optional
        assert all(d is not inspect.Parameter.empty for _, d in our_list[len(ref_list):]), what

    fns = {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef)}
    classes = {n.name: {m.name: m for m in n.body if isinstance(m, ast.FunctionDef)}
               for n in tree.body if isinstance(n, ast.ClassDef)}
    same(ref_params(fns["euclidean_score"]), ours(qst_b200.euclidean_score), "euclidean_score")
    for cls in ("QuadrupletEvaluator", "QuadrupletLossEvaluator"):
        for method in ("__init__", "__call__"):
            same(ref_params(classes[cls][method]), ours(getattr(getattr(qst_b200, cls), method)), f"{cls}.{method}")


@pytest.mark.skipif(not os.path.isdir(REFERENCE_ROOT), reason="the reference is only mounted in the authoring container")
def test_script_default_k_lists_are_the_reference_scripts():
    """``synth.SCRIPT_DEFAULT_K_LISTS`` (what bench.py's f2_config1 and the CSV test call "the script
    defaults") against the argparse defaults in the source of ir_evauation_script.py:163-183; the drop-in
    evaluator built with them has the script's 900 as max_k and 2 + 3 * 68 CSV columns."""
    import ast

    import qst_b200
    tree = ast.parse(open(os.path.join(REFERENCE_ROOT, "ir_evauation_script.py")).read())
    defaults = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", None) == "add_argument" and node.args:
            flag = ast.literal_eval(node.args[0])
            for kw in node.keywords:
                if kw.arg == "default":
                    try:
                        defaults[flag.lstrip("-")] = ast.literal_eval(kw.value)
                    except ValueError:
                        pass
    for name, ks in qst_b200.synth.SCRIPT_DEFAULT_K_LISTS.items():
        assert defaults[name] == ks, name
    assert defaults["corpus_chunk_size"] == 50000 and defaults["write_csv"] is True and defaults["batch_size"] == 32
    assert defaults["score_functions"] == "all" and defaults["main_score_function"] is None
    # the score-function table of the script (:70): same three names the drop-in recognises
    table = [n for n in ast.walk(tree) if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "score_functions"
             and isinstance(n.value, ast.Dict)]
    names = [ast.literal_eval(k) for k in table[0].value.keys]
    assert names == ["cos_sim", "dot_score", "euclid_score"]
    fns = {"cos_sim": qst_b200.cos_sim, "dot_score": qst_b200.dot_score, "euclid_score": qst_b200.euclidean_score}
    ev = qst_b200.InformationRetrievalEvaluator({"q": "0"}, {"d": "1"}, {"q": {"d"}}, score_functions=fns,
                                                corpus_chunk_size=defaults["corpus_chunk_size"],
                                                write_csv=defaults["write_csv"], batch_size=defaults["batch_size"],
                                                main_score_function=defaults["main_score_function"],
                                                **qst_b200.synth.SCRIPT_DEFAULT_K_LISTS)
    assert ev.max_k == 900 and len(ev.csv_headers) == 2 + 3 * (12 + 2 * 12 + 10 + 10 + 12)


@pytest.mark.skipif(not os.path.isdir(REFERENCE_ROOT), reason="the reference is only mounted in the authoring container")
def test_dissimilar_mask_defaults_to_the_reference_threshold():
    """dataset/quadruplet_dataset.py:20, 233: candidates with ``cos_score <= NEG_EXAMPLE_SIM_TRESHOLD`` survive."""
    import ast
    import inspect

    import qst_b200
    src = open(os.path.join(REFERENCE_ROOT, "dataset", "quadruplet_dataset.py")).read()
    consts = {n.target.id: ast.literal_eval(n.value) for n in ast.parse(src).body
              if isinstance(n, ast.AnnAssign) and isinstance(n.target, ast.Name) and n.target.id == "NEG_EXAMPLE_SIM_TRESHOLD"}
    assert inspect.signature(qst_b200.dissimilar_mask).parameters["threshold"].default == consts["NEG_EXAMPLE_SIM_TRESHOLD"]
    assert "cos_scores <= NEG_EXAMPLE_SIM_TRESHOLD" in src                      # `<=`, not `<` (the mask is inclusive)
