"""The host side of SURVEY 8f rows 3-4 (and a8) pinned to the reference's OWN code.

``tests/golden/evaluators_golden.json`` holds what ``/root/reference/models/evaluators.py`` itself produces
(``QuadrupletEvaluator`` sampling / global accuracy / CSV, ``QuadrupletLossEvaluator`` incremental mean / JSON
log, ``euclidean_score``, the evaluation-set file of ``create_ir_evaluation_set`` and its reloads) when only
its third-party collaborators are scripted -- see ``tests/reference_sandbox.py`` and
``tests/golden/make_evaluators_golden.py``.  Here the drop-in classes and the oracle restatements are held against those vectors on the CPU (the device work underneath -- embeddings,
the nine comparison counts, the per-batch losses -- is scripted the same way and is covered by the ``-m gpu``
tests), and, when the reference is mounted, the fixture is re-derived from the reference and must not have
drifted."""
import json
import os
import random
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "golden"))

import make_evaluators_golden as gen  # noqa: E402
import reference_sandbox as rs  # noqa: E402


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(HERE, "golden", "evaluators_golden.json")) as fp:
        return json.load(fp)


def test_sampling_of_quadruplets_follows_the_reference_draw_for_draw(golden):
    """models/evaluators.py:224-262 and :264-343: same ``random`` stream -> same sentences, at construction
    and at the re-sampling after 5 epochs (and NOT before)."""
    import qst_b200
    examples = gen.dict_examples(golden["sampling"]["n"])
    random.seed(golden["sampling"]["seed"])
    ev = qst_b200.QuadrupletEvaluator.from_input_examples(examples, gamma=0.6, name="s")
    assert [ev.anchors, ev.positives, ev.partially_positives, ev.negatives] == golden["sampling"]["first"]
    for epoch in range(5):
        if epoch == 4:
            assert [ev.anchors, ev.positives, ev.partially_positives, ev.negatives] == golden["sampling"]["first"]
        ev._reset_examples()
    assert [ev.anchors, ev.positives, ev.partially_positives, ev.negatives] == golden["sampling"]["after_5_epochs"]
    # InputExample-shaped items and (item, label) pairs take the texts as they are
    items = [(rs.InputExample(texts=[f"a{i}", f"p{i}", f"q{i}", f"n{i}"]), 0) for i in range(3)]
    ev = qst_b200.QuadrupletEvaluator.from_input_examples(items)
    assert ev.positives == ["p0", "p1", "p2"] and ev.negatives == ["n0", "n1", "n2"]


# examples per call such that count / n reproduces every scripted accuracy of that call exactly
_N_PER_CALL = [100, 81, 7, 7, 30]


def _count(acc, n):
    c = round(acc * n)
    assert c / n == acc, (acc, n)
    return c


@pytest.mark.parametrize("main", ["COSINE", "MANHATTAN", "EUCLIDEAN", None])
def test_quadruplet_evaluator_host_path_equals_the_reference(golden, tmp_path, monkeypatch, main):
    """Global accuracy (:367), return value and CSV (:374-387) for scripted comparison counts: the drop-in's
    ``__call__`` with the device work replaced by the counts that give the reference's scripted triplet
    accuracies -- returned floats ``==``, CSV text byte for byte."""
    import qst_b200
    from qst_b200 import quad_evaluator as qe
    column = {"COSINE": 0, "MANHATTAN": 1, "EUCLIDEAN": 2, None: 1}[main]
    for case in golden["quadruplet_evaluator"]:
        out = tmp_path / f"{main}_{case['gamma']}"
        out.mkdir()
        ev = qst_b200.QuadrupletEvaluator(["a"], ["p"], ["pp"], ["n"], gamma=case["gamma"], name="val",
                                          main_distance_function=None if main is None else rs.SimilarityFunction[main])
        monkeypatch.setattr(ev, "_encode", lambda model, sentences: None)
        returned = []
        for i, (epoch, steps) in enumerate(golden["calls"]):
            n = _N_PER_CALL[i]
            counts = torch.zeros(3, 3, dtype=torch.int64)                  # [metric, pair]; others lose (None: max)
            for j, pair in enumerate(("pos_part", "pos_neg", "part_neg")):
                counts[column, j] = _count(golden["triplet_script"][pair][i], n)
            monkeypatch.setattr(qe, "paired_distance_counts", lambda *embs, _c=counts: _c)
            ev.anchors = ev.positives = ev.partially_positives = ev.negatives = ["x"] * n
            returned.append(ev(None, output_path=str(out), epoch=epoch, steps=steps))
        assert returned == case["returned"]
        assert ev.csv_file == case["csv_file"]
        assert open(out / ev.csv_file, newline="", encoding="utf-8").read() == case["csv_text"]


def test_loss_evaluator_running_mean_and_log_equal_the_reference(golden, tmp_path, monkeypatch):
    """models/evaluators.py:84-128 for scripted batch losses: returned value (the reference returns the 0-dim
    float32 tensor, the drop-in its float) and the JSON log text, two calls appending to one file."""
    import qst_b200
    from oracle import loss_eval_oracle
    for case in golden["loss_evaluator"]:
        n_batches = len(case["batch_sizes_seen"])
        losses = golden["batch_losses"][:n_batches]
        out = tmp_path / f"{case['n_items']}_{case['batch_size']}"
        out.mkdir()
        ev = qst_b200.QuadrupletLossEvaluator(list(range(case["n_items"])), None, batch_size=case["batch_size"])
        # the drop-in batches the dataset like the reference's DataLoader does
        from qst_b200.loss_evaluator import _batches
        assert [len(b) for b in _batches(list(range(case["n_items"])), case["batch_size"])] == case["batch_sizes_seen"]
        monkeypatch.setattr(ev, "batch_losses", lambda model, _l=losses: torch.tensor(_l, dtype=torch.float32))
        first = ev(None, output_path=str(out), epoch=0, steps=-1)
        monkeypatch.setattr(ev, "batch_losses", lambda model, _l=losses: torch.tensor(_l[::-1], dtype=torch.float32))
        second = ev(None, output_path=str(out), epoch=1, steps=40)
        assert [first, second] == case["returned"] and isinstance(first, float)
        assert open(out / "_quadruplet_loss_eval.json").read() == case["json_text"]
        # the oracle restatement of the same expression
        for seq, want in ((losses, case["returned"][0]), (losses[::-1], case["returned"][1])):
            got = loss_eval_oracle.running_average([torch.tensor(v, dtype=torch.float32) for v in seq])
            assert got.dtype == torch.float32 and float(got) == want
            assert float(qst_b200.loss_evaluator.incremental_mean_f32(np.asarray(seq, dtype=np.float32))) == want


def test_euclidean_score_oracle_equals_the_reference_function(golden):
    """models/evaluators.py:392-405 (tensor / list / 1-D inputs).  1e-6: ``torch.cdist`` is not bit-stable
    across host CPUs (matmul formulation above 25 rows); on the generating host it is exact, see the live test."""
    from oracle import ir_oracle
    g = golden["euclidean_score"]
    a, b = torch.tensor(g["a"]), torch.tensor(g["b"])
    torch.testing.assert_close(ir_oracle.euclidean_score(a, b), torch.tensor(g["scores"]), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(ir_oracle.euclidean_score(a[0], b[1]), torch.tensor(g["one_d"]), rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(ir_oracle.euclidean_score(a[:2].tolist(), b[:3].tolist()),
                               torch.tensor(g["from_lists"]), rtol=1e-6, atol=1e-7)
    assert torch.tensor(g["one_d"]).shape == (1, 1)


def test_evaluation_set_written_by_the_reference_loads_as_the_reference_reloads_it(golden, tmp_path):
    """SURVEY 8f row 4.  The file is the one ``create_ir_evaluation_set`` (models/evaluators.py:406-530) wrote;
    ``load_ir_evaluation_set`` gives the per-query sets the function itself returns and reloads (:414-431),
    ``reference_compatible=True`` the sets ``get_sequential_evaluator`` (:556-561) and the script
    (ir_evauation_script.py:94-96) hand to the evaluator -- every query's set = all query ids.  The keyword
    set the reference then passes (:572-588) builds the drop-in evaluator."""
    import qst_b200
    fns = {"cos_sim": qst_b200.cos_sim, "dot_score": qst_b200.dot_score}
    for i, case in enumerate(golden["ir_evaluation_set"]):
        path = tmp_path / f"created_eval_queries_{i}.json"
        path.write_text(case["file_text"])
        queries, corpus, relevant = qst_b200.load_ir_evaluation_set(str(path))
        assert queries == case["queries"] and corpus == case["corpus"]
        assert list(queries) == list(case["queries"]) and list(corpus) == list(case["corpus"])     # order kept
        assert relevant == {q: set(v) for q, v in case["relevant"].items()}
        q2, c2, literal = qst_b200.load_ir_evaluation_set(str(path), reference_compatible=True)
        assert (q2, c2) == (queries, corpus)
        assert literal == {q: set(v) for q, v in case["relevant_after_sequential_evaluator_reload"].items()}
        assert all(v == set(queries) for v in literal.values())
        ev = qst_b200.InformationRetrievalEvaluator(
            queries=queries, corpus=corpus, relevant_docs=relevant,
            score_functions={n: fns[n] for n in case["ire_score_function_names"]}, **case["ire_kwargs"])
        assert ev.csv_file == "Information-Retrieval_evaluation_val_results.csv"
        assert ev.queries_ids == [q for q in queries if relevant[q]]
        assert case["sequential_order"][1:] == ["QuadrupletEvaluator", "QuadrupletLossEvaluator"]


def test_loss_evaluator_oracle_equals_the_reference_full_stack(golden):
    """``loss_eval_oracle.evaluate`` -- what the GPU test holds the drop-in ``QuadrupletLossEvaluator`` against --
    compared with the recorded result of the reference's complete, un-scripted stack (its evaluator, its
    loss-model wrapper, its loss module) on the same table-lookup model.  1e-6 relative across host CPUs; bit
    for bit on the host that runs the reference (live test below)."""
    from oracle import loss_eval_oracle
    n, table = gen.full_stack_table()
    a, p, pp, neg = table[:n], table[n:2 * n], table[2 * n:3 * n], table[3 * n:]
    for case in golden["full_loss_stack"]:
        got, _ = loss_eval_oracle.evaluate(a, p, pp, neg, case["batch_size"], **case["loss_kwargs"])
        assert float(got) == pytest.approx(case["average_loss"], rel=1e-6)


@pytest.mark.skipif(not rs.available(), reason="the reference is only mounted in the authoring container")
def test_fixture_is_what_the_reference_produces_now():
    """Re-derives every vector from /root/reference and compares with the committed file: the fixture cannot
    drift from the code it claims to come from."""
    from oracle import ir_oracle
    with open(os.path.join(HERE, "golden", "evaluators_golden.json")) as fp:
        want = json.load(fp)
    ns = rs.load(*gen.LIFTED)
    assert [gen.run_quadruplet_evaluator(ns, g) for g in gen.GAMMAS] == want["quadruplet_evaluator"]
    got = [gen.run_ir_evaluation_set(ns, *(c["flags"][k] for k in ("use_pos", "use_part_pos", "add_part_pos_corpus")))
           for c in want["ir_evaluation_set"]]
    assert got == want["ir_evaluation_set"]
    got = [gen.run_loss_evaluator(ns, c["n_items"], c["batch_size"]) for c in want["loss_evaluator"]]
    assert got == want["loss_evaluator"]
    # the un-scripted loss stack: the reference's evaluator + loss-model wrapper + loss module == the oracle
    from oracle import loss_eval_oracle
    n, table = gen.full_stack_table()
    for (kw, bs), rec in zip(gen.FULL_STACK_CASES, want["full_loss_stack"]):
        theirs = gen.run_full_loss_stack(kw, bs)
        ours, _ = loss_eval_oracle.evaluate(table[:n], table[n:2 * n], table[2 * n:3 * n], table[3 * n:], bs, **kw)
        assert theirs.dtype == torch.float32 and torch.equal(theirs, ours)
        assert float(theirs) == rec["average_loss"] and rec["loss_kwargs"] == kw and rec["batch_size"] == bs
    random.seed(14)
    ev = ns["QuadrupletEvaluator"].from_input_examples(gen.dict_examples(), gamma=0.6, name="s")
    assert [ev.anchors, ev.positives, ev.partially_positives, ev.negatives] == want["sampling"]["first"]
    # the three inner evaluators are built on the sampled lists with the reference's pairing (:188-220)
    built = {t.name: t for t in rs.ScriptedTriplet.built[-3:]}
    assert built["pos_part"].negatives is ev.partially_positives and built["part_neg"].positives is ev.partially_positives
    assert built["pos_neg"].positives is ev.positives and built["pos_neg"].negatives is ev.negatives
    # euclidean_score: reference function vs oracle restatement, same host -> bit for bit
    g = torch.Generator().manual_seed(3)
    for shape_a, shape_b in (((5, 12), (9, 12)), ((40, 64), (70, 64)), ((1, 3), (2, 3))):
        a, b = torch.randn(*shape_a, generator=g), torch.randn(*shape_b, generator=g)
        assert torch.equal(ns["euclidean_score"](a, b), ir_oracle.euclidean_score(a, b))
    assert torch.equal(ns["euclidean_score"](a[0], b), ir_oracle.euclidean_score(a[0], b))
    assert torch.equal(ns["euclidean_score"](a.tolist(), b.tolist()), ir_oracle.euclidean_score(a.tolist(), b.tolist()))


@pytest.mark.skipif(not rs.available(), reason="the reference is only mounted in the authoring container")
def test_reference_loss_model_wrapper_drives_the_drop_in_loss_up_to_the_device_boundary():
    """The loss protocol of SURVEY 8b from the caller's side: the reference's own
    ``QuadrupletSentenceTransformerLossModel`` (models/quadruplet_sentence_transformer.py:9-78) takes the drop-in
    ``GammaQuadrupletLoss`` as its ``quadruplet_loss`` (an ``nn.Module``, registered as a sub-module), runs the
    sentence model on the four texts and calls it with the reference's keywords -- and on a box without a GPU
    the drop-in answers with its no-CPU-fallback error instead of computing on the host.  A per-batch
    ``additional_loss_kwargs`` entry (``reduction``) reaches ``forward`` like it reaches the reference's."""
    import qst_b200
    wrapper_cls = rs.load_loss_model()
    n, table = gen.full_stack_table()
    model = rs.TableSentenceModel(table)
    loss = qst_b200.GammaQuadrupletLoss(gamma=0.6, margin_pos_neg=1.0, margin_pos_part=0.5, margin_part_neg=0.5)
    wrapper = wrapper_cls(st_model=model, quadruplet_loss=loss)
    assert dict(wrapper.named_modules())["_quadruplet_loss"] is loss and loss.gamma == 0.6
    features = [torch.arange(0, 8), torch.arange(n, n + 8), torch.arange(2 * n, 2 * n + 8), torch.arange(3 * n, 3 * n + 8)]
    seen = {}
    real_forward = loss.forward

    def spy(*args, **kwargs):
        seen["args"], seen["kwargs"] = args, dict(kwargs)
        return real_forward(*args, **kwargs)

    loss.forward = spy
    with pytest.raises(qst_b200.QstError):
        wrapper(features, None)
    assert seen["args"] == () and sorted(seen["kwargs"]) == ["x_anchor", "x_neg", "x_part", "x_pos"]
    assert torch.equal(seen["kwargs"]["x_pos"], table[n:n + 8])
    # dict-shaped features + an additional loss kwarg taken from the batch (:60-66)
    wrapper = wrapper_cls(st_model=model, quadruplet_loss=loss, additional_loss_kwargs=["reduction"])
    batch = {"reference": features[0], "positive": features[1], "part_positive": features[2], "negative": features[3],
             "reduction": "sum"}
    with pytest.raises(qst_b200.QstError):
        wrapper(batch, None)
    assert seen["kwargs"]["reduction"] == "sum" and torch.equal(seen["kwargs"]["x_part"], table[2 * n:2 * n + 8])
    # the same wrapper around the reference's loss computes; the drop-in must be interchangeable in that slot
    from oracle import loss_oracle
    ref_loss = loss_oracle.load_reference_losses().GammaQuadrupletLoss(gamma=0.6, margin_pos_neg=1.0,
                                                                       margin_pos_part=0.5, margin_part_neg=0.5)
    value = wrapper_cls(st_model=model, quadruplet_loss=ref_loss, additional_loss_kwargs=["reduction"])(batch, None)
    want = loss_oracle.gamma_quadruplet_loss(table[:8], table[n:n + 8], table[2 * n:2 * n + 8], table[3 * n:3 * n + 8],
                                             gamma=0.6, margin_pos_neg=1.0, margin_pos_part=0.5, margin_part_neg=0.5,
                                             reduction="sum")
    assert torch.equal(value, want)


@pytest.mark.skipif(not rs.available(), reason="the reference is only mounted in the authoring container")
def test_text_order_of_a_dataset_instance_is_the_reference_to_input_example_order():
    """models/quadruplet_sentence_transformer.py:83-97 lays a dict instance out as [reference, positive,
    part_positive, negative]; the drop-in QuadrupletLossEvaluator reads dict instances, InputExamples and
    (instance, label) pairs in that order."""
    import ast

    from qst_b200.loss_evaluator import QUADRUPLET_KEYS, _texts_of
    tree = ast.parse(open(rs.LOSS_MODEL).read())
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("to_input_example", "select_single_example")]
    ns = {"torch": torch, "random": random, "InputExample": rs.InputExample}
    ns.update({k: getattr(__import__("typing"), k) for k in ("Tuple", "Any", "Optional", "List", "Dict", "Union")})
    ns.update({k: v for k, v in rs._constants().items() if k.isupper()})
    for node in wanted:
        exec(compile(ast.Module(body=[node], type_ignores=[]), rs.LOSS_MODEL, "exec"), ns)
    instance = {"reference": "r", "positive": "p", "part_positive": "pp", "negative": "n"}
    example = ns["to_input_example"](dict(instance))
    assert example.texts == ["r", "p", "pp", "n"] == [instance[k] for k in QUADRUPLET_KEYS]
    assert _texts_of(instance) == example.texts and _texts_of(example) == example.texts
    assert _texts_of((example, torch.tensor(0))) == example.texts and _texts_of((instance, 0)) == example.texts
