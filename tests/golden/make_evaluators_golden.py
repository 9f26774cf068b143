"""Generate tests/golden/evaluators_golden.json from the REFERENCE's own evaluator code.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_evaluators_golden.py

``tests/reference_sandbox.py`` lifts ``euclidean_score``, ``QuadrupletEvaluator`` and
``QuadrupletLossEvaluator`` out of ``/root/reference/models/evaluators.py`` and runs them with the
third-party pieces scripted (see its docstring).  Stored here: the scripted inputs and what the reference
code made of them -- sampled quadruplets for a seeded ``random``, return values and CSV text of
``QuadrupletEvaluator.__call__`` (``:345-389``), return values and JSON text of
``QuadrupletLossEvaluator.__call__`` (``:49-128``), ``euclidean_score`` (``:392-405``) on a small
seeded case, and the evaluation-set file ``create_ir_evaluation_set`` (``:406-530``) writes together with
what its two reload paths (its own, and ``get_sequential_evaluator``'s ``:556-561``) make of it.  Seed 14 is
the reference's RANDOM_SEED (``dataset/constants.py:5``).
"""
import json
import os
import random
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import reference_sandbox as rs  # noqa: E402


LIFTED = ("euclidean_score", "QuadrupletLossEvaluator", "QuadrupletEvaluator", "create_ir_evaluation_set",
          "get_sequential_evaluator")


def dict_examples(n=12):
    """Dataset items of the shape dataset/quadruplet_dataset.py yields: lists to draw from or plain strings."""
    out = []
    for i in range(n):
        out.append({"reference": f"anchor {i}",
                    "positive": [f"pos {i}.{j}" for j in range(1 + i % 4)],
                    "part_positive": [f"part {i}.{j}" for j in range(1 + (i * 3) % 5)] if i % 3 else f"part {i}",
                    "negative": [f"neg {i}.{j}" for j in range(2 + i % 3)]})
    return out


TRIPLET_SCRIPT = {  # accuracy returned by each inner TripletEvaluator, per call
    "pos_part": [0.61, 0.7283950617283951, 1.0, 0.0, 1 / 3],
    "pos_neg": [0.93, 0.9753086419753086, 1.0, 0.0, 2 / 3],
    "part_neg": [0.88, 0.8641975308641975, 1.0, 0.0, 0.1],
}
CALLS = [(-1, -1), (0, -1), (0, 500), (1, -1), (2, 10)]           # (epoch, steps)
GAMMAS = [0.6, 0.0, 1.0, 0.25]
BATCH_LOSSES = [1.7320508, 0.25, 3.1415927, 0.0, 2.5e-7, 12.75, 0.333333343, 1.0, 0.99999994, 7.0]


def run_quadruplet_evaluator(ns, gamma):
    rs.ScriptedTriplet.script = {k: list(v) for k, v in TRIPLET_SCRIPT.items()}
    ev = ns["QuadrupletEvaluator"](["a"], ["p"], ["pp"], ["n"], gamma=gamma, name="val")
    with tempfile.TemporaryDirectory() as tmp:
        returned = [ev(None, output_path=tmp, epoch=e, steps=s) for e, s in CALLS]
        text = open(os.path.join(tmp, ev.csv_file), newline="", encoding="utf-8").read()
    return {"gamma": gamma, "returned": returned, "csv_file": ev.csv_file, "csv_text": text}


def run_loss_evaluator(ns, n_items, batch_size):
    n_batches = -(-n_items // batch_size)
    rs.ScriptedLossModel.losses = [torch.tensor(v, dtype=torch.float32) for v in BATCH_LOSSES[:n_batches]]
    rs.ScriptedLossModel.seen_batches = []
    ev = ns["QuadrupletLossEvaluator"](list(range(n_items)), None, batch_size=batch_size)
    with tempfile.TemporaryDirectory() as tmp:
        first = ev(rs.StubSentenceModel(), output_path=tmp, epoch=0, steps=-1)
        rs.ScriptedLossModel.losses = [torch.tensor(v, dtype=torch.float32) for v in reversed(BATCH_LOSSES[:n_batches])]
        second = ev(rs.StubSentenceModel(), output_path=tmp, epoch=1, steps=40)
        text = open(os.path.join(tmp, "_quadruplet_loss_eval.json")).read()
    return {"n_items": n_items, "batch_size": batch_size, "batch_sizes_seen": rs.ScriptedLossModel.seen_batches[:n_batches],
            "returned": [float(first), float(second)], "returned_dtype": str(first.dtype), "json_text": text}


def list_examples(n=12):
    """Items as create_ir_evaluation_set iterates them (:455-490): every example field is a list."""
    out = dict_examples(n)
    for item in out:
        if isinstance(item["part_positive"], str):
            item["part_positive"] = [item["part_positive"]]
    return out


class LossWithGamma:
    gamma = 0.6            # get_sequential_evaluator reads loss.gamma (:593)


def run_ir_evaluation_set(ns, use_pos, use_part_pos, add_part_pos_corpus):
    """create_ir_evaluation_set (:406-530) on a seeded dataset, then the reload paths: its own (:414-431,
    per-query sets) and get_sequential_evaluator's (:556-561, the set of ALL query ids for every query),
    whose result is what the reference hands to InformationRetrievalEvaluator (:572-588)."""
    random.seed(14)
    with tempfile.TemporaryDirectory() as tmp:
        path = os.path.join(tmp, "created_eval_queries.json")
        made = ns["create_ir_evaluation_set"](list_examples(), n_queries=5, out_path=path, use_pos=use_pos,
                                              use_part_pos=use_part_pos, use_cross_encoder=False,
                                              add_part_pos_corpus=add_part_pos_corpus)
        text = open(path).read()
        again = ns["create_ir_evaluation_set"](list_examples(), out_path=path)          # file exists: reload branch
        rs.RecordingEvaluator.built = []
        ns["get_sequential_evaluator"](list_examples(), LossWithGamma(), evaluation_queries_path=path, name="val")
        ire, seq = rs.RecordingEvaluator.built
    assert made["queries"] == again["queries"] and made["relevant"] == again["relevant"]
    kw = dict(ire.kwargs)
    return {"flags": {"use_pos": use_pos, "use_part_pos": use_part_pos, "add_part_pos_corpus": add_part_pos_corpus},
            "file_text": text,
            "queries": made["queries"], "corpus": made["corpus"],
            "relevant": {q: sorted(v) for q, v in made["relevant"].items()},
            "relevant_after_sequential_evaluator_reload": {q: sorted(v) for q, v in kw["relevant_docs"].items()},
            "ire_kwargs": {k: v for k, v in kw.items() if k not in ("queries", "corpus", "relevant_docs", "score_functions")},
            "ire_score_function_names": sorted(kw["score_functions"]),
            "sequential_order": [type(e).__name__ for e in seq.kwargs["evaluators"]]}


FULL_STACK_CASES = [   # (loss kwargs, batch size) for the un-scripted run below
    (dict(gamma=0.6, margin_pos_neg=1.0, margin_pos_part=0.5, margin_part_neg=0.5), 16),
    (dict(gamma=0.3, margin_pos_neg=0.7, margin_pos_part=0.2, margin_part_neg=0.9, p=1.0, swap=True, reduction="sum"), 32),
    (dict(gamma=1.0, margin_pos_neg=0.1, margin_pos_part=2.0, margin_part_neg=0.5, p=3.0), 7),
]


def full_stack_table(n=70, d=48):
    g = torch.Generator().manual_seed(14)
    return n, torch.randn(4 * n, d, generator=g)


def run_full_loss_stack(loss_kwargs, batch_size):
    """Nothing scripted: the reference's QuadrupletLossEvaluator (:34-128) drives the reference's
    QuadrupletSentenceTransformerLossModel (models/quadruplet_sentence_transformer.py:9-78) and the reference's
    GammaQuadrupletLoss (models/losses/losses.py:241-303) over a table-lookup sentence model on the CPU."""
    from oracle import loss_oracle
    ns = rs.load("QuadrupletLossEvaluator", real_loss_model=True)
    n, table = full_stack_table()
    items = [(i, n + i, 2 * n + i, 3 * n + i) for i in range(n)]
    loss = loss_oracle.load_reference_losses().GammaQuadrupletLoss(**loss_kwargs)
    ev = ns["QuadrupletLossEvaluator"](items, loss, batch_size=batch_size)
    with tempfile.TemporaryDirectory() as tmp:
        return ev(rs.TableSentenceModel(table), output_path=tmp, epoch=0, steps=0)


def main():
    if not rs.available():
        raise SystemExit("/root/reference not present: golden vectors can only be made in the authoring container")
    ns = rs.load(*LIFTED)
    out = {"calls": CALLS, "triplet_script": TRIPLET_SCRIPT, "batch_losses": BATCH_LOSSES}

    # sampling (:224-262) and the re-sampling every 5 epochs (:264-343)
    examples = dict_examples()
    random.seed(14)
    ev = ns["QuadrupletEvaluator"].from_input_examples(examples, gamma=0.6, name="s")
    first = [ev.anchors, ev.positives, ev.partially_positives, ev.negatives]
    for _ in range(5):
        ev._reset_examples()
    out["sampling"] = {"seed": 14, "n": len(examples), "first": first,
                       "after_5_epochs": [ev.anchors, ev.positives, ev.partially_positives, ev.negatives]}

    out["quadruplet_evaluator"] = [run_quadruplet_evaluator(ns, g) for g in GAMMAS]
    out["loss_evaluator"] = [run_loss_evaluator(ns, n, b) for n, b in ((10, 1), (10, 3), (7, 32), (64, 8))]

    out["ir_evaluation_set"] = [run_ir_evaluation_set(ns, *flags) for flags in
                                ((True, True, True), (True, False, True), (False, True, True), (True, False, False))]

    out["full_loss_stack"] = [{"loss_kwargs": kw, "batch_size": bs, "average_loss": float(run_full_loss_stack(kw, bs))}
                              for kw, bs in FULL_STACK_CASES]

    g = torch.Generator().manual_seed(14)
    a, b = torch.randn(5, 12, generator=g), torch.randn(9, 12, generator=g)
    out["euclidean_score"] = {"a": a.tolist(), "b": b.tolist(), "scores": ns["euclidean_score"](a, b).tolist(),
                              "one_d": ns["euclidean_score"](a[0], b[1]).tolist(),
                              "from_lists": ns["euclidean_score"](a[:2].tolist(), b[:3].tolist()).tolist()}
    path = os.path.join(HERE, "evaluators_golden.json")
    with open(path, "w") as fp:
        json.dump(out, fp, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
