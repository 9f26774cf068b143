"""SURVEY 8f row 2, "ir_evauation_script.py runs unmodified against the drop-in", as far as a box without a GPU
can show it: tests/run_reference_script.py imports the reference's script from /root/reference as it is, with
sentence-transformers' evaluator and score functions replaced by this package's, and runs its ``main``."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(not os.path.isfile("/root/reference/ir_evauation_script.py"),
                    reason="the reference is only mounted in the authoring container")
def test_reference_script_runs_unmodified_up_to_the_device_boundary():
    run = subprocess.run([sys.executable, os.path.join(HERE, "run_reference_script.py")], capture_output=True,
                         text=True, timeout=600)
    lines = [ln for ln in run.stdout.splitlines() if ln.startswith("RESULT ")]
    assert run.returncode == 0 and len(lines) == 1, (run.returncode, run.stdout[-2000:], run.stderr[-2000:])
    r = json.loads(lines[0][len("RESULT "):])
    # the script's names are bound to the drop-in (evaluator, cos_sim) and to the reference's own euclidean_score
    assert r["bound_evaluator"] == "qst_b200.ir_evaluator" and r["bound_cos_sim"] == "qst_b200.scoring"
    assert r["bound_euclidean_score"] == "models.evaluators"
    # main() got through dataset split, create_ir_evaluation_set, output folders and evaluator construction ...
    assert r["evaluators_built"] == 1 and r["files_written"] == ["command_line_args.json", "created_eval_queries.json"]
    assert r["keywords"] == sorted(["queries", "corpus", "relevant_docs", "corpus_chunk_size", "mrr_at_k", "ndcg_at_k",
                                    "accuracy_at_k", "precision_recall_at_k", "map_at_k", "show_progress_bar",
                                    "batch_size", "write_csv", "score_functions", "main_score_function", "name"])
    # ... with the script's defaults: all three score functions (its own euclidean_score is recognised), k up to
    # 900, CSV on, 50 000-row chunks; 4 positives + 4 partial positives relevant per query (:36-43, :102-105)
    assert r["score_function_names"] == ["cos_sim", "dot_score", "euclid_score"]
    assert r["score_function_modules"]["euclid_score"] == "models.evaluators"
    assert r["max_k"] == 900 and r["csv_columns"] == 2 + 3 * 68 and r["write_csv"] is True
    assert r["corpus_chunk_size"] == 50000 and r["relevant_per_query"] == [8]
    assert r["csv_file"] == "Information-Retrieval_evaluation_trained_exp5_results.csv"
    assert r["queries"] >= 5 and r["corpus"] > 100
    # ... and stopped exactly where the device is needed: there is no CPU fallback
    assert r["stopped_at"] == ["QstError", "InformationRetrievalEvaluator needs a CUDA device (no CPU fallback)"]


@pytest.mark.skipif(not os.path.isfile("/root/reference/ir_evauation_script.py"),
                    reason="the reference is only mounted in the authoring container")
def test_reference_script_runs_to_completion_with_the_oracle_evaluator():
    """The same unmodified script with the CPU ORACLE standing in for sentence-transformers' evaluator and
    score functions (and the reference's own ``euclidean_score`` called for real): it runs to the end, i.e. the
    restatement has the constructor, call protocol and CSV layout the script needs from ST 2.2.2 -- baseline and
    model evaluated into one CSV (header + two rows of 2 + 3 * 68 cells), the returned value = the best
    score function's MAP@900, and with a 169-document corpus every relevant document is inside the top 900."""
    run = subprocess.run([sys.executable, os.path.join(HERE, "run_reference_script.py"), "--oracle"],
                         capture_output=True, text=True, timeout=600)
    lines = [ln for ln in run.stdout.splitlines() if ln.startswith("RESULT ")]
    assert run.returncode == 0 and len(lines) == 1, (run.returncode, run.stdout[-2000:], run.stderr[-2000:])
    r = json.loads(lines[0][len("RESULT "):])
    assert r["bound_evaluator"] == "oracle.ir_oracle" and r["stopped_at"] is None
    assert r["files_written"] == ["Information-Retrieval_evaluation_trained_exp5_results.csv", "command_line_args.json",
                                  "created_eval_queries.json"]
    rows = [ln.split(",") for ln in r["csv_text"].strip().splitlines()]
    assert len(rows) == 3 and [len(x) for x in rows] == [206, 206, 206]
    header = rows[0]
    assert header[:3] == ["epoch", "steps", "cos_sim-Accuracy@1"] and header[-1] == "euclid_score-MAP@900"
    assert len(r["returned"]) == 2
    for row, returned in zip(rows[1:], r["returned"]):
        cell = dict(zip(header, row))
        assert cell["epoch"] == "-1" and cell["steps"] == "-1"
        for fn in ("cos_sim", "dot_score", "euclid_score"):
            assert float(cell[f"{fn}-Recall@900"]) == 1.0 and float(cell[f"{fn}-Accuracy@900"]) == 1.0
            assert 0.0 < float(cell[f"{fn}-MAP@900"]) <= 1.0
            recalls = [float(cell[f"{fn}-Recall@{k}"]) for k in (1, 3, 5, 10, 20, 30, 40, 50, 100, 200, 500, 900)]
            assert recalls == sorted(recalls)                               # recall is monotone in k
        assert returned == max(float(cell[f"{fn}-MAP@900"]) for fn in ("cos_sim", "dot_score", "euclid_score"))
        assert 0.3 < returned < 1.0                                         # the stand-in embeddings are not trivial


def test_known_foreign_score_callables_select_the_fused_path():
    """sentence-transformers' ``util.cos_sim`` / ``util.dot_score`` and the reference's
    ``models.evaluators.euclidean_score`` are recognised by module and name (the unmodified reference hands
    exactly these over, ir_evauation_script.py:70); anything else is still refused."""
    import qst_b200
    from qst_b200 import scoring

    def foreign(module, name):
        def fn(a, b):
            raise AssertionError("the evaluator never calls a score function")
        fn.__module__, fn.__name__ = module, name
        return fn

    assert scoring.score_name_of(foreign("sentence_transformers.util", "cos_sim")) == "cos_sim"
    assert scoring.score_name_of(foreign("sentence_transformers.util", "dot_score")) == "dot_score"
    assert scoring.score_name_of(foreign("sentence_transformers.util.similarity", "cos_sim")) == "cos_sim"
    assert scoring.score_name_of(foreign("models.evaluators", "euclidean_score")) == "euclid_score"
    assert scoring.score_name_of(foreign("sentence_transformers.util", "manhattan_sim")) is None
    assert scoring.score_name_of(foreign("somewhere.else", "cos_sim")) is None
    assert scoring.score_name_of(qst_b200.euclidean_score) == "euclid_score"
    fns = {"cos_sim": foreign("sentence_transformers.util", "cos_sim"),
           "euclid_score": foreign("models.evaluators", "euclidean_score")}
    ev = qst_b200.InformationRetrievalEvaluator({"q": "0"}, {"d": "1"}, {"q": {"d"}}, score_functions=fns)
    assert ev.score_function_names == ["cos_sim", "euclid_score"]
    with pytest.raises(TypeError):
        qst_b200.InformationRetrievalEvaluator({"q": "0"}, {"d": "1"}, {"q": {"d"}},
                                               score_functions={"f": foreign("somewhere.else", "cos_sim")})
