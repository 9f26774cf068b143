"""Runs the reference's UNMODIFIED ``ir_evauation_script.py`` against the drop-in (TEST INFRASTRUCTURE; run as a
subprocess by tests/test_reference_script.py, authoring container only).

What INTEGRATION.md section A promises -- "swap the imports, nothing else" -- executed: a stand-in
``sentence_transformers`` package whose ``evaluation.InformationRetrievalEvaluator`` and ``util.cos_sim`` /
``util.dot_score`` ARE this package's, put in front of the reference's own modules.  The script is imported from
``/root/reference`` as it is (its ``models.evaluators`` with ``create_ir_evaluation_set`` and ``euclidean_score``
too), its own argparse block builds the arguments, and ``main(args)`` runs: dataset split, evaluation-set
creation, evaluator construction with the script's keyword set and score-function table, output folders --
up to the first ``evaluator(model=..., output_path=...)``, where the drop-in asks for a CUDA device (there is
no CPU fallback; on a GPU box the same call is what tests/test_gpu_scoring.py exercises).  Stand-ins beyond
sentence-transformers: the dataset class (needs the COCO chunk files) and the packages the reference imports
for dataset creation (nlpaug, nltk, openai, sortedcollections, torchvision), none of them on the path.
``--oracle``: the CPU oracle's evaluator and score functions stand in instead of the drop-in's; the script then
runs to completion on the CPU (the oracle restatement must offer everything the script asks of ST 2.2.2).
Prints one JSON line prefixed with ``RESULT ``."""
import ast
import importlib
import importlib.abc
import importlib.machinery
import json
import os
import random
import sys
import tempfile
import types
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE_ROOT = "/root/reference"
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True         # the reference's modules are imported from where they lie: leave no __pycache__ there

import torch  # noqa: E402

import qst_b200  # noqa: E402


class _DatasetCreationStubs(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    ROOTS = ("nlpaug", "openai", "nltk", "sortedcollections", "torchvision")

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in self.ROOTS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = mock.MagicMock(name=spec.name)
        m.__path__, m.__spec__, m.__name__ = [], spec, spec.name
        return m

    def exec_module(self, module):
        pass


class SentenceTransformer:
    """Stand-in sentence model with HOST embeddings (the drop-in stops at its device check; the oracle
    evaluator runs through): "anchor i" is a seeded Gaussian row, "pos i.j" / "part i.j" lie at 1.0 / 2.0 noise
    from it, "neg i.j" anywhere; two model names give two different tables."""
    DIM = 32

    def __init__(self, name, device=None):
        self.name, self.device = name, device
        self.salt = sum(ord(ch) for ch in name)

    def _row(self, text):
        kind, rest = text.split(" ", 1)
        i = int(rest.split(".")[0])
        base = torch.randn(self.DIM, generator=torch.Generator().manual_seed(1000 * self.salt + i))
        if kind == "anchor":
            return base
        noise = torch.randn(self.DIM, generator=torch.Generator().manual_seed(
            7_000_000 + 1000 * self.salt + sum(ord(ch) * (n + 1) for n, ch in enumerate(text))))
        return {"pos": base + 1.0 * noise, "part": base + 2.0 * noise, "neg": 1.5 * noise}[kind]

    def encode(self, sentences, **kw):
        return torch.stack([self._row(t) for t in sentences]) if len(sentences) else torch.zeros(0, self.DIM)


class InputExample:
    def __init__(self, guid="", texts=None, label=0):
        self.guid, self.texts, self.label = guid, texts, label


class CrossEncoder:
    def __init__(self, name):
        self.name = name


class SequentialEvaluator:
    def __init__(self, evaluators, main_score_function=None):
        self.evaluators = evaluators


def install_sentence_transformers_shim(use_oracle: bool = False):
    """``use_oracle``: the CPU oracle's evaluator and score functions instead of the drop-in's (same script, run
    to completion on the CPU)."""
    st = types.ModuleType("sentence_transformers")
    st.__path__ = []
    ev = types.ModuleType("sentence_transformers.evaluation")
    ut = types.ModuleType("sentence_transformers.util")
    st.SentenceTransformer, st.InputExample, st.CrossEncoder, st.util, st.evaluation = \
        SentenceTransformer, InputExample, CrossEncoder, ut, ev
    ev.SentenceEvaluator, ev.TripletEvaluator, ev.SequentialEvaluator = object, object, SequentialEvaluator
    ev.SimilarityFunction = qst_b200.SimilarityFunction
    ev.InformationRetrievalEvaluator = qst_b200.InformationRetrievalEvaluator          # <- the swap
    ut.cos_sim, ut.dot_score = qst_b200.cos_sim, qst_b200.dot_score                    # <- the swap
    if use_oracle:
        from oracle import ir_oracle
        ev.InformationRetrievalEvaluator = ir_oracle.InformationRetrievalEvaluatorOracle
        ut.cos_sim, ut.dot_score = ir_oracle.cos_sim, ir_oracle.dot_score
    ut.batch_to_device = lambda batch, device: batch
    sys.modules.update({"sentence_transformers": st, "sentence_transformers.evaluation": ev,
                        "sentence_transformers.util": ut})


class FakeQuadrupletDataset:
    """200 items of the shape dataset/quadruplet_dataset.py yields (the real class reads COCO chunk files)."""

    def __init__(self, path, *chunks, hard_contrastive_mode=None, n_pos=1, n_neg=1, n_part_pos=1, cache_size=None,
                 transform=None):
        self.items = [{"reference": f"anchor {i}", "positive": [f"pos {i}.{j}" for j in range(n_pos)],
                       "part_positive": [f"part {i}.{j}" for j in range(n_part_pos)],
                       "negative": [f"neg {i}.{j}" for j in range(n_neg)]} for i in range(200)]

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        return self.items[i]


def main():
    use_oracle = "--oracle" in sys.argv[1:]
    sys.meta_path.insert(0, _DatasetCreationStubs())
    install_sentence_transformers_shim(use_oracle)
    sys.path.insert(0, REFERENCE_ROOT)
    script = importlib.import_module("ir_evauation_script")            # the reference's file, as it is
    import models.evaluators as reference_evaluators                   # noqa: E402  (the reference's module)
    reference_evaluators.generate_variations = lambda sentence, n=1: [sentence]     # text augmentation (nlpaug)
    script.QuadrupletDataset = FakeQuadrupletDataset

    built = []

    class Recording(script.InformationRetrievalEvaluator):
        def __init__(self, *a, **k):
            super().__init__(*a, **k)
            built.append((self, k))
            self.returned = []

        def __call__(self, *a, **k):
            self.returned.append(super().__call__(*a, **k))
            return self.returned[-1]

    script.InformationRetrievalEvaluator = Recording

    # the script's own argparse block (everything under `if __name__ == '__main__':` but parse + main call)
    tree = ast.parse(open(os.path.join(REFERENCE_ROOT, "ir_evauation_script.py")).read())
    block = next(n for n in tree.body if isinstance(n, ast.If))
    stmts = [n for n in block.body
             if not (isinstance(n, ast.Assign) and getattr(n.targets[0], "id", "") == "arguments")
             and not (isinstance(n, ast.Expr) and isinstance(n.value, ast.Call) and getattr(n.value.func, "id", "") == "main")]
    ns = dict(vars(script))
    exec(compile(ast.Module(body=stmts, type_ignores=[]), "ir_evauation_script.py", "exec"), ns)

    result = {"bound_evaluator": script.InformationRetrievalEvaluator.__mro__[1].__module__,
              "bound_cos_sim": script.cos_sim.__module__, "bound_euclidean_score": script.euclidean_score.__module__}
    with tempfile.TemporaryDirectory() as tmp:
        torch.save(3, os.path.join(tmp, "chunk_n.pt"))
        args = ns["parser"].parse_args(["--dataset_path_train", tmp, "--out_path", os.path.join(tmp, "out"),
                                        "--evaluation_queries_path", os.path.join(tmp, "absent.json")])
        random.seed(14)
        torch.manual_seed(14)
        try:
            script.main(args)
            result["stopped_at"] = None
        except Exception as e:  # noqa: BLE001
            result["stopped_at"] = [type(e).__name__, str(e)]
        written = sorted(f for _, _, files in os.walk(os.path.join(tmp, "out")) for f in files)
        csv_text = None
        for folder, _, files in os.walk(os.path.join(tmp, "out")):
            for f in files:
                if f.endswith("_results.csv"):
                    csv_text = open(os.path.join(folder, f)).read()
    ev, kwargs = built[0]
    result["csv_text"], result["returned"] = csv_text, [float(v) for v in ev.returned]
    result.update({
        "evaluators_built": len(built), "keywords": sorted(kwargs), "score_function_names": ev.score_function_names,
        "score_function_modules": {n: f.__module__ for n, f in ev.score_functions.items()},
        "max_k": ev.max_k, "csv_columns": len(ev.csv_headers), "csv_file": ev.csv_file,
        "queries": len(ev.queries), "corpus": len(ev.corpus), "write_csv": ev.write_csv,
        "corpus_chunk_size": ev.corpus_chunk_size, "files_written": written,
        "relevant_per_query": sorted({len(ev.relevant_docs[q]) for q in ev.queries_ids}),
    })
    print("RESULT " + json.dumps(result))


if __name__ == "__main__":
    main()
