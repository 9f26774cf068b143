"""Run the reference's OWN in-repo evaluator code on the CPU (TEST INFRASTRUCTURE, authoring container only).

``/root/reference/models/evaluators.py`` cannot be imported: it needs ``sentence-transformers==2.2.2`` (absent
from this image) and downloads a cross-encoder at import time (``:31``).  The pieces of it that lie on the
SURVEY section-8 path and are the reference's own code -- ``euclidean_score`` (``:392-405``),
``QuadrupletLossEvaluator`` (``:34-128``), ``QuadrupletEvaluator`` (``:130-389``) -- are therefore lifted out
of the file with ``ast`` AT TEST TIME (nothing is copied into this repository) and executed in a namespace in
which only the third-party names are stand-ins:

* ``TripletEvaluator`` (sentence-transformers) -> ``ScriptedTriplet``: records its constructor arguments and
  returns the accuracy the test scripted for it (what the real one computes is restated in
  ``oracle/quad_eval_oracle.py`` and stays unpinned);
* ``QuadrupletSentenceTransformerLossModel`` -> ``ScriptedLossModel``: returns the per-batch loss tensors the
  test scripted (the loss arithmetic itself is pinned by ``tests/golden/loss_golden.npz``);
* ``SentenceEvaluator`` -> ``object``, ``InputExample`` -> a two-field class, ``tqdm`` / ``autocast`` /
  ``batch_to_device`` -> no-ops;
* for ``create_ir_evaluation_set`` / ``get_sequential_evaluator`` (``:406-614``): sentence-transformers'
  ``InformationRetrievalEvaluator`` / ``SequentialEvaluator`` -> ``RecordingEvaluator`` (keeps the constructor
  arguments), ``generate_variations`` (text augmentation of the query) -> identity.

``load(..., real_loss_model=True)`` puts the reference's own ``QuadrupletSentenceTransformerLossModel`` (lifted
from models/quadruplet_sentence_transformer.py the same way) in place of ``ScriptedLossModel``; with the
reference's loss module and ``TableSentenceModel`` (texts = rows of an embedding table) nothing on the loss
side is scripted any more.

Everything else the lifted code executes (sampling with ``random``, the 5-epoch re-sampling, the global
accuracy formula, the incremental mean in tensor arithmetic, CSV and JSON writing) is the reference's.
"""
import ast
import contextlib
import csv
import enum
import json
import logging
import os
import random
import typing

import numpy as np
import torch

REFERENCE_ROOT = "/root/reference"
EVALUATORS = os.path.join(REFERENCE_ROOT, "models", "evaluators.py")
CONSTANTS = os.path.join(REFERENCE_ROOT, "dataset", "constants.py")


def available() -> bool:
    return os.path.isfile(EVALUATORS) and os.path.isfile(CONSTANTS)


class InputExample:
    """Stand-in for sentence_transformers.InputExample (``texts`` + ``label``)."""

    def __init__(self, guid="", texts=None, label=0):
        self.guid, self.texts, self.label = guid, texts, label


class SimilarityFunction(enum.Enum):
    """Members and values of sentence_transformers.evaluation.SimilarityFunction (2.2.2)."""
    COSINE = 0
    EUCLIDEAN = 1
    MANHATTAN = 2
    DOT_PRODUCT = 3


class ScriptedTriplet:
    """Stand-in for TripletEvaluator: ``script[name]`` is the list of accuracies its calls return, in order."""
    script = {}
    built = []

    def __init__(self, anchors, positives, negatives, main_distance_function=None, name="", batch_size=16,
                 show_progress_bar=False, write_csv=True):
        self.anchors, self.positives, self.negatives, self.name = anchors, positives, negatives, name
        self.main_distance_function, self.write_csv = main_distance_function, write_csv
        ScriptedTriplet.built.append(self)

    def __call__(self, model, output_path=None, epoch=-1, steps=-1):
        return ScriptedTriplet.script[self.name].pop(0)


class ScriptedLossModel:
    """Stand-in for QuadrupletSentenceTransformerLossModel: returns the scripted batch losses in order."""
    losses = []
    seen_batches = []

    def __init__(self, st_model, quadruplet_loss, additional_model_kwargs=None, additional_loss_kwargs=None):
        pass

    def __call__(self, features, labels):
        ScriptedLossModel.seen_batches.append(len(labels))
        return ScriptedLossModel.losses.pop(0)


class StubSentenceModel:
    """What the lifted QuadrupletLossEvaluator touches of a SentenceTransformer."""
    device = torch.device("cpu")

    @staticmethod
    def smart_batching_collate(batch):
        return [batch], torch.zeros(len(batch))


class RecordingEvaluator:
    """Stand-in for sentence-transformers' InformationRetrievalEvaluator / SequentialEvaluator: keeps what the
    reference's code hands to the constructor."""
    built = []

    def __init__(self, *args, **kwargs):
        self.args, self.kwargs = args, kwargs
        RecordingEvaluator.built.append(self)


def cos_sim_marker(a, b):
    raise AssertionError("stand-in for sentence_transformers.util.cos_sim: never called in the sandbox")


def dot_score_marker(a, b):
    raise AssertionError("stand-in for sentence_transformers.util.dot_score: never called in the sandbox")


class _Bar:
    def __init__(self, it=None, **kw):
        self._it = it

    def __iter__(self):
        return iter(self._it)

    def update(self, n=1):
        pass

    def set_description(self, desc=None):
        pass

    def close(self):
        pass


def _constants():
    ns = {"Final": typing.Final, "final": typing.final, "os": os}
    tree = ast.parse(open(CONSTANTS).read())
    body = [n for n in tree.body if isinstance(n, (ast.Assign, ast.AnnAssign, ast.Import, ast.ImportFrom))]
    exec(compile(ast.Module(body=body, type_ignores=[]), CONSTANTS, "exec"), ns)
    return ns


LOSS_MODEL = os.path.join(REFERENCE_ROOT, "models", "quadruplet_sentence_transformer.py")


def load_loss_model():
    """The reference's ``QuadrupletSentenceTransformerLossModel`` (models/quadruplet_sentence_transformer.py:9-78):
    runs the sentence model on the four texts of a batch and calls the quadruplet loss with keywords."""
    consts = _constants()
    ns = {"torch": torch, "random": random, "Tuple": typing.Tuple, "Any": typing.Any, "Optional": typing.Optional,
          "List": typing.List, "Dict": typing.Dict, "Union": typing.Union, "SentenceTransformer": object,
          "InputExample": InputExample, "QuadrupletLoss": object}
    for key in ("REFERENCE_EXAMPLE", "POS_EXAMPLES", "PART_POS_EXAMPLES", "NEG_EXAMPLES"):
        ns[key] = consts[key]
    tree = ast.parse(open(LOSS_MODEL).read())
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "QuadrupletSentenceTransformerLossModel")
    exec(compile(ast.Module(body=[node], type_ignores=[]), LOSS_MODEL, "exec"), ns)
    return ns["QuadrupletSentenceTransformerLossModel"]


class TableSentenceModel:
    """Stand-in for a SentenceTransformer whose 'texts' are row ids of an embedding table: what the lifted
    QuadrupletLossEvaluator and loss model touch (``device``, ``smart_batching_collate``, ``__call__`` ->
    ``{'sentence_embedding': ...}``).  Dataset items are 4-tuples of row ids."""

    def __init__(self, table: torch.Tensor):
        self.table, self.device = table, table.device

    def smart_batching_collate(self, batch):
        cols = list(zip(*batch))
        return [torch.tensor(c, dtype=torch.long) for c in cols], torch.zeros(len(batch))

    def __call__(self, features, **kwargs):
        return {"sentence_embedding": self.table[features]}


def load(*names, real_loss_model: bool = False):
    """Namespace in which the named top-level definitions of models/evaluators.py have been executed.
    ``real_loss_model``: the reference's own loss-model wrapper instead of ``ScriptedLossModel``."""
    consts = _constants()
    ns = {
        "csv": csv, "json": json, "logging": logging, "os": os, "random": random, "np": np, "torch": torch,
        "Optional": typing.Optional, "List": typing.List, "Callable": typing.Callable, "Dict": typing.Dict,
        "Union": typing.Union, "Set": typing.Set, "final": typing.final, "Final": typing.Final,
        "Tensor": torch.Tensor, "DataLoader": torch.utils.data.DataLoader,
        "SentenceTransformer": object, "SentenceEvaluator": object, "QuadrupletDataset": object,
        "QuadrupletLoss": object, "InputExample": InputExample, "SimilarityFunction": SimilarityFunction,
        "TripletEvaluator": ScriptedTriplet, "QuadrupletSentenceTransformerLossModel": ScriptedLossModel,
        "tqdm": _Bar, "autocast": contextlib.nullcontext, "batch_to_device": lambda batch, device: batch,
        "LOGGER": logging.getLogger("reference.models.evaluators"),
        # create_ir_evaluation_set / get_sequential_evaluator (:406-614)
        "Subset": torch.utils.data.Subset, "GammaQuadrupletLoss": object,
        "InformationRetrievalEvaluator": RecordingEvaluator, "SequentialEvaluator": RecordingEvaluator,
        "cos_sim": cos_sim_marker, "dot_score": dot_score_marker,
        "generate_variations": lambda sentence, n=1: [sentence],      # dataset/sentence_compr_dataset_creation.py (text augmentation)
    }
    for key in ("REFERENCE_EXAMPLE", "POS_EXAMPLES", "PART_POS_EXAMPLES", "NEG_EXAMPLES", "RANDOM_SEED"):
        ns[key] = consts[key]
    if real_loss_model:
        ns["QuadrupletSentenceTransformerLossModel"] = load_loss_model()
    tree = ast.parse(open(EVALUATORS).read())
    # the module-level constants of the file (IR_EVALUATION_PATH, N_IR_SAMPLES, SIMILARITY_THRESHOLD)
    consts_here = [n for n in tree.body if isinstance(n, ast.AnnAssign) and isinstance(n.target, ast.Name)
                   and n.target.id.isupper()]
    exec(compile(ast.Module(body=consts_here, type_ignores=[]), EVALUATORS, "exec"), ns)
    found = {n.name: n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef))}
    for name in names:
        exec(compile(ast.Module(body=[found[name]], type_ignores=[]), EVALUATORS, "exec"), ns)
    return ns
