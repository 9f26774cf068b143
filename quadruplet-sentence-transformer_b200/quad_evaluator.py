"""Drop-in ``QuadrupletEvaluator``.

The reference's ``QuadrupletEvaluator`` (``/root/reference/models/evaluators.py:130-389``) wraps three
sentence-transformers 2.2.2 ``TripletEvaluator`` objects -- (anchor, pos, part), (anchor, pos, neg),
(anchor, part, neg) -- each of which encodes its sentences again, computes sklearn paired cosine /
manhattan / euclidean distances on the CPU and counts ``d(anchor, first) < d(anchor, second)``; the
three accuracies are combined as ``(((1-gamma)*pos_part + gamma*part_neg) + pos_neg) / 2`` (``:367``).

Here the four sentence lists are encoded ONCE, the embeddings stay on the GPU and one kernel
(``qst_quadruplet_eval``) produces all nine comparison counts.  Constructor, ``from_input_examples``,
``__call__(model, output_path, epoch, steps) -> float``, the CSV files and their columns are the
reference's; the re-sampling of examples every 5 epochs (``:266-345``) is kept.
"""
from __future__ import annotations

import csv
import logging
import os
import random
from enum import Enum
from typing import List, Optional

import torch

from . import _lib

logger = logging.getLogger(__name__)

# dataset item keys (/root/reference/dataset/constants.py:21-24)
REFERENCE_EXAMPLE, POS_EXAMPLES, NEG_EXAMPLES, PART_POS_EXAMPLES = "reference", "positive", "negative", "part_positive"


class SimilarityFunction(Enum):
    """Same members and values as sentence_transformers.evaluation.SimilarityFunction."""
    COSINE = 0
    EUCLIDEAN = 1
    MANHATTAN = 2
    DOT_PRODUCT = 3


_PAIRS = ("pos_part", "pos_neg", "part_neg")


def normalize_similarity_function(f):
    """``main_distance_function`` as the local enum.  The reference call sites pass
    ``sentence_transformers.evaluation.SimilarityFunction`` members (``models/evaluators.py:10, 148``):
    a foreign enum never compares equal to ours, so members are matched by ``.name`` (then ``.value``),
    plain strings / ints are accepted too, and anything unknown raises instead of silently falling
    through to "best of the three"."""
    if f is None or isinstance(f, SimilarityFunction):
        return f
    name = getattr(f, "name", f if isinstance(f, str) else None)
    if isinstance(name, str) and name.upper() in SimilarityFunction.__members__:
        return SimilarityFunction[name.upper()]
    value = getattr(f, "value", f)
    if isinstance(value, int) and not isinstance(value, bool):
        try:
            return SimilarityFunction(value)
        except ValueError:
            pass
    raise ValueError(f"main_distance_function must be None or a SimilarityFunction (COSINE, EUCLIDEAN, MANHATTAN, "
                     f"DOT_PRODUCT), {f!r} given")


def paired_distance_counts(anchor: torch.Tensor, pos: torch.Tensor, part: torch.Tensor, neg: torch.Tensor,
                           want_distances: bool = False):
    """``qst_quadruplet_eval``: counts [3 metrics (cos, manhattan, euclid), 3 pairs] as a CPU int64
    tensor (and the [B, 9] distances when asked)."""
    lib = _lib.load()
    xs = [anchor, pos, part, neg]
    _lib.require_cuda(*xs)
    dt = xs[0].dtype if xs[0].dtype in (torch.float32, torch.float16, torch.bfloat16) else torch.float32
    xs = [x.to(dt).contiguous() for x in xs]
    B, D = xs[0].shape
    dev = xs[0].device
    with torch.cuda.device(dev):
        counts = torch.empty(9, dtype=torch.int64, device=dev)
        dist = torch.empty((B, 9), dtype=torch.float32, device=dev) if want_distances else None
        _lib.check(lib.qst_quadruplet_eval(xs[0].data_ptr(), xs[1].data_ptr(), xs[2].data_ptr(), xs[3].data_ptr(),
                                           _lib.dtype_code(dt), B, D, _lib.ptr(dist), counts.data_ptr(),
                                           _lib.stream_ptr(dev)))
    counts = counts.cpu().view(3, 3)
    return (counts, dist) if want_distances else counts


def _sample_quadruplets(examples):
    """One (anchor, positive, partial positive, negative) per dataset item, as
    ``models/evaluators.py:224-262`` samples them (InputExample-like objects or dict items whose
    positive / part_positive / negative entries may be lists to draw from)."""
    cols = ([], [], [], [])
    for example in examples:
        if isinstance(example, tuple):
            example = example[0]
        texts = getattr(example, "texts", None)
        if texts is None:
            texts = [example[REFERENCE_EXAMPLE]]
            for key in (POS_EXAMPLES, PART_POS_EXAMPLES, NEG_EXAMPLES):
                v = example[key]
                texts.append(v[random.randint(0, len(v) - 1)] if isinstance(v, list) else v)
        for col, t in zip(cols, texts):
            col.append(t)
    return cols


class QuadrupletEvaluator:
    N_EPOCHS_RESET_EXAMPLES = 5  # number of epochs after which examples are re-sampled

    def __init__(self, anchors: List[str], positives: List[str], partially_positives: List[str],
                 negatives: List[str], gamma: float = 0.6, main_distance_function: SimilarityFunction = None,
                 name: str = "", batch_size: int = 16, show_progress_bar: bool = False, write_csv: bool = True,
                 all_examples=None):
        assert len(anchors) == len(positives) == len(partially_positives) == len(negatives)
        self.anchors, self.positives = anchors, positives
        self.partially_positives, self.negatives = partially_positives, negatives
        self.name = name
        self._gamma = gamma
        self._all_examples = all_examples
        self.main_distance_function = normalize_similarity_function(main_distance_function)
        self.batch_size = batch_size
        if show_progress_bar is None:     # models/evaluators.py:179-183: follow the logger's level
            show_progress_bar = logger.getEffectiveLevel() in (logging.INFO, logging.DEBUG)
        self.show_progress_bar = bool(show_progress_bar)
        self.write_csv = write_csv
        self.csv_file = "quadruplet_evaluation" + ("_" + name if name else "") + "_results.csv"
        self.csv_headers = ["epoch", "steps", "pos_part_accuracy", "pos_neg_accuracy", "part_neg_accuracy",
                            "global_accuracy"]
        # the three inner TripletEvaluators of the reference write their own files
        self.triplet_csv_files = {p: "triplet_evaluation_" + p + "_results.csv" for p in _PAIRS}
        self.triplet_csv_headers = ["epoch", "steps", "accuracy_cosinus", "accuracy_manhattan", "accuracy_euclidean"]
        self._epoch_counter = 0
        self.last_accuracies = None

    @classmethod
    def from_input_examples(cls, examples, **kwargs):
        a, p, pp, n = _sample_quadruplets(examples)
        return cls(a, p, pp, n, all_examples=examples, **kwargs)

    def _reset_examples(self):
        self._epoch_counter += 1
        if self._all_examples is not None and self._epoch_counter % self.N_EPOCHS_RESET_EXAMPLES == 0:
            self.anchors, self.positives, self.partially_positives, self.negatives = \
                _sample_quadruplets(self._all_examples)

    def _encode(self, model, sentences):
        emb = model.encode(sentences, batch_size=self.batch_size, show_progress_bar=self.show_progress_bar,
                           convert_to_tensor=True)
        if not isinstance(emb, torch.Tensor):
            emb = torch.as_tensor(emb)
        if not emb.is_cuda:
            if not torch.cuda.is_available():
                raise _lib.QstError("QuadrupletEvaluator needs a CUDA device (no CPU fallback)")
            emb = emb.cuda()
        return emb

    def _pick(self, acc_cos: float, acc_manhattan: float, acc_euclid: float) -> float:
        """Return value of one TripletEvaluator call (ST 2.2.2)."""
        f = normalize_similarity_function(self.main_distance_function)   # the attribute may be re-assigned
        if f == SimilarityFunction.COSINE:
            return acc_cos
        if f == SimilarityFunction.MANHATTAN:
            return acc_manhattan
        if f == SimilarityFunction.EUCLIDEAN:
            return acc_euclid
        return max(acc_cos, acc_manhattan, acc_euclid)

    def __call__(self, model, output_path: str = None, epoch: int = -1, steps: int = -1) -> float:
        self._reset_examples()
        embs = [self._encode(model, s) for s in (self.anchors, self.positives, self.partially_positives,
                                                 self.negatives)]
        counts = paired_distance_counts(*embs)                    # [metric, pair]
        n = len(self.anchors)
        acc = {}
        for j, pair in enumerate(_PAIRS):
            a_cos, a_man, a_euc = (int(counts[m, j]) / n for m in range(3))
            acc[pair] = self._pick(a_cos, a_man, a_euc)
            if output_path is not None and self.write_csv:
                self._append_csv(os.path.join(output_path, self.triplet_csv_files[pair]), self.triplet_csv_headers,
                                 [epoch, steps, a_cos, a_man, a_euc])
        g = self._gamma
        glob_accuracy = (((1 - g) * acc["pos_part"] + g * acc["part_neg"]) + acc["pos_neg"]) / 2
        self.last_accuracies = dict(acc, global_accuracy=glob_accuracy)
        logger.info("Pos-Part Accuracy Distance:   \t{:.2f}".format(acc["pos_part"] * 100))
        logger.info("Pos-Neg Accuracy Distance:   \t{:.2f}".format(acc["pos_neg"] * 100))
        logger.info("Part-Neg Distance:   \t{:.2f}".format(acc["part_neg"] * 100))
        logger.info("Accuracy Distance:   \t{:.2f}".format(glob_accuracy * 100))
        if output_path is not None and self.write_csv:
            self._append_csv(os.path.join(output_path, self.csv_file), self.csv_headers,
                             [epoch, steps, acc["pos_part"], acc["pos_neg"], acc["part_neg"], glob_accuracy])
        return glob_accuracy

    @staticmethod
    def _append_csv(path: str, headers: list, row: list):
        new = not os.path.isfile(path)
        with open(path, newline="", mode="w" if new else "a", encoding="utf-8") as f:
            w = csv.writer(f)
            if new:
                w.writerow(headers)
            w.writerow(row)
