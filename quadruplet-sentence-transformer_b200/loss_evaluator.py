"""Drop-in for the reference's ``QuadrupletLossEvaluator``.

Reference: ``/root/reference/models/evaluators.py:34-128``.  It walks a quadruplet dataset in batches,
runs the sentence model on the four texts of every instance, evaluates the quadruplet loss under
``torch.no_grad()`` (optionally ``autocast``) and keeps the incremental mean
``avg <- avg + 1/(i+1) * (loss_i - avg)`` (``:98``), which it returns and appends to
``<output_path>/_quadruplet_loss_eval.json``.

Here every batch loss is one launch of the fused loss kernel (``qst_quadruplet_fwd`` through the loss
module handed in); the per-batch scalars stay on the device in one buffer and are read back once, and
the incremental mean is replayed on the host in float32 with the reference's operation order, so the
returned value is the one the reference's tensor arithmetic produces for the same batch losses.
"""
from __future__ import annotations

import json
import os
from typing import Iterable, List, Optional, Sequence

import numpy as np
import torch

from . import _lib

# keys of a dict-shaped dataset instance (/root/reference/dataset/constants.py:22-25), in the order the
# reference's to_input_example() lays the texts out (models/quadruplet_sentence_transformer.py:86-110)
QUADRUPLET_KEYS = ("reference", "positive", "part_positive", "negative")
LOG_FILE = "_quadruplet_loss_eval.json"


def incremental_mean_f32(batch_losses: Sequence[float]) -> np.float32:
    """``models/evaluators.py:98`` replayed in float32: the running value starts as the Python float
    0.0, every later operand is a float32 tensor, and the Python scalar ``1/(i+1)`` enters the
    multiplication rounded to float32 (torch's scalar-operand rule)."""
    avg = np.float32(0.0)
    for i, loss in enumerate(batch_losses):
        step = np.float32(1 / (i + 1)) * (np.float32(loss) - avg)
        avg = np.float32(avg + step)
    return avg


def _texts_of(instance) -> List[str]:
    """The four texts (anchor, positive, partially positive, negative) of one dataset instance."""
    if isinstance(instance, tuple) and len(instance) == 2 and not isinstance(instance[0], str):
        instance = instance[0]                               # (instance, label) pairs
    texts = getattr(instance, "texts", None)                 # sentence-transformers InputExample
    if texts is None and isinstance(instance, dict):
        texts = []
        for key in QUADRUPLET_KEYS:
            v = instance[key]
            texts.append(v if isinstance(v, str) else v[0])
    if texts is None:
        texts = list(instance)
    if len(texts) != 4:
        raise ValueError(f"a quadruplet instance carries 4 texts, got {len(texts)}")
    return list(texts)


def _batches(dataset: Iterable, batch_size: int):
    batch = []
    for instance in dataset:
        batch.append(instance)
        if len(batch) == batch_size:
            yield batch
            batch = []
    if batch:
        yield batch


class QuadrupletLossEvaluator:
    """Same constructor and call protocol as ``models/evaluators.py:35-128``."""

    def __init__(self, quadruplet_dataset, quadruplet_loss, batch_size: int = 32,
                 additional_model_kwargs: Optional[List[str]] = None,
                 additional_loss_kwargs: Optional[List[str]] = None, use_amp: bool = False):
        if batch_size < 1:
            raise ValueError(f"batch_size must be >= 1, {batch_size} given.")
        self._quadruplet_dataset = quadruplet_dataset
        self._quadruplet_loss = quadruplet_loss
        self._batch_size = batch_size
        self._additional_model_kwargs = additional_model_kwargs
        self._additional_loss_kwargs = additional_loss_kwargs
        self._use_amp = use_amp
        self.last_batch_losses: Optional[np.ndarray] = None

    def _encode(self, model, sentences: List[str], extra: dict) -> torch.Tensor:
        emb = model.encode(sentences, batch_size=self._batch_size, show_progress_bar=False,
                           convert_to_tensor=True, **extra)
        if not isinstance(emb, torch.Tensor):
            emb = torch.as_tensor(emb)
        if not emb.is_cuda:
            if not torch.cuda.is_available():
                raise _lib.QstError("QuadrupletLossEvaluator needs a CUDA device (no CPU fallback)")
            emb = emb.cuda()
        return emb

    def _kwargs_for(self, names: Optional[List[str]], batch) -> dict:
        if not names:
            return {}
        out = {}
        for name in names:
            vals = [inst[name] for inst in batch]
            out[name] = vals[0] if all(v == vals[0] for v in vals) else vals
        return out

    def batch_losses(self, model) -> torch.Tensor:
        """One loss value per batch, float32 on the device (no host synchronisation)."""
        losses = []
        with torch.no_grad():
            for batch in _batches(self._quadruplet_dataset, self._batch_size):
                columns = list(zip(*[_texts_of(inst) for inst in batch]))
                model_kw = self._kwargs_for(self._additional_model_kwargs, batch)
                loss_kw = self._kwargs_for(self._additional_loss_kwargs, batch)
                anchor, pos, part, neg = (self._encode(model, list(col), model_kw) for col in columns)
                if self._use_amp:
                    with torch.autocast("cuda"):
                        value = self._quadruplet_loss(x_anchor=anchor, x_pos=pos, x_part=part, x_neg=neg, **loss_kw)
                else:
                    value = self._quadruplet_loss(x_anchor=anchor, x_pos=pos, x_part=part, x_neg=neg, **loss_kw)
                if value.dim() != 0:
                    raise ValueError("QuadrupletLossEvaluator needs a scalar loss (reduction 'mean' or 'sum')")
                losses.append(value.float())
        if not losses:
            return torch.empty(0, dtype=torch.float32)
        return torch.stack(losses)

    def __call__(self, model, output_path: str = None, epoch: int = -1, steps: int = -1) -> float:
        per_batch = self.batch_losses(model).cpu().numpy()          # the only device->host read
        self.last_batch_losses = per_batch
        average_loss = float(incremental_mean_f32(per_batch))
        if output_path is not None:
            full_out_path = os.path.join(output_path, LOG_FILE)
            log = {}
            if os.path.exists(full_out_path):
                with open(full_out_path, "r") as fp:
                    log = json.load(fp)
            for key, value in (("epoch", epoch), ("steps", steps), ("average_loss", average_loss)):
                log.setdefault(key, []).append(value)
            with open(full_out_path, "w") as fp:
                json.dump(log, fp, indent=2)
        return average_loss


def dissimilar_mask(reference_embedding: torch.Tensor, candidate_embeddings: torch.Tensor,
                    threshold: float = 0.2):
    """The similarity filter of the reference's negative mining
    (``/root/reference/dataset/quadruplet_dataset.py:229-234`` with ``compute_cosine_scores``,
    ``dataset/positive_examples_selection.py:50-56``): cosine score of one reference embedding
    against the candidates, and the mask ``score <= NEG_EXAMPLE_SIM_TRESHOLD`` (0.2, ``:20``) of the
    candidates that may serve as negatives.  Returns ``(mask, scores)`` on the device."""
    from .scoring import cos_sim
    scores = cos_sim(reference_embedding, candidate_embeddings)[0]
    return scores <= threshold, scores
