"""CPU restatement of the gamma-quadruplet loss (TEST INFRASTRUCTURE).

PINNED against the reference's own module through ``tests/golden/loss_golden.npz``
(made by ``tests/golden/make_loss_golden.py`` importing
``/root/reference/models/losses/losses.py`` by file path) -- see
``tests/test_oracle_loss.py``.

Follows ``/root/reference/models/losses/losses.py``:

* validation ............................ ``:20-32``
* three triplet terms a, b, c ........... ``:35-61``
* reductions none / sum / mean .......... ``:64-69``

with ``F.triplet_margin_loss`` expanded into its published definition
([UPSTREAM torch] ``nn/functional.py`` ``triplet_margin_loss`` ->
``ATen/native/Loss.cpp``): ``clamp_min(margin + d(a,p) - d(a,n), 0)`` where
``d(u,v) = || u - v + eps ||_p`` (eps = 1e-6 is added to the *difference*) and
``swap=True`` replaces ``d(a,n)`` by ``min(d(a,n), d(p,n))``.

Gradients come from torch autograd over these plain ops, which is the same
graph the reference builds.
"""
from __future__ import annotations

import importlib.util
import math
import os

import torch

REDUCTIONS = ("mean", "sum", "none")
EPS = 1e-6
REFERENCE_LOSSES_PATH = "/root/reference/models/losses/losses.py"


def _validate(gamma, margin_pos_neg, margin_pos_part, margin_part_neg, p, reduction):
    if gamma < 0 or gamma > 1:
        raise ValueError(f"gamma must be between 0 and 1, {gamma} given")
    if margin_pos_neg <= 0:
        raise ValueError(f"margin_pos_neg must be positive, {margin_pos_neg} given")
    if margin_pos_part <= 0:
        raise ValueError(f"margin_pos_part must be positive, {margin_pos_part} given")
    if margin_part_neg <= 0:
        raise ValueError(f"margin_part_neg must be positive, {margin_part_neg} given")
    if reduction not in REDUCTIONS:
        raise ValueError(f"reduction must be one of: {REDUCTIONS}, {reduction} given")
    if p <= 0:
        raise ValueError(f"p must be positive, {p} given")


def pairwise_distance(u: torch.Tensor, v: torch.Tensor, p: float, eps: float = EPS) -> torch.Tensor:
    return torch.linalg.vector_norm(u - v + eps, ord=p, dim=-1)


def triplet_term(anchor, positive, negative, margin, p, swap):
    d_pos = pairwise_distance(anchor, positive, p)
    d_neg = pairwise_distance(anchor, negative, p)
    if swap:
        d_neg = torch.minimum(d_neg, pairwise_distance(positive, negative, p))
    return torch.clamp_min(margin + d_pos - d_neg, 0)


def gamma_quadruplet_loss(x_anchor, x_pos, x_part, x_neg, gamma=0.6, margin_pos_neg=1.0,
                          margin_pos_part=0.5, margin_part_neg=0.5, p=2.0, swap=False,
                          reduction="mean"):
    _validate(gamma, margin_pos_neg, margin_pos_part, margin_part_neg, p, reduction)
    a = triplet_term(x_anchor, x_pos, x_neg, margin_pos_neg, p, swap)
    b = triplet_term(x_anchor, x_part, x_neg, margin_part_neg, p, swap)
    c = triplet_term(x_anchor, x_pos, x_part, margin_pos_part, p, swap)
    if reduction == "none":
        return a + gamma * b + (1 - gamma) * c
    if reduction == "sum":
        return a.sum() + (gamma * b).sum() + ((1 - gamma) * c).sum()
    return a.mean() + (gamma * b).mean() + ((1 - gamma) * c).mean()


def loss_and_grads(x_anchor, x_pos, x_part, x_neg, **kw):
    """Loss (any reduction) and d(sum of loss)/d(inputs) via autograd, all float32 CPU."""
    xs = [t.detach().clone().float().requires_grad_(True) for t in (x_anchor, x_pos, x_part, x_neg)]
    out = gamma_quadruplet_loss(*xs, **kw)
    out.sum().backward()
    return out.detach(), [t.grad for t in xs]


def load_reference_losses(path: str = REFERENCE_LOSSES_PATH):
    """Import the reference's own ``losses.py`` by file path (authoring container only).

    Never called by ``-m gpu`` tests, ``smoke()`` or ``bench.py``: ``/root/reference``
    does not exist on the GPU box.
    """
    if not os.path.isfile(path):
        return None
    import sys
    spec = importlib.util.spec_from_file_location("_reference_losses", path)
    mod = importlib.util.module_from_spec(spec)
    keep, sys.dont_write_bytecode = sys.dont_write_bytecode, True      # never leave a __pycache__ in /root/reference
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.dont_write_bytecode = keep
    return mod


def is_inf(p: float) -> bool:
    return math.isinf(p)
