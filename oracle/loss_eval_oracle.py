"""CPU oracle for the QuadrupletLossEvaluator running mean.  TEST INFRASTRUCTURE ONLY.

Restates ``/root/reference/models/evaluators.py:84-98``: per-batch loss values (from
``oracle.loss_oracle``) folded with ``average_loss = average_loss + 1 / (i + 1) * (loss_value -
average_loss)`` in torch tensor arithmetic, starting from the Python float 0.0.  The expression is
the reference's own; the loss values under it are pinned by ``tests/golden/loss_golden.npz``.
PINNED as a whole: ``evaluate`` equals, bit for bit, what the reference's un-scripted stack (its
``QuadrupletLossEvaluator``, its ``QuadrupletSentenceTransformerLossModel``, its ``GammaQuadrupletLoss``) returns
for a table-lookup sentence model (``tests/test_reference_evaluators.py``, fixture ``full_loss_stack`` of
``tests/golden/evaluators_golden.json``).
"""
import torch

from . import loss_oracle


def running_average(batch_losses):
    """`batch_losses`: iterable of 0-dim float32 tensors (CPU)."""
    average_loss = 0.0
    for i, loss_value in enumerate(batch_losses):
        average_loss = average_loss + 1 / (i + 1) * (loss_value - average_loss)
    return average_loss


def evaluate(anchor, pos, part, neg, batch_size, **loss_kwargs):
    """Average loss over consecutive batches of the four [n, D] embedding tables."""
    losses = []
    for s in range(0, anchor.shape[0], batch_size):
        e = s + batch_size
        losses.append(loss_oracle.gamma_quadruplet_loss(anchor[s:e], pos[s:e], part[s:e], neg[s:e], **loss_kwargs))
    return running_average(losses), torch.stack(losses)
