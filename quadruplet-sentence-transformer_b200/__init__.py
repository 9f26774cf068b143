"""B200-native retrieval scoring + quadruplet loss (hot path of
lucastrefezza/quadruplet-sentence-transformer), drop-in behind the reference's Python protocols.

The directory name carries a hyphen (it mirrors the reference's repository name), so import it
through the alias module at the repository root::

    import qst_b200
    from qst_b200 import GammaQuadrupletLoss, InformationRetrievalEvaluator, cos_sim

Public surface (same names and call signatures as the reference side uses):

* ``GammaQuadrupletLoss``, ``QuadrupletLoss``, ``gamma_quadruplet_loss``
  (``/root/reference/models/losses/losses.py``)
* ``InformationRetrievalEvaluator``, ``cos_sim``, ``dot_score``
  (sentence-transformers 2.2.2, as constructed at ``ir_evauation_script.py:107-123``) and the
  reference's own ``euclidean_score`` (``models/evaluators.py:392-405``)
* ``QuadrupletEvaluator`` (``models/evaluators.py:130-389``) with its paired distances on the device
* ``QuadrupletLossEvaluator`` (``models/evaluators.py:34-128``): running mean of the fused loss
* ``CorpusIndex``, ``topk``: the scoring engine underneath
* ``ShardedCorpus``: corpus-sharded retrieval over NCCL

All arithmetic runs in ``libqst.so`` (hand-written CUDA for sm_100a).  There is no CPU fallback.
"""
from . import _lib
from ._lib import QstError, QstLibraryError
from .quad_loss import (GammaQuadrupletLoss, QuadrupletLoss, gamma_quadruplet_loss,
                        gamma_quadruplet_loss_and_grads)
from .scoring import (CorpusIndex, HostTopkPipeline, TopkResult, cos_sim, dot_score, euclidean_score, prepare_rows,
                      topk, topk_host)
from .ir_evaluator import InformationRetrievalEvaluator, load_ir_evaluation_set
from .quad_evaluator import QuadrupletEvaluator, SimilarityFunction, paired_distance_counts
from .loss_evaluator import QuadrupletLossEvaluator, dissimilar_mask, incremental_mean_f32
from . import comm, metrics, synth
from .sharded import ShardedCorpus

__all__ = [
    "GammaQuadrupletLoss", "QuadrupletLoss", "gamma_quadruplet_loss", "gamma_quadruplet_loss_and_grads",
    "InformationRetrievalEvaluator", "cos_sim", "dot_score", "euclidean_score", "CorpusIndex", "TopkResult", "topk",
    "topk_host", "HostTopkPipeline", "prepare_rows", "QuadrupletEvaluator", "SimilarityFunction", "paired_distance_counts",
    "QuadrupletLossEvaluator", "dissimilar_mask", "incremental_mean_f32",
    "load_ir_evaluation_set", "ShardedCorpus", "comm", "metrics", "synth", "QstError", "QstLibraryError",
]
