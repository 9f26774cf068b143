"""IR metrics on the device (K4) with the reference's float64 operation order.

Per-query Accuracy / Precision / Recall / reciprocal rank / NDCG / average precision are computed
by ``qst_ir_metrics``; the cross-query reductions are done here with exactly the reductions
sentence-transformers 2.2.2 ``InformationRetrievalEvaluator.compute_metrics`` uses
(``numpy.mean`` for Precision/Recall/NDCG/MAP, a sequential ``+=`` then ``/ len(queries)`` for
Accuracy and MRR), so the final numbers are bit-identical to the reference's given identical
rankings.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np
import torch

from . import _lib

ACC, PREC, REC, RR, NDCG, AP = range(6)


def relevance_csr(relevant_positions: Sequence[Sequence[int]], device) -> tuple:
    """CSR (rowptr int64 [Q+1], cols int64 sorted per row) of relevant corpus positions.

    Relevant ids that are not in the corpus must still be listed (with any position >= N) so the
    row length equals ``len(relevant_docs[qid])``, which the reference uses as denominator.
    """
    lens = np.fromiter((len(r) for r in relevant_positions), dtype=np.int64, count=len(relevant_positions))
    rowptr = np.zeros(len(relevant_positions) + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    cols = np.empty(int(rowptr[-1]), dtype=np.int64)
    for i, r in enumerate(relevant_positions):
        cols[rowptr[i]:rowptr[i + 1]] = np.sort(np.asarray(list(r), dtype=np.int64))
    return torch.from_numpy(rowptr).to(device), torch.from_numpy(cols).to(device)


def _tables(K: int):
    # scalar calls, exactly as the reference evaluates np.log2(i + 2) (an array call may take a
    # SIMD code path that rounds differently)
    log2_tab = np.array([np.log2(i + 2) for i in range(K)], dtype=np.float64)
    idcg = np.zeros(K + 1, dtype=np.float64)
    acc = 0.0
    for i in range(K):                                             # sequential, as compute_dcg_at_k
        acc = acc + 1 / log2_tab[i]
        idcg[i + 1] = acc
    return log2_tab, idcg


def per_query_metrics(ranked_idx: torch.Tensor, rel_rowptr: torch.Tensor, rel_cols: torch.Tensor,
                      ks: Sequence[int]) -> torch.Tensor:
    """[6, len(ks), Q] float64 per-query values on the device (``qst_ir_metrics``)."""
    lib = _lib.load()
    _lib.require_cuda(ranked_idx, rel_rowptr, rel_cols)
    ranked_idx = ranked_idx.contiguous()
    Q, K = ranked_idx.shape
    dev = ranked_idx.device
    log2_tab, idcg = _tables(max(K, max(ks)))
    with torch.cuda.device(dev):
        ks_t = torch.tensor(list(ks), dtype=torch.int32, device=dev)
        log2_t = torch.from_numpy(log2_tab).to(dev)
        idcg_t = torch.from_numpy(idcg).to(dev)
        out = torch.empty((6, len(ks), Q), dtype=torch.float64, device=dev)
        _lib.check(lib.qst_ir_metrics(ranked_idx.data_ptr(), Q, K, rel_rowptr.data_ptr(), rel_cols.data_ptr(),
                                      ks_t.data_ptr(), len(ks), log2_t.data_ptr(), idcg_t.data_ptr(),
                                      out.data_ptr(), _lib.stream_ptr(dev)))
    return out


def _sequential_sum(x: np.ndarray) -> float:
    # a plain left-to-right float64 sum (numpy's cumsum is sequential; Python >= 3.12's sum() is
    # compensated and np.sum is pairwise, neither matches the reference's `+=` loop)
    return float(np.cumsum(x, dtype=np.float64)[-1]) if x.size else 0.0


def reduce_like_reference(per_query: np.ndarray, ks: Sequence[int], accuracy_at_k: List[int],
                          precision_recall_at_k: List[int], mrr_at_k: List[int], ndcg_at_k: List[int],
                          map_at_k: List[int]) -> Dict[str, Dict[int, float]]:
    """Cross-query reductions of compute_metrics, from the [6, n_ks, Q] per-query array."""
    pos = {k: i for i, k in enumerate(ks)}
    n_q = per_query.shape[2]
    return {
        'accuracy@k': {k: int(per_query[ACC, pos[k]].sum()) / n_q for k in accuracy_at_k},
        'precision@k': {k: np.mean(per_query[PREC, pos[k]]) for k in precision_recall_at_k},
        'recall@k': {k: np.mean(per_query[REC, pos[k]]) for k in precision_recall_at_k},
        'ndcg@k': {k: np.mean(per_query[NDCG, pos[k]]) for k in ndcg_at_k},
        'mrr@k': {k: _sequential_sum(per_query[RR, pos[k]]) / n_q for k in mrr_at_k},
        'map@k': {k: np.mean(per_query[AP, pos[k]]) for k in map_at_k},
    }
