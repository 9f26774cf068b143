"""Synthetic inputs of the BASELINE.json configurations.

Every tensor comes from its own ``torch.Generator().manual_seed(14 + i)``; 14 is the reference's
``RANDOM_SEED`` (``/root/reference/dataset/constants.py:5``).  Generation happens on the CPU so
the same bits feed the CPU oracle and the GPU path.
"""
from __future__ import annotations

from typing import Dict, List, Set, Tuple

import torch

SEED = 14


# argparse defaults of the reference's evaluation script (``/root/reference/ir_evauation_script.py:163-173``):
# MRR / NDCG at ten cut-offs, Accuracy / Precision-Recall / MAP at twelve, the largest 900.  Held against the
# script's source by tests/test_host_logic.py::test_script_default_k_lists_are_the_reference_scripts.
_K10 = [5, 10, 20, 30, 40, 50, 100, 200, 500, 900]
_K12 = [1, 3] + _K10
SCRIPT_DEFAULT_K_LISTS = {"mrr_at_k": _K10, "ndcg_at_k": _K10, "accuracy_at_k": _K12,
                          "precision_recall_at_k": _K12, "map_at_k": _K12}


def _gen(i: int) -> torch.Generator:
    return torch.Generator().manual_seed(SEED + i)


def gaussian_embeddings(n: int, d: int, stream: int, dtype=torch.float32) -> torch.Tensor:
    return torch.randn(n, d, generator=_gen(stream), dtype=torch.float32).to(dtype)


def clustered_embeddings(n: int, d: int, stream: int, n_centroids: int = 64, noise: float = 0.3) -> torch.Tensor:
    """corpus = centroid + noise * N(0,1): dense near-ties, stresses the exactness certificate."""
    g = _gen(stream)
    cent = torch.randn(n_centroids, d, generator=g)
    which = torch.randint(0, n_centroids, (n,), generator=g)
    return cent[which] + noise * torch.randn(n, d, generator=g)


def ir_eval_set(n_queries: int = 1000, n_corpus: int = 10000, d: int = 384, n_pos: int = 4, n_part: int = 4,
                use_part_pos: bool = True) -> Tuple[torch.Tensor, torch.Tensor, Dict[str, str], Dict[str, str],
                                                    Dict[str, Set[str]]]:
    """Config 1: queries with planted positives (q + 0.3 e) and partial positives (q + 0.8 e).

    Mirrors n_pos=4 / n_part_pos=4 of ``ir_evauation_script.py:36-43``; relevant = positives
    (+ partial positives when ``use_part_pos``, the script's default flags ``:102-105``).
    Returns (query_emb, corpus_emb, queries, corpus, relevant_docs); the dict values are
    stringified row ids of the table ``torch.cat([query_emb, corpus_emb])`` that ``TableModel``
    serves (queries first).
    """
    q = gaussian_embeddings(n_queries, d, 0)
    c = gaussian_embeddings(n_corpus, d, 1)
    g = _gen(2)
    planted = torch.randperm(n_corpus, generator=g)[:n_queries * (n_pos + n_part)].view(n_queries, n_pos + n_part)
    e = torch.randn(n_queries, n_pos + n_part, d, generator=g)
    scale = torch.cat([torch.full((n_pos,), 0.3), torch.full((n_part,), 0.8)]).view(1, -1, 1)
    c[planted.reshape(-1)] = (q.unsqueeze(1) + scale * e).reshape(-1, d)
    queries = {f"q{i}": str(i) for i in range(n_queries)}
    corpus = {f"d{i}": str(n_queries + i) for i in range(n_corpus)}
    n_rel = n_pos + (n_part if use_part_pos else 0)
    relevant = {f"q{i}": {f"d{int(j)}" for j in planted[i, :n_rel]} for i in range(n_queries)}
    return q, c, queries, corpus, relevant


def quadruplet_batch(b: int = 4096, d: int = 768) -> List[torch.Tensor]:
    """Config 2: anchor / positive / partial positive / negative, N(0,1)."""
    return [gaussian_embeddings(b, d, 10 + i) for i in range(4)]


class TableModel:
    """Stand-in for a SentenceTransformer: ``encode`` looks rows up in a fixed embedding table
    (sentences are stringified row ids).  The 'fake backend' of SURVEY.md section 4."""

    def __init__(self, table: torch.Tensor):
        self.table = table

    def encode(self, sentences, batch_size: int = 32, show_progress_bar: bool = False,
               convert_to_tensor: bool = True, **_):
        rows = torch.tensor([int(s) for s in sentences], dtype=torch.long, device=self.table.device)
        return self.table[rows]
