"""CPU restatement of the IR scoring / top-k / metric path (TEST INFRASTRUCTURE).

PARITY UNPINNED: the algorithm restated here is that of the third-party package
``sentence-transformers==2.2.2`` (pinned at ``/root/reference/requirements.txt:8``
and ``requirements_cuda.txt:8``), which is absent from ``/root/reference`` and
from this image.  The reference only *calls* it:

* ``util.cos_sim`` / ``util.dot_score``  -> ``ir_evauation_script.py:10,70``,
  ``models/evaluators.py:12,545-546``
* ``InformationRetrievalEvaluator(...)`` -> ``ir_evauation_script.py:107-123``,
  ``models/evaluators.py:572-588``; called at ``ir_evauation_script.py:130-131``
* ``euclidean_score`` is in-repo: ``models/evaluators.py:392-405``

Published 2.2.2 algorithm, restated:

score      a, b -> tensors, 1-D -> [1, D];  cos: F.normalize(p=2, dim=1, eps=1e-12)
           on both, then torch.mm(a_n, b_n.T);  dot: torch.mm(a, b.T)
top-k      per corpus chunk (corpus_chunk_size rows): torch.topk(scores,
           min(max_k, chunk_len), dim=1, largest=True, sorted=False); values and
           indices leave the device as Python lists; one {'corpus_id','score'}
           dict per hit is appended to the query's list, chunk after chunk
ranking    per query: sorted(list, key=score, reverse=True)  (stable)
metrics    Accuracy/Precision/Recall/MRR/NDCG/MAP @k in Python float64,
           MRR and Accuracy accumulated sequentially and divided by the number
           of queries, the others averaged with numpy.mean
"""
from __future__ import annotations

import logging
import os
from typing import Callable, Dict, List, Optional, Set

import numpy as np
import torch

logger = logging.getLogger(__name__)


# --------------------------------------------------------------------------- score functions
def _as_2d(x) -> torch.Tensor:
    if not isinstance(x, torch.Tensor):
        x = torch.tensor(x)
    if x.dim() == 1:
        x = x.unsqueeze(0)
    return x


def cos_sim(a, b) -> torch.Tensor:
    """[UPSTREAM ST 2.2.2 util.cos_sim]; used at ir_evauation_script.py:70."""
    a, b = _as_2d(a), _as_2d(b)
    a_n = torch.nn.functional.normalize(a, p=2, dim=1)
    b_n = torch.nn.functional.normalize(b, p=2, dim=1)
    return torch.mm(a_n, b_n.transpose(0, 1))


def dot_score(a, b) -> torch.Tensor:
    """[UPSTREAM ST 2.2.2 util.dot_score]; used at ir_evauation_script.py:70."""
    a, b = _as_2d(a), _as_2d(b)
    return torch.mm(a, b.transpose(0, 1))


def euclidean_score(a, b) -> torch.Tensor:
    """Restates /root/reference/models/evaluators.py:392-405."""
    a, b = _as_2d(a), _as_2d(b)
    return 1 / (1 + torch.cdist(a, b, p=2))


# --------------------------------------------------------------------------- evaluator
class InformationRetrievalEvaluatorOracle:
    """Restatement of [UPSTREAM ST 2.2.2] ``InformationRetrievalEvaluator``.

    Constructor keywords are the ones the reference passes at
    ``ir_evauation_script.py:107-123`` and ``models/evaluators.py:572-588``.
    """

    def __init__(self,
                 queries: Dict[str, str],
                 corpus: Dict[str, str],
                 relevant_docs: Dict[str, Set[str]],
                 corpus_chunk_size: int = 50000,
                 mrr_at_k: List[int] = [10],
                 ndcg_at_k: List[int] = [10],
                 accuracy_at_k: List[int] = [1, 3, 5, 10],
                 precision_recall_at_k: List[int] = [1, 3, 5, 10],
                 map_at_k: List[int] = [100],
                 show_progress_bar: bool = False,
                 batch_size: int = 32,
                 name: str = '',
                 write_csv: bool = True,
                 score_functions: Dict[str, Callable] = None,
                 main_score_function: str = None):
        if score_functions is None:
            score_functions = {'cos_sim': cos_sim, 'dot_score': dot_score}
        # queries without any relevant document are dropped
        self.queries_ids = [qid for qid in queries
                            if qid in relevant_docs and len(relevant_docs[qid]) > 0]
        self.queries = [queries[qid] for qid in self.queries_ids]
        self.corpus_ids = list(corpus.keys())
        self.corpus = [corpus[cid] for cid in self.corpus_ids]
        self.relevant_docs = relevant_docs
        self.corpus_chunk_size = corpus_chunk_size
        self.mrr_at_k = mrr_at_k
        self.ndcg_at_k = ndcg_at_k
        self.accuracy_at_k = accuracy_at_k
        self.precision_recall_at_k = precision_recall_at_k
        self.map_at_k = map_at_k
        self.show_progress_bar = show_progress_bar
        self.batch_size = batch_size
        self.name = name
        self.write_csv = write_csv
        self.score_functions = score_functions
        self.score_function_names = sorted(score_functions.keys())
        self.main_score_function = main_score_function

        self.csv_file = "Information-Retrieval_evaluation" + ("_" + name if name else "") + "_results.csv"
        self.csv_headers = ["epoch", "steps"]
        for fn in self.score_function_names:
            self.csv_headers += ["{}-Accuracy@{}".format(fn, k) for k in accuracy_at_k]
            for k in precision_recall_at_k:
                self.csv_headers += ["{}-Precision@{}".format(fn, k), "{}-Recall@{}".format(fn, k)]
            self.csv_headers += ["{}-MRR@{}".format(fn, k) for k in mrr_at_k]
            self.csv_headers += ["{}-NDCG@{}".format(fn, k) for k in ndcg_at_k]
            self.csv_headers += ["{}-MAP@{}".format(fn, k) for k in map_at_k]

    # -- SentenceEvaluator protocol: models/evaluators.py:49, ir_evauation_script.py:130-131
    def __call__(self, model, output_path: Optional[str] = None, epoch: int = -1, steps: int = -1,
                 *args, **kwargs) -> float:
        scores = self.compute_metrices(model, *args, **kwargs)
        if output_path is not None and self.write_csv:
            csv_path = os.path.join(output_path, self.csv_file)
            new_file = not os.path.isfile(csv_path)
            with open(csv_path, mode="w" if new_file else "a", encoding="utf-8") as f:
                if new_file:
                    f.write(",".join(self.csv_headers) + "\n")
                row = [epoch, steps]
                for fn in self.score_function_names:
                    row += [scores[fn]['accuracy@k'][k] for k in self.accuracy_at_k]
                    for k in self.precision_recall_at_k:
                        row += [scores[fn]['precision@k'][k], scores[fn]['recall@k'][k]]
                    row += [scores[fn]['mrr@k'][k] for k in self.mrr_at_k]
                    row += [scores[fn]['ndcg@k'][k] for k in self.ndcg_at_k]
                    row += [scores[fn]['map@k'][k] for k in self.map_at_k]
                f.write(",".join(map(str, row)) + "\n")
        if self.main_score_function is None:
            return max(scores[fn]['map@k'][max(self.map_at_k)] for fn in self.score_function_names)
        return scores[self.main_score_function]['map@k'][max(self.map_at_k)]

    @property
    def max_k(self) -> int:
        return max(max(self.mrr_at_k), max(self.ndcg_at_k), max(self.accuracy_at_k),
                   max(self.precision_recall_at_k), max(self.map_at_k))

    def collect_hits(self, model, corpus_model=None, corpus_embeddings=None):
        """Hot loops 1+2 of SURVEY.md section 3.1: chunked score -> topk -> Python lists."""
        if corpus_model is None:
            corpus_model = model
        max_k = self.max_k
        query_embeddings = model.encode(self.queries, show_progress_bar=self.show_progress_bar,
                                        batch_size=self.batch_size, convert_to_tensor=True)
        hits = {fn: [[] for _ in range(len(query_embeddings))] for fn in self.score_functions}
        for start in range(0, len(self.corpus), self.corpus_chunk_size):
            end = min(start + self.corpus_chunk_size, len(self.corpus))
            if corpus_embeddings is None:
                sub = corpus_model.encode(self.corpus[start:end], show_progress_bar=False,
                                          batch_size=self.batch_size, convert_to_tensor=True)
            else:
                sub = corpus_embeddings[start:end]
            for fn, score_function in self.score_functions.items():
                pair_scores = score_function(query_embeddings, sub)
                vals, idx = torch.topk(pair_scores, min(max_k, len(pair_scores[0])), dim=1,
                                       largest=True, sorted=False)
                vals = vals.cpu().tolist()
                idx = idx.cpu().tolist()
                for q in range(len(query_embeddings)):
                    dst = hits[fn][q]
                    for sub_id, score in zip(idx[q], vals[q]):
                        dst.append({'corpus_id': self.corpus_ids[start + sub_id], 'score': score})
        return hits

    def compute_metrices(self, model, corpus_model=None, corpus_embeddings=None) -> Dict[str, dict]:
        hits = self.collect_hits(model, corpus_model, corpus_embeddings)
        return {fn: self.compute_metrics(hits[fn]) for fn in self.score_functions}

    def compute_metrics(self, queries_result_list: List[list]) -> Dict[str, Dict[int, float]]:
        """Hot loop 3: per-query stable sort then float64 metric loops."""
        hits_at_k = {k: 0 for k in self.accuracy_at_k}
        precisions = {k: [] for k in self.precision_recall_at_k}
        recalls = {k: [] for k in self.precision_recall_at_k}
        mrr = {k: 0 for k in self.mrr_at_k}
        ndcg = {k: [] for k in self.ndcg_at_k}
        avep = {k: [] for k in self.map_at_k}

        for q, result in enumerate(queries_result_list):
            relevant = self.relevant_docs[self.queries_ids[q]]
            top_hits = sorted(result, key=lambda h: h['score'], reverse=True)
            is_rel = [h['corpus_id'] in relevant for h in top_hits]

            for k in self.accuracy_at_k:
                if any(is_rel[0:k]):
                    hits_at_k[k] += 1

            for k in self.precision_recall_at_k:
                num_correct = sum(1 for r in is_rel[0:k] if r)
                precisions[k].append(num_correct / k)
                recalls[k].append(num_correct / len(relevant))

            for k in self.mrr_at_k:
                for rank, r in enumerate(is_rel[0:k]):
                    if r:
                        mrr[k] += 1.0 / (rank + 1)
                        break

            for k in self.ndcg_at_k:
                predicted = [1 if r else 0 for r in is_rel[0:k]]
                ideal = [1] * len(relevant)
                ndcg[k].append(self.compute_dcg_at_k(predicted, k) / self.compute_dcg_at_k(ideal, k))

            for k in self.map_at_k:
                num_correct = 0
                sum_precisions = 0
                for rank, r in enumerate(is_rel[0:k]):
                    if r:
                        num_correct += 1
                        sum_precisions += num_correct / (rank + 1)
                avep[k].append(sum_precisions / min(k, len(relevant)))

        n_q = len(self.queries)
        return {
            'accuracy@k': {k: hits_at_k[k] / n_q for k in hits_at_k},
            'precision@k': {k: np.mean(precisions[k]) for k in precisions},
            'recall@k': {k: np.mean(recalls[k]) for k in recalls},
            'ndcg@k': {k: np.mean(ndcg[k]) for k in ndcg},
            'mrr@k': {k: mrr[k] / n_q for k in mrr},
            'map@k': {k: np.mean(avep[k]) for k in avep},
        }

    @staticmethod
    def compute_dcg_at_k(relevances, k):
        dcg = 0
        for i in range(min(len(relevances), k)):
            dcg += relevances[i] / np.log2(i + 2)
        return dcg


# --------------------------------------------------------------------------- helpers for tests / bench
class PrecomputedEmbeddingModel:
    """The 'fake backend' of SURVEY.md section 4: ``encode`` returns rows of a fixed
    embedding table; sentences are stringified row ids."""

    def __init__(self, table: torch.Tensor):
        self.table = table

    def encode(self, sentences, batch_size: int = 32, show_progress_bar: bool = False,
               convert_to_tensor: bool = True, **_):
        rows = torch.tensor([int(s) for s in sentences], dtype=torch.long)
        return self.table[rows]


def ranked_ids(hits_for_fn: List[list], k: int) -> List[List[str]]:
    """Top-k corpus ids per query, after the same stable descending sort the metrics use."""
    out = []
    for result in hits_for_fn:
        top = sorted(result, key=lambda h: h['score'], reverse=True)[:k]
        out.append([h['corpus_id'] for h in top])
    return out


def topk_dense(queries: torch.Tensor, corpus: torch.Tensor, k: int, score: str = "cos_sim",
               corpus_chunk_size: int = 50000):
    """score -> per-chunk topk -> global descending order, as tensors (values, indices).

    Same arithmetic as ``collect_hits`` + the sort of ``compute_metrics`` but without Python
    dict lists, for sizes where those lists are the bottleneck.  Ties keep chunk order
    (stable sort), as the list version does.
    """
    fn = {"cos_sim": cos_sim, "dot_score": dot_score, "euclid_score": euclidean_score}[score]
    vals_all, idx_all = [], []
    for start in range(0, corpus.shape[0], corpus_chunk_size):
        sub = corpus[start:start + corpus_chunk_size]
        s = fn(queries, sub)
        v, i = torch.topk(s, min(k, s.shape[1]), dim=1, largest=True, sorted=False)
        vals_all.append(v)
        idx_all.append(i + start)
    v = torch.cat(vals_all, dim=1)
    i = torch.cat(idx_all, dim=1)
    order = torch.sort(v, dim=1, descending=True, stable=True).indices[:, :k]
    return torch.gather(v, 1, order), torch.gather(i, 1, order)
