"""CPU restatement of the QuadrupletEvaluator arithmetic (TEST INFRASTRUCTURE).

Follows ``/root/reference/models/evaluators.py:130-389`` (three TripletEvaluators + the gamma-weighted
global accuracy at ``:367``) with the inner sentence-transformers 2.2.2 ``TripletEvaluator.__call__``
restated: numpy embeddings -> sklearn ``paired_cosine_distances`` / ``paired_manhattan_distances`` /
``paired_euclidean_distances`` -> count ``pos_distance < neg_distance`` -> accuracy; the evaluator
returns the accuracy of ``main_distance_function`` or the max of the three.
PARITY UNPINNED for the TripletEvaluator part (third-party, absent from /root/reference, no golden
vectors); the combination formula is the reference's own.
"""
from __future__ import annotations

import numpy as np
from sklearn.metrics.pairwise import paired_cosine_distances, paired_euclidean_distances, paired_manhattan_distances


def triplet_accuracies(anchors: np.ndarray, positives: np.ndarray, negatives: np.ndarray):
    """(accuracy_cos, accuracy_manhattan, accuracy_euclidean) of one TripletEvaluator."""
    out = []
    for fn in (paired_cosine_distances, paired_manhattan_distances, paired_euclidean_distances):
        pos, neg = fn(anchors, positives), fn(anchors, negatives)
        correct = 0
        for i in range(len(pos)):
            if pos[i] < neg[i]:
                correct += 1
        out.append(correct / len(pos))
    return tuple(out)


def quadruplet_accuracy(anchor, pos, part, neg, gamma: float = 0.6, main: str = None):
    """Global accuracy and the three triplet accuracies; ``main`` in {None, 'cos', 'manhattan', 'euclid'}."""
    pick = {None: max, "cos": lambda a, m, e: a, "manhattan": lambda a, m, e: m, "euclid": lambda a, m, e: e}[main]
    a, p, pp, n = (np.asarray(x, dtype=np.float32) for x in (anchor, pos, part, neg))
    pos_part = pick(*triplet_accuracies(a, p, pp))
    pos_neg = pick(*triplet_accuracies(a, p, n))
    part_neg = pick(*triplet_accuracies(a, pp, n))
    glob = (((1 - gamma) * pos_part + gamma * part_neg) + pos_neg) / 2
    return glob, dict(pos_part=pos_part, pos_neg=pos_neg, part_neg=part_neg)


def paired_distances(anchor, other):
    a, o = np.asarray(anchor, dtype=np.float32), np.asarray(other, dtype=np.float32)
    return paired_cosine_distances(a, o), paired_manhattan_distances(a, o), paired_euclidean_distances(a, o)
