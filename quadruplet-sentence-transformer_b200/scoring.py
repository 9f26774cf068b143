"""Query-by-corpus scoring + exact top-k on B200 (K1 prep, K2 tensor-core select, K3 rescore).

Host-side mirror of what sentence-transformers 2.2.2 ``InformationRetrievalEvaluator.
compute_metrices`` does per corpus chunk -- ``score_function(q, c)`` then ``torch.topk`` -- for the
score functions the reference wires in at ``/root/reference/ir_evauation_script.py:70``
(``cos_sim``, ``dot_score``).  The corpus stays resident in HBM as a ``CorpusIndex`` (fp32 master,
normalised bf16 operand, inverse norms); ``topk`` runs the three kernels on the current stream.

Nothing here falls back to PyTorch arithmetic: CPU tensors raise, a missing ``libqst.so`` raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib

SCORE_CODES = {"cos_sim": _lib.QST_SCORE_COS, "dot_score": _lib.QST_SCORE_DOT,
               "euclid_score": _lib.QST_SCORE_EUCLID}
# how corpus rows / query rows become tensor-core operands for each score function (qst_prep_mode)
CORPUS_PREP = {"cos_sim": _lib.QST_PREP_COS, "dot_score": _lib.QST_PREP_RAW,
               "euclid_score": _lib.QST_PREP_EUCLID_CORPUS}
QUERY_PREP = {"cos_sim": _lib.QST_PREP_COS, "dot_score": _lib.QST_PREP_RAW,
              "euclid_score": _lib.QST_PREP_EUCLID_QUERY}


def cos_sim(a, b):
    """Marker + dense entry for the ``cos_sim`` score function (``ir_evauation_script.py:70``).

    The evaluator recognises this callable and runs the fused path (the [Q, N] matrix is never
    materialised).  Called directly it returns the dense fp32 cosine matrix computed from the
    exact rescoring kernel, for the small inputs the reference uses it on elsewhere.
    """
    return _dense_scores(a, b, "cos_sim")


def dot_score(a, b):
    """Marker + dense entry for ``dot_score`` (``ir_evauation_script.py:70``)."""
    return _dense_scores(a, b, "dot_score")


def euclidean_score(a, b):
    """Marker + dense entry for the reference's own score function ``1 / (1 + cdist(a, b))``
    (``/root/reference/models/evaluators.py:392-405``, wired in at ``ir_evauation_script.py:70``)."""
    return _dense_scores(a, b, "euclid_score")


cos_sim.qst_score = "cos_sim"
dot_score.qst_score = "dot_score"
euclidean_score.qst_score = "euclid_score"


# The three callables the reference's unmodified code hands over (``ir_evauation_script.py:8-16, 70``,
# ``models/evaluators.py:9-12, 545``): sentence-transformers' ``util.cos_sim`` / ``util.dot_score`` and the
# reference's own ``models.evaluators.euclidean_score``.  Their arithmetic is known, so they select the fused
# path like this package's own markers -- the evaluator drops in without touching the score-function table.
_KNOWN_SCORE_FUNCTIONS = {("sentence_transformers.util", "cos_sim"): "cos_sim",
                          ("sentence_transformers.util", "dot_score"): "dot_score",
                          ("models.evaluators", "euclidean_score"): "euclid_score"}


def score_name_of(fn) -> Optional[str]:
    """Name of the fused score function behind a callable, or None for a foreign callable."""
    name = getattr(fn, "qst_score", None)
    if name is not None:
        return name
    module, fname = getattr(fn, "__module__", None), getattr(fn, "__name__", None)
    if isinstance(module, str) and module.startswith("sentence_transformers.util"):
        module = "sentence_transformers.util"          # later releases moved the functions into a sub-module
    return _KNOWN_SCORE_FUNCTIONS.get((module, fname))


def _as_2d_cuda(x) -> torch.Tensor:
    if not isinstance(x, torch.Tensor):
        x = torch.tensor(x)
    if x.dim() == 1:
        x = x.unsqueeze(0)
    _lib.require_cuda(x)
    return x


@dataclass
class PreparedRows:
    f32: torch.Tensor        # [n, D] fp32 master (contiguous)
    bf16: torch.Tensor       # [n, D_pad] bf16 tensor-core operand
    inv_norm: torch.Tensor   # [n] 1/max(||x||, 1e-12)
    sq_norm: torch.Tensor    # [n] ||x||^2
    err: torch.Tensor        # [n] ||bf16(row) - row||
    stats: torch.Tensor      # [2] max err, max used norm

    @property
    def n(self) -> int:
        return self.f32.shape[0]

    @property
    def d(self) -> int:
        return self.f32.shape[1]


def prepare_rows(x: torch.Tensor, mode, want_bf16: bool = True) -> PreparedRows:
    """K1 (``qst_prep_rows``): norms, bf16 operand for the given ``qst_prep_mode`` (``True``/``False``
    are accepted as cos / raw), zero padding."""
    mode = int(mode)
    lib = _lib.load()
    x = _as_2d_cuda(x)
    if x.dtype not in (torch.float32, torch.float16, torch.bfloat16):
        x = x.float()
    x = x.contiguous()
    n, d = x.shape
    dev = x.device
    d_pad = lib.qst_padded_dim_for(d, mode)
    with torch.cuda.device(dev):
        bf = torch.empty((n, d_pad), dtype=torch.bfloat16, device=dev) if want_bf16 else None
        inv = torch.empty(n, dtype=torch.float32, device=dev)
        sq = torch.empty(n, dtype=torch.float32, device=dev)
        err = torch.empty(n, dtype=torch.float32, device=dev)
        stats = torch.zeros(2, dtype=torch.float32, device=dev)
        _lib.check(lib.qst_prep_rows(x.data_ptr(), _lib.dtype_code(x.dtype), n, d, mode, _lib.ptr(bf),
                                     inv.data_ptr(), sq.data_ptr(), err.data_ptr(), stats.data_ptr(),
                                     _lib.stream_ptr(dev)))
    f32 = x if x.dtype == torch.float32 else x.float()
    return PreparedRows(f32, bf, inv, sq, err, stats)


class CorpusIndex:
    """One corpus shard resident in HBM, prepared once for a given score function.

    ``idx_offset`` is the global id of the shard's first row (corpus-sharded retrieval,
    SURVEY.md section 8e).
    """

    def __init__(self, corpus_embeddings: torch.Tensor, score: str = "cos_sim", idx_offset: int = 0):
        if score not in SCORE_CODES:
            raise ValueError(f"score must be one of {sorted(SCORE_CODES)}, {score} given")
        self.score = score
        self.idx_offset = int(idx_offset)
        self.rows = prepare_rows(corpus_embeddings, CORPUS_PREP[score])

    @property
    def n(self) -> int:
        return self.rows.n

    @property
    def d(self) -> int:
        return self.rows.d

    @property
    def device(self) -> torch.device:
        return self.rows.f32.device


_ws_cache = {}


def _workspace(nbytes: int, device: torch.device, tag: str) -> torch.Tensor:
    # one buffer per (purpose, device, stream, host thread): work queued by different host threads on
    # the same stream interleaves in submission order, so they must not share scratch memory
    key = (tag, device.index, torch.cuda.current_stream(device).cuda_stream, threading.get_ident())
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1), dtype=torch.uint8, device=device)
        _ws_cache[key] = ws
    return ws


def release_workspaces(thread_ident: Optional[int] = None) -> int:
    """Drops the cached scratch buffers of one host thread (default: the calling one) -- or of every
    thread with ``thread_ident=-1``.  The cache only ever grows otherwise (one buffer per purpose, device,
    stream and thread, sized for the largest call seen).  Returns the number of buffers released."""
    me = threading.get_ident() if thread_ident is None else thread_ident
    keys = [k for k in _ws_cache if thread_ident == -1 or k[3] == me]
    for k in keys:
        del _ws_cache[k]
    return len(keys)


def first_pass_kprime(k: int, kprime: int) -> int:
    """k' of K3's first pass: k + 62 % of the default head-room, in steps of 16 (k = 100: 176 of 224;
    measured at config 3: 33 of 10 000 queries take the second pass, K3 1.49 -> 1.28 ms; 160 would send
    15 % there).  0 (one pass) when that leaves nothing to save.  ``QST_K3_FIRST`` overrides (0 = off)."""
    env = os.environ.get("QST_K3_FIRST")
    if env is not None:
        return int(env)
    first = ((k + int(0.62 * (kprime - k)) + 15) // 16) * 16
    return first if k <= first <= kprime - 16 else 0


def make_plan(Q: int, N: int, D: int, k: int, kprime: int = 0, score: str = "cos_sim",
              sm_count: int = 0) -> _lib.TopkPlan:
    plan = _lib.TopkPlan()
    _lib.check(_lib.load().qst_topk_plan_make(Q, N, D, k, kprime, SCORE_CODES[score], sm_count, C.byref(plan)))
    return plan


@dataclass
class TopkResult:
    values: torch.Tensor    # [Q, k] fp32, descending
    indices: torch.Tensor   # [Q, k] int64 global corpus ids, -1 where the shard has fewer than k rows
    margin: torch.Tensor    # [Q] fp32 certificate (> 0: proven exact; +inf after an exact re-scan)
    plan: _lib.TopkPlan


# Query rows per K2 launch.  The candidate workspace grows with the number of query rows (about 100 KB
# per row at 1M corpus rows); larger batches are walked in tiles of this many rows, which costs
# nothing (a tile already fills the machine many times over) and bounds the workspace to ~3 GB.
QUERY_TILE = 32768


def topk(queries: torch.Tensor, index: CorpusIndex, k: int, kprime: int = 0, exact: bool = True,
         prepared_queries: Optional[PreparedRows] = None) -> TopkResult:
    """Exact top-``k`` corpus rows per query under ``index.score``.

    bf16 tensor-core pass keeps ``kprime`` candidates per query, fp32 rescoring orders them; with
    ``exact=True`` queries whose certificate fails are re-scanned in fp32 on the device -- all of them
    (the re-scan queues as many passes of 8192 queries as the batch could need).  ``margin > 0`` marks a
    proven-exact row; the only way a row comes back with ``margin == 0`` is more than 2048 documents
    tied at or above its k-th score.  Nothing is read back here (the call is asynchronous); callers that
    hand rankings to a user check the margins where they synchronise anyway
    (``InformationRetrievalEvaluator.last_uncertified``).
    """
    lib = _lib.load()
    pq = prepared_queries if prepared_queries is not None else prepare_rows(queries, QUERY_PREP[index.score])
    if pq.d != index.d:
        raise ValueError(f"query dim {pq.d} != corpus dim {index.d}")
    dev = index.device
    Q, N, D = pq.n, index.n, index.d
    if Q == 0:
        z = torch.empty((0, k), device=dev)
        return TopkResult(z, z.long(), torch.empty(0, device=dev), None)
    if Q > QUERY_TILE:
        with torch.cuda.device(dev):
            vals = torch.empty((Q, k), dtype=torch.float32, device=dev)
            idx = torch.empty((Q, k), dtype=torch.int64, device=dev)
            margin = torch.empty(Q, dtype=torch.float32, device=dev)
            for s in range(0, Q, QUERY_TILE):
                e = min(Q, s + QUERY_TILE)
                tile = PreparedRows(pq.f32[s:e], pq.bf16[s:e], pq.inv_norm[s:e], pq.sq_norm[s:e], pq.err[s:e],
                                    pq.stats)
                r = topk(None, index, k, kprime, exact, prepared_queries=tile)
                vals[s:e], idx[s:e], margin[s:e] = r.values, r.indices, r.margin
        return TopkResult(vals, idx, margin, r.plan)
    plan = make_plan(Q, N, D, k, kprime, index.score)
    st = _lib.stream_ptr(dev)
    cos = index.score == "cos_sim"
    with torch.cuda.device(dev):
        ws = _workspace(plan.ws_bytes, dev, "select")
        vals = torch.empty((Q, k), dtype=torch.float32, device=dev)
        idx = torch.empty((Q, k), dtype=torch.int64, device=dev)
        margin = torch.empty(Q, dtype=torch.float32, device=dev)
        c = index.rows
        _lib.check(lib.qst_score_select(C.byref(plan), pq.bf16.data_ptr(), c.bf16.data_ptr(), ws.data_ptr(), st))
        # K3 in two passes when k' was left to the default: a narrower first pass certifies all but a
        # fraction of a percent of the queries, the full k' is only spent on the rest
        first = first_pass_kprime(plan.k, plan.kprime) if (kprime <= 0 and exact) else 0
        _lib.check(lib.qst_finalize_topk_adaptive(C.byref(plan), first, ws.data_ptr(), pq.f32.data_ptr(),
                                                  pq.inv_norm.data_ptr() if cos else None, pq.err.data_ptr(),
                                                  c.f32.data_ptr(), c.inv_norm.data_ptr() if cos else None,
                                                  c.stats.data_ptr(), index.idx_offset, vals.data_ptr(),
                                                  idx.data_ptr(), margin.data_ptr(), st))
        if exact:
            scratch = _workspace(lib.qst_exact_rescan_workspace_bytes(Q, k), dev, "rescan")
            _lib.check(lib.qst_exact_rescan(Q, N, D, k, SCORE_CODES[index.score], pq.f32.data_ptr(),
                                            pq.inv_norm.data_ptr() if cos else None, c.f32.data_ptr(),
                                            c.inv_norm.data_ptr() if cos else None, index.idx_offset,
                                            vals.data_ptr(), idx.data_ptr(), margin.data_ptr(),
                                            scratch.data_ptr(), st))
    return TopkResult(vals, idx, margin, plan)


def dense_tensorcore_scores(q_bf16: torch.Tensor, c_bf16: torch.Tensor) -> torch.Tensor:
    """Raw bf16 tensor-core scores [Q, N] (``qst_score_dense``); validation aid for K2."""
    lib = _lib.load()
    _lib.require_cuda(q_bf16, c_bf16)
    Q, d_pad = q_bf16.shape
    N = c_bf16.shape[0]
    out = torch.empty((Q, N), dtype=torch.float32, device=q_bf16.device)
    with torch.cuda.device(q_bf16.device):
        _lib.check(lib.qst_score_dense(q_bf16.data_ptr(), Q, c_bf16.data_ptr(), N, d_pad, out.data_ptr(),
                                       _lib.stream_ptr(q_bf16.device)))
    return out


def _dense_scores(a, b, score: str) -> torch.Tensor:
    """Dense fp32 [Q, N] scores with K3's exact arithmetic (``qst_dense_scores``: fp32 dot products on
    CUDA cores, bit-identical to the scores ``topk`` reports).  This is what the score functions return
    when they are called directly, as the reference does on small inputs
    (``dataset/positive_examples_selection.py:55``, ``dataset/quadruplet_dataset.py:229-234``,
    ``training/main.py:57``); the evaluator never comes here."""
    lib = _lib.load()
    a, b = _as_2d_cuda(a), _as_2d_cuda(b)
    if a.shape[1] != b.shape[1]:
        raise ValueError(f"embedding dims differ: {a.shape[1]} vs {b.shape[1]}")
    dev = a.device
    cos = score == "cos_sim"
    pa = prepare_rows(a, _lib.QST_PREP_COS if cos else _lib.QST_PREP_RAW, want_bf16=False)
    pb = prepare_rows(b.to(dev), _lib.QST_PREP_COS if cos else _lib.QST_PREP_RAW, want_bf16=False)
    Q, N, D = pa.n, pb.n, pa.d
    out = torch.empty((Q, N), dtype=torch.float32, device=dev)
    if Q == 0 or N == 0:
        return out
    with torch.cuda.device(dev):
        for s0 in range(0, Q, 65535):
            s1 = min(Q, s0 + 65535)
            _lib.check(lib.qst_dense_scores(s1 - s0, N, D, SCORE_CODES[score], pa.f32[s0:s1].data_ptr(),
                                            pa.inv_norm[s0:s1].data_ptr() if cos else None, pb.f32.data_ptr(),
                                            pb.inv_norm.data_ptr() if cos else None, out[s0:s1].data_ptr(),
                                            _lib.stream_ptr(dev)))
    return out


_pinned_cache = {}


def _pinned_outputs(shape, device: torch.device):
    """Reusable pinned host buffers for the ranking (allocating pinned memory costs milliseconds), one
    pair per (device, stream, shape)."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream, shape)
    buf = _pinned_cache.get(key)
    if buf is None:
        buf = (torch.empty(shape, dtype=torch.float32, pin_memory=True),
               torch.empty(shape, dtype=torch.int64, pin_memory=True))
        _pinned_cache[key] = buf
    return buf


def topk_host(queries_host: torch.Tensor, index: CorpusIndex, k: int, kprime: int = 0,
              exact: bool = True, out: Optional[Tuple[torch.Tensor, torch.Tensor]] = None
              ) -> Tuple[torch.Tensor, torch.Tensor]:
    """End-to-end call with HOST buffers: pinned queries in, pinned (values, indices) out.

    This is the boundary ``bench.py`` times as ``e2e``: H2D copy of the query embeddings, K1 on the
    queries, K2, K3, D2H copy of the ranking.

    The returned tensors are REUSED pinned buffers (one pair per device, stream and result shape): the
    next ``topk_host`` call with the same shape on the same stream overwrites them -- ``clone()`` what
    has to outlive that call, or pass ``out=(values, indices)`` pinned tensors of your own.
    """
    dev = index.device
    q_dev = queries_host.to(dev, non_blocking=True)
    res = topk(q_dev, index, k, kprime, exact)
    vals, idx = out if out is not None else _pinned_outputs(tuple(res.values.shape), dev)
    vals.copy_(res.values, non_blocking=True)
    idx.copy_(res.indices, non_blocking=True)
    torch.cuda.current_stream(dev).synchronize()
    return vals, idx


class HostTopkPipeline:
    """Back-to-back end-to-end calls with HOST buffers, double-buffered.

    ``topk_host`` serialises copy-in, kernels and copy-out of one batch.  A stream of batches (the
    reference evaluates query sets batch after batch against the same corpus) does not have to:
    every slot owns a CUDA stream, a pinned result buffer and -- through the per-stream workspace
    cache -- its own K2 workspace, so the H2D copy of batch i+1 and the D2H copy of batch i-1 run on
    the copy engines while the kernels of batch i occupy the SMs.  ``submit`` returns at once;
    ``result`` waits for that batch only.  Per batch the same work is done as in ``topk_host``
    (pinned queries in, K1, K2, K3, re-scan, pinned ranking out).
    """

    def __init__(self, index: CorpusIndex, k: int, kprime: int = 0, exact: bool = True, depth: int = 2):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.index, self.k, self.kprime, self.exact = index, k, kprime, exact
        dev = index.device
        self._streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]
        for st in self._streams:                         # the index was built on the caller's stream
            st.wait_stream(torch.cuda.current_stream(dev))
        self._events = [None] * depth
        self._out = [None] * depth
        self._keep = [None] * depth          # device tensors of the batch in flight on a slot
        self._next = 0

    def submit(self, queries_host: torch.Tensor) -> int:
        """Queue one batch of pinned host queries; returns the ticket for ``result``."""
        slot = self._next % len(self._streams)
        if self._events[slot] is not None:
            self._events[slot].synchronize()             # the slot's previous batch has fully left
        dev = self.index.device
        st = self._streams[slot]
        with torch.cuda.stream(st):
            q_dev = queries_host.to(dev, non_blocking=True)
            res = topk(q_dev, self.index, self.k, self.kprime, self.exact)
            shape = tuple(res.values.shape)
            if self._out[slot] is None or tuple(self._out[slot][0].shape) != shape:
                self._out[slot] = (torch.empty(shape, dtype=torch.float32, pin_memory=True),
                                   torch.empty(shape, dtype=torch.int64, pin_memory=True))
            self._out[slot][0].copy_(res.values, non_blocking=True)
            self._out[slot][1].copy_(res.indices, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(st)
        self._events[slot] = ev
        self._keep[slot] = (q_dev, res)                  # alive until the slot is reused
        ticket = self._next
        self._next += 1
        return ticket

    def result(self, ticket: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Pinned (values, indices) of a submitted batch; valid until its slot is submitted to again
        (``depth`` submissions later)."""
        if not (self._next - len(self._streams) <= ticket < self._next):
            raise ValueError(f"ticket {ticket} is not in flight (next {self._next}, depth {len(self._streams)})")
        slot = ticket % len(self._streams)
        self._events[slot].synchronize()
        return self._out[slot]

    @property
    def streams(self):
        return list(self._streams)

    def drain(self):
        for ev in self._events:
            if ev is not None:
                ev.synchronize()
