"""Corpus-sharded retrieval across the GPUs of one node.

One process per GPU (``torch.distributed``, backend ``nccl``; ``gloo`` for the CPU tests of the
plumbing).  Rank r keeps rows ``shard_bounds(N, G, r)`` of the corpus resident in HBM, every rank
sees all queries, computes its exact local top-k (global ids = local row + shard offset), the
per-rank lists are exchanged with ONE all-gather per query tile over NVLink/NVSwitch and merged on
the device (``qst_merge_topk``: ties -> lower global id).  The all-gather of tile t runs on a side
stream while tile t+1 is being scored.

The reference has no multi-GPU path; its only scale-out knob is the sequential corpus chunk loop
(``corpus_chunk_size``, ``/root/reference/ir_evauation_script.py:161``), which this replaces in
space instead of time.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib, scoring


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced row range of shard ``rank`` (first ``n % world`` shards get one extra)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_gather_topk(vals: torch.Tensor, idx: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """[Q, k] per rank -> [G, Q, k] on every rank (rank-major).  Works on any backend."""
    world = dist.get_world_size(group)
    vals, idx = vals.contiguous(), idx.contiguous()
    # flat [G*Q, k] outputs (concatenation along dim 0) are accepted by both nccl and gloo
    gv = torch.empty((world * vals.shape[0],) + tuple(vals.shape[1:]), dtype=vals.dtype, device=vals.device)
    gi = torch.empty((world * idx.shape[0],) + tuple(idx.shape[1:]), dtype=idx.dtype, device=idx.device)
    dist.all_gather_into_tensor(gv, vals, group=group)
    dist.all_gather_into_tensor(gi, idx, group=group)
    gv = gv.view((world,) + tuple(vals.shape))
    gi = gi.view((world,) + tuple(idx.shape))
    return gv, gi


def merge_topk(gv: torch.Tensor, gi: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """K6 on the device: G descending lists per query -> global top-k."""
    lib = _lib.load()
    _lib.require_cuda(gv, gi)
    G, Q, k = gv.shape
    dev = gv.device
    with torch.cuda.device(dev):
        ov = torch.empty((Q, k), dtype=torch.float32, device=dev)
        oi = torch.empty((Q, k), dtype=torch.int64, device=dev)
        _lib.check(lib.qst_merge_topk(gv.contiguous().data_ptr(), gi.contiguous().data_ptr(), G, Q, k,
                                      ov.data_ptr(), oi.data_ptr(), _lib.stream_ptr(dev)))
    return ov, oi


class ShardedCorpus:
    """This rank's shard of an N-row corpus + the collective top-k over all shards."""

    def __init__(self, shard_embeddings: torch.Tensor, n_total: int, score: str = "cos_sim", group=None,
                 query_tile: int = 16384):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = n_total
        self.start, self.end = shard_bounds(n_total, self.world, self.rank)
        if shard_embeddings.shape[0] != self.end - self.start:
            raise ValueError(f"rank {self.rank} expects rows [{self.start}, {self.end}) = {self.end - self.start} "
                             f"rows, got {shard_embeddings.shape[0]}")
        self.index = scoring.CorpusIndex(shard_embeddings, score, idx_offset=self.start)
        self.query_tile = query_tile
        self._side = torch.cuda.Stream(device=self.index.device) if self.world > 1 else None

    def topk(self, queries: torch.Tensor, k: int, kprime: int = 0, exact: bool = True):
        """Global exact top-k for every query: (values [Q, k], global ids [Q, k], local margins [Q])."""
        dev = self.index.device
        queries = queries.to(dev)
        Q = queries.shape[0]
        if self.world == 1:
            r = scoring.topk(queries, self.index, k, kprime, exact)
            return r.values, r.indices, r.margin
        main = torch.cuda.current_stream(dev)
        out_v: List[torch.Tensor] = []
        out_i: List[torch.Tensor] = []
        margins: List[torch.Tensor] = []
        pending: Optional[tuple] = None
        for q0 in range(0, Q, self.query_tile):
            r = scoring.topk(queries[q0:q0 + self.query_tile], self.index, k, kprime, exact)
            margins.append(r.margin)
            done = torch.cuda.Event()
            done.record(main)
            if pending is not None:           # finish the previous tile's exchange
                out_v.append(pending[0]); out_i.append(pending[1])
                main.wait_event(pending[2])
            with torch.cuda.stream(self._side):
                self._side.wait_event(done)
                gv, gi = all_gather_topk(r.values, r.indices, self.group)
                mv, mi = merge_topk(gv, gi)
                fin = torch.cuda.Event()
                fin.record(self._side)
            for t in (r.values, r.indices, gv, gi, mv, mi):
                t.record_stream(self._side)
            pending = (mv, mi, fin)
        out_v.append(pending[0]); out_i.append(pending[1])
        main.wait_event(pending[2])
        return torch.cat(out_v), torch.cat(out_i), torch.cat(margins)
