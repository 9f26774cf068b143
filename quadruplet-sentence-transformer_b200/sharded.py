"""Corpus-sharded retrieval across the GPUs of one node.

One process per GPU (``torch.distributed``, backend ``nccl``; ``gloo`` for the CPU tests of the
plumbing).  Rank r keeps rows ``shard_bounds(N, G, r)`` of the corpus resident in HBM, every rank
sees all queries, computes its exact local top-k (global ids = local row + shard offset), the
per-rank lists are exchanged with ONE all-gather per query tile over NVLink/NVSwitch and merged on
the device (``qst_merge_topk``: ties -> lower global id).  The all-gather of tile t runs on a side
stream while tile t+1 is being scored.

Two exchange strategies:

* ``master="sharded"`` (default when no full master is given): every rank rescoring its own k'
  candidates exactly, one all-gather of the exact per-shard top-k lists, merge (K6).  Works for
  corpora whose fp32 master does not fit one GPU; rescoring work grows with the number of ranks.
* ``master="replicated"``: only the bf16 tensor-core operand is sharded; the fp32 master
  (N*D*4 bytes, 3 GB at 1M x 768) is resident on every rank.  Each shard lists its m best
  candidates per query by bf16 key (``qst_select_candidates``), ONE all-to-all sends every query's
  lists to the rank that owns the query, which rescoring-finalises them (``qst_finalize_lists``) and
  re-scans uncertified ones; an all-gather distributes the final rankings.  Rescoring work per rank
  stays constant as ranks are added.

The reference has no multi-GPU path; its only scale-out knob is the sequential corpus chunk loop
(``corpus_chunk_size``, ``/root/reference/ir_evauation_script.py:161``), which this replaces in
space instead of time.
"""
from __future__ import annotations

import ctypes as C
import os
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


def candidates_per_shard(kprime: int, world: int) -> int:
    """m of ``qst_select_candidates``: twice a shard's fair share of the k' overall candidates plus
    slack (placement of the best documents over shards is uneven), never more than k'."""
    m = ((2 * -(-kprime // world) + 32 + 31) // 32) * 32
    return max(32, min(m, kprime))


def exchange_candidate_lists(lists: torch.Tensor, group=None) -> torch.Tensor:
    """[G*Qown, m+1, 2] int32 (rows grouped by owner rank) -> [G, Qown, m+1, 2] received from every
    shard for the queries this rank owns.  One all-to-all; any backend."""
    world = dist.get_world_size(group)
    out = torch.empty_like(lists)
    dist.all_to_all_single(out, lists.contiguous(), group=group)
    return out.view((world, lists.shape[0] // world) + tuple(lists.shape[1:]))


class PeerHints:
    """Threshold-hint arrays every rank of the node can write (CUDA IPC), two generations.

    Rank r's K2 pushes a query's new threshold into the hint array of every peer with a remote
    atomicMax over NVLink, so all shards filter against the best threshold found on ANY shard
    (``qst_score_select_peers``).  Generation g is used by step g mod 2 and cleared (stream-ordered)
    right after the previous step's K2, so a fast rank never pushes into memory that a slow rank is
    about to clear, and hints of one step never leak into the next (different queries).
    """

    def __init__(self, rows: int, group, device: torch.device):
        lib = _lib.load()
        self.rows, self.group, self.device = rows, group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.gen_bytes = ((rows * 4 + 255) // 256) * 256
        self.step = 0
        self.local = C.c_void_p()
        self.peers = []
        handle = C.create_string_buffer(64)
        ok = 1
        with torch.cuda.device(device):
            if lib.qst_peer_buffer_create(2 * self.gen_bytes, C.byref(self.local), handle) != 0:
                ok, self.local = 0, C.c_void_p()
            mine = torch.tensor(list(handle.raw), dtype=torch.uint8, device=device)
            every = torch.empty(self.world * 64, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(every, mine, group=group)
            every = every.cpu().view(self.world, 64)
            if ok:
                for r in range(self.world):
                    if r == self.rank:
                        continue
                    p = C.c_void_p()
                    if lib.qst_peer_buffer_open(bytes(every[r].tolist()), C.byref(p)) != 0:
                        ok = 0
                        break
                    self.peers.append(p)
            flag = torch.tensor([ok], dtype=torch.int32, device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)   # all ranks or none
        self.ok = bool(int(flag))
        if not self.ok:
            self.close()

    def launch_args(self):
        """(local pointer, ctypes array of peer pointers, n_peers) of the current generation."""
        off = (self.step % 2) * self.gen_bytes
        arr = (C.c_void_p * max(1, len(self.peers)))(*[C.c_void_p(p.value + off) for p in self.peers])
        return C.c_void_p(self.local.value + off), arr, len(self.peers)

    def advance(self, stream_ptr):
        """Call right after K2 was enqueued: clear the other generation for the next step."""
        nxt = ((self.step + 1) % 2) * self.gen_bytes
        _lib.check(_lib.load().qst_peer_buffer_clear(self.local, nxt, self.gen_bytes, stream_ptr))
        self.step += 1

    def close(self):
        lib = _lib.load()
        for p in self.peers:
            lib.qst_peer_buffer_close(p)
        self.peers = []
        if self.local:
            lib.qst_peer_buffer_destroy(self.local)
            self.local = C.c_void_p()


class ShardedCorpus:
    """This rank's shard of an N-row corpus + the collective top-k over all shards."""

    def __init__(self, shard_embeddings: torch.Tensor, n_total: int, score: str = "cos_sim", group=None,
                 query_tile: int = 16384, full_master: Optional[torch.Tensor] = None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = n_total
        self.score = score
        self.start, self.end = shard_bounds(n_total, self.world, self.rank)
        if shard_embeddings.shape[0] != self.end - self.start:
            raise ValueError(f"rank {self.rank} expects rows [{self.start}, {self.end}) = {self.end - self.start} "
                             f"rows, got {shard_embeddings.shape[0]}")
        self.index = scoring.CorpusIndex(shard_embeddings, score, idx_offset=self.start)
        self.query_tile = query_tile
        self._side = torch.cuda.Stream(device=self.index.device) if self.world > 1 else None
        self._peer_hints: Optional[PeerHints] = None
        self._peer_hints_off = bool(os.environ.get("QST_NO_PEER_HINTS"))
        self._timing = [] if os.environ.get("QST_SHARD_TIMING") else None   # debug: per-stage CUDA events
        self.master = None
        if full_master is not None:
            if full_master.shape[0] != n_total:
                raise ValueError(f"full_master must have {n_total} rows, got {full_master.shape[0]}")
            # norms / residual statistics of the whole corpus for rescoring and the certificate; the
            # bf16 operand is NOT built for it
            self.master = scoring.prepare_rows(full_master, scoring.CORPUS_PREP[score], want_bf16=False)

    # ---------------------------------------------------------------------------------------------
    def topk(self, queries: torch.Tensor, k: int, kprime: int = 0, exact: bool = True):
        """Global exact top-k for every query: (values [Q, k], global ids [Q, k], margins [Q])."""
        dev = self.index.device
        queries = queries.to(dev)
        if self.world == 1:
            r = scoring.topk(queries, self.index, k, kprime, exact)
            return r.values, r.indices, r.margin
        if self.master is not None:
            return self._topk_candidate_exchange(queries, k, kprime, exact)
        return self._topk_list_exchange(queries, k, kprime, exact)

    def _hints_for(self, rows: int, dev) -> Optional[PeerHints]:
        """Peer-visible hint arrays for `rows` query rows (collective on first use / growth)."""
        if self._peer_hints_off:
            return None
        if self._peer_hints is None or self._peer_hints.rows < rows:
            if self._peer_hints is not None:
                torch.cuda.synchronize(dev)
                dist.barrier(self.group)          # nobody may still be pushing into the old buffers
                self._peer_hints.close()
            self._peer_hints = PeerHints(rows, self.group, dev)
            if not self._peer_hints.ok:
                self._peer_hints_off = True       # IPC not available here: per-shard thresholds only
                return None
        return self._peer_hints

    def _mark(self, marks, name):
        if marks is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((name, e))

    def enable_stage_timing(self):
        """Record CUDA events between the stages of every following call (cheap; for bench/profiling)."""
        self._timing = []

    def stage_ms(self) -> dict:
        """Mean per-stage device time over the most recent half of the timed calls."""
        if not self._timing:
            return {}
        torch.cuda.synchronize()
        acc = {}
        for marks in self._timing[len(self._timing) // 2:]:
            for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
                acc.setdefault(n1, []).append(e0.elapsed_time(e1))
        return {n: sum(v) / len(v) for n, v in acc.items()}

    def timing_report(self) -> str:
        return "  ".join(f"{n} {v:.3f}" for n, v in self.stage_ms().items())

    # ---- master="replicated": lists of bf16 candidates go to the owner of each query -----------
    def _topk_candidate_exchange(self, queries, k, kprime, exact):
        """All queries in, all rankings out (every rank): slices the batch by owner, runs
        ``topk_owned`` and all-gathers the rankings."""
        dev = self.index.device
        G = self.world
        Q = queries.shape[0]
        q_own = -(-Q // G)
        q_pad = q_own * G
        if q_pad != Q:     # pad with copies of the last query so that every rank owns q_own rows
            queries = torch.cat([queries, queries[-1:].expand(q_pad - Q, -1)])
        vals, idx, margin = self.topk_owned(queries[self.rank * q_own:(self.rank + 1) * q_own], k, kprime, exact)
        with torch.cuda.device(dev):
            gv, gi = all_gather_topk(vals, idx, self.group)                 # [G, q_own, k]
            gm = torch.empty(q_pad, dtype=torch.float32, device=dev)
            dist.all_gather_into_tensor(gm, margin, group=self.group)
            if self._timing:
                self._mark(self._timing[-1], "all_gather")
        return gv.view(q_pad, k)[:Q], gi.view(q_pad, k)[:Q], gm[:Q]

    def topk_owned(self, own_queries: torch.Tensor, k: int, kprime: int = 0, exact: bool = True):
        """Collective: every rank passes ITS OWN slice of the query batch (same number of rows on
        every rank; global query id = rank * rows + i) and gets the exact global top-k of that slice
        back: (values [q_own, k], global ids [q_own, k], margins [q_own]).

        Only the bf16 tensor-core operands of the queries travel between ranks (one all-gather over
        NVLink); fp32 queries, rescoring and results stay with the owner.  Needs the replicated
        fp32 master (``full_master=``).
        """
        if self.master is None:
            raise _lib.QstError("topk_owned needs ShardedCorpus(full_master=...) (candidate exchange)")
        lib = _lib.load()
        dev = self.index.device
        G = self.world
        own_queries = own_queries.to(dev)
        q_own = own_queries.shape[0]
        q_pad = q_own * G
        score = self.score
        cos = score == "cos_sim"
        code = scoring.SCORE_CODES[score]
        marks = [] if self._timing is not None else None
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            self._mark(marks, "start")
            pq = scoring.prepare_rows(own_queries, scoring.QUERY_PREP[score])
            if G > 1:
                q_bf16 = torch.empty((q_pad, pq.bf16.shape[1]), dtype=torch.bfloat16, device=dev)
                dist.all_gather_into_tensor(q_bf16, pq.bf16, group=self.group)
            else:
                q_bf16 = pq.bf16
            self._mark(marks, "prep+gather_q")
            # k' of the whole corpus decides how many candidates every shard lists (m); the shard's
            # own K2 then only has to retain its m best
            kprime_all = scoring.make_plan(q_pad, self.n_total, self.index.d, k, kprime, score).kprime
            m = candidates_per_shard(kprime_all, G)
            plan = scoring.make_plan(q_pad, self.index.n, self.index.d, min(k, m), m, score)
            hints = self._hints_for(plan.m_tiles * plan.rows_per_unit, dev) if G > 1 else None
            if hints is not None:
                # thresholds are shared by the units of ALL shards: size the per-unit retention for
                # k' of the whole corpus spread over stripes x shards units (same Poisson-tail rule as
                # qst_topk_plan_make)
                ku = max(16, -(-(3 * -(-kprime_all // (plan.stripes * G)) + 8) // 8) * 8)
                _lib.check(lib.qst_topk_plan_set_kunit(C.byref(plan), min(ku, plan.kunit)))
            ws = scoring._workspace(plan.ws_bytes, dev, "select")
            if hints is not None:
                local, peers, n_peers = hints.launch_args()
                _lib.check(lib.qst_score_select_peers(C.byref(plan), q_bf16.data_ptr(),
                                                      self.index.rows.bf16.data_ptr(), ws.data_ptr(), local, peers,
                                                      n_peers, st))
                hints.advance(st)
            else:
                _lib.check(lib.qst_score_select(C.byref(plan), q_bf16.data_ptr(), self.index.rows.bf16.data_ptr(),
                                                ws.data_ptr(), st))
            self._mark(marks, "K2")
            lists = torch.empty((q_pad, m + 1, 2), dtype=torch.int32, device=dev)
            _lib.check(lib.qst_select_candidates(C.byref(plan), ws.data_ptr(), m, self.start, lists.data_ptr(), st))
            self._mark(marks, "select")
            recv = exchange_candidate_lists(lists, self.group) if G > 1 else lists.view(1, q_pad, m + 1, 2)
            self._mark(marks, "all_to_all")
            vals = torch.empty((q_own, k), dtype=torch.float32, device=dev)
            idx = torch.empty((q_own, k), dtype=torch.int64, device=dev)
            margin = torch.empty(q_own, dtype=torch.float32, device=dev)
            scratch = scoring._workspace(lib.qst_finalize_lists_scratch_bytes(q_own, G), dev, "lists")
            mst = self.master
            q_inv = pq.inv_norm if cos else None
            _lib.check(lib.qst_finalize_lists(q_own, G, m, k, kprime_all, code, self.index.d, recv.data_ptr(),
                                              pq.f32.data_ptr(), _lib.ptr(q_inv), pq.err.data_ptr(),
                                              mst.f32.data_ptr(), mst.inv_norm.data_ptr() if cos else None,
                                              mst.stats.data_ptr(), vals.data_ptr(), idx.data_ptr(),
                                              margin.data_ptr(), scratch.data_ptr(), st))
            self._mark(marks, "finalize_lists")
            if exact:
                rs = scoring._workspace(lib.qst_exact_rescan_workspace_bytes(q_own, k), dev, "rescan")
                _lib.check(lib.qst_exact_rescan(q_own, self.n_total, self.index.d, k, code, pq.f32.data_ptr(),
                                                _lib.ptr(q_inv), mst.f32.data_ptr(),
                                                mst.inv_norm.data_ptr() if cos else None, 0, vals.data_ptr(),
                                                idx.data_ptr(), margin.data_ptr(), rs.data_ptr(), st))
            self._mark(marks, "rescan")
            if marks is not None:
                self._timing.append(marks)
        return vals, idx, margin

    # ---- master="sharded": exact per-shard top-k lists, all-gather, merge ---------------------------
    def _topk_list_exchange(self, queries, k, kprime, exact):
        dev = self.index.device
        Q = queries.shape[0]
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
