"""Corpus-sharded retrieval across the GPUs of one node.

One process per GPU.  Rank r keeps rows ``shard_bounds(N, G, r)`` of the corpus resident in HBM -- the
bf16 tensor-core operand and, by default, the fp32 master of THOSE ROWS ONLY (the partition of
BASELINE.json's north star).  The collectives go through a ``qst_b200.comm.Comm``: NCCL over
NVLink/NVSwitch (``torch.distributed`` or the C ABI's ``qst_comm_*``), ``gloo`` in the CPU tests, or
threads of one process for the single-GPU test tier.

``ShardedCorpus.topk_owned`` (data-parallel entry: every rank brings the queries it owns):

1. query distribution: K1 on the rank's own slice; bf16 operand + inverse norms all-gathered, fp32 rows
   pushed to every rank by the copy engines underneath K2 (``PeerGather``).  A stream of batches announces
   the next one (``prefetch=``): then all of it is pushed underneath the CURRENT K2 and a step starts with
   no collective at all;
2. K2 of all queries against the local shard; thresholds are shared between the shards THROUGH PEER
   MEMORY while the kernels run (``qst_score_select_peers``);
3. every shard lists its m best candidates per query by bf16 key and the kernel stores each list straight
   into the receive buffer of the rank that owns the query (``qst_select_candidates_scatter``,
   ``ExchangeArena``); a flag barrier over peer memory (``qst_peer_barrier``) is what is left of the
   all-to-all.  Without peer-mapped memory: ``qst_select_candidates`` + ONE all-to-all;
4. sharded master: the owner selects the k' best overall and stores into every shard's buffer the rows it
   wants rescored (``qst_select_requests_scatter`` -> barrier), the shard computes their exact fp32 scores
   from its own rows and stores them into the owners' buffers (``qst_rescore_requests_scatter`` ->
   barrier), ``qst_finalize_exact`` orders them and evaluates the certificate.  Rescoring work per rank is
   q*k' rows whatever G is; per-GPU memory is the shard only.
   Replicated master (``full_master=``, 3 GB per 1M x 768 rows on every rank): the owner
   rescoring-finalises the lists itself (``qst_finalize_lists``), no requests travel;
5. queries whose certificate failed are re-scanned exactly -- by every shard over its own rows, merged
   at the owner (sharded master), or by the owner over its copy (replicated master).

``ShardedCorpus.topk(..., strategy="allgather_merge")`` is the literal "local top-k, all-gather, merge":
every shard finalises its exact local top-k of ALL queries (``qst_finalize_topk``), one all-gather per
query tile, ``qst_merge_topk`` (ties -> lower global id).  Rescoring work grows with G.

The reference has no multi-GPU path; its only scale-out knob is the sequential corpus chunk loop
(``corpus_chunk_size``, ``/root/reference/ir_evauation_script.py:161``), which this replaces in
space instead of time.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib, scoring
from .comm import Comm, TorchComm, default_comm


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
    """Threshold-hint arrays every rank of the node can write, two generations.

    Rank r's K2 pushes a query's new threshold into the hint array of every peer with a remote
    atomicMax over NVLink, so all shards filter against the best threshold found on ANY shard
    (``qst_score_select_peers``).  Generation g is used by step g mod 2 and cleared (stream-ordered)
    right after the previous step's K2, so a fast rank never pushes into memory that a slow rank is
    about to clear, and hints of one step never leak into the next (different queries).  The buffers
    come from the communicator (CUDA IPC between processes, plain allocations between the threads of
    an emulated node).
    """

    def __init__(self, rows: int, comm: Comm, device: torch.device):
        self.rows, self.comm, self.device = rows, comm, device
        self.gen_bytes = ((rows * 4 + 255) // 256) * 256
        self.step = 0
        self.buffers = comm.shared_buffers(2 * self.gen_bytes, device)
        self.ok = self.buffers is not None

    def launch_args(self):
        """(local pointer, ctypes array of peer pointers, n_peers) of the current generation."""
        off = (self.step % 2) * self.gen_bytes
        peers = self.buffers.peers
        arr = (C.c_void_p * max(1, len(peers)))(*[C.c_void_p(p + off) for p in peers])
        return C.c_void_p(self.buffers.local + off), arr, len(peers)

    def advance(self, stream_ptr):
        """Call right after K2 was enqueued: clear the other generation for the next step."""
        nxt = ((self.step + 1) % 2) * self.gen_bytes
        _lib.check(_lib.load().qst_peer_buffer_clear(C.c_void_p(self.buffers.local), nxt, self.gen_bytes, stream_ptr))
        self.step += 1

    def close(self):
        if self.buffers is not None:
            self.buffers.close()
            self.buffers = None


class PeerGather:
    """All-gather of one block per rank and REGION through the COPY ENGINES into peer-mapped buffers
    (regions: the fp32 queries, their bf16 tensor-core operand, their inverse norms).

    The sharded-master path needs the fp32 queries of every rank on every rank -- but only after K2, for
    the exact rescoring.  An NCCL all-gather would sit on the critical path in front of K2 (its kernel
    cannot become resident beside the persistent K2); copy-engine transfers use no SM and run underneath
    K2.  Every rank pushes its block into the gather buffer of every rank on a side stream; the consumer
    kernel is ordered after (a) this rank's own pushes (event) and (b) a later collective in which every
    rank takes part only after ITS pushes have completed -- so every block has landed.  A step that knows
    the NEXT batch (``topk_owned(..., prefetch=)``) pushes all three regions of that batch while its own
    exchanges run, and the next step starts K2 straight from the gathered buffers: no collective in front
    of K2 at all.  Three generations rotate between pushes: a rank can be at most one step ahead of
    another (there are collectives / barriers in every step), so a fast rank never writes into a buffer a
    slow rank is still reading -- and the generation of step i is still intact while step i+1 runs (its
    prefetch writes the third one), which is what lets the deferred exact re-scan of step i wait until
    K2 of step i+1 has been queued."""

    GENERATIONS = 3

    def __init__(self, region_bytes: Sequence[int], comm: Comm, device: torch.device, own_stream: bool):
        self.region_bytes, self.comm, self.device = tuple(int(b) for b in region_bytes), comm, device
        self.region_off, off = [], 0
        for b in self.region_bytes:
            self.region_off.append(off)
            off += ((b * comm.world + 255) // 256) * 256
        self.gen_bytes = off
        self.buffers = comm.shared_buffers(self.GENERATIONS * self.gen_bytes, device)
        self.ok = self.buffers is not None
        self.side = torch.cuda.Stream(device=device) if (self.ok and own_stream) else None
        self.step = 0

    def push(self, blocks: Sequence[Optional[torch.Tensor]], after: Optional[torch.cuda.Event] = None):
        """Start pushing this rank's contiguous block of every region (``None``: that region is not used
        in this generation) to every rank, after ``after`` (default: everything queued so far on the
        current stream).  Returns (this rank's gathered-buffer pointer per region for this generation,
        event to wait for before the collective that publishes the pushes)."""
        lib = _lib.load()
        off = (self.step % self.GENERATIONS) * self.gen_bytes
        self.step += 1
        main = torch.cuda.current_stream(self.device)
        st = self.side if self.side is not None else main
        if st is not main:
            if after is not None:
                st.wait_event(after)
            else:
                st.wait_stream(main)                   # the blocks have been produced on the main stream
        with torch.cuda.stream(st):
            for block, nbytes, roff in zip(blocks, self.region_bytes, self.region_off):
                if block is None:
                    continue
                if block.numel() * block.element_size() != nbytes or not block.is_contiguous():
                    raise ValueError("PeerGather.push: block does not match its region")
                dst_off = off + roff + self.comm.rank * nbytes
                for base in [self.buffers.local] + list(self.buffers.peers):
                    _lib.check(lib.qst_peer_copy(C.c_void_p(base + dst_off), block.data_ptr(), nbytes, st.cuda_stream))
                block.record_stream(st)
            ev = torch.cuda.Event()
            ev.record(st)
        return [self.buffers.local + off + roff for roff in self.region_off], ev

    def close(self):
        if self.buffers is not None:
            self.buffers.close()
            self.buffers = None


class ExchangeArena:
    """Receive buffers of the three exchanges of the sharded step (candidate lists -> owners, requests ->
    shards, exact scores -> owners), peer-mapped, two generations, plus the flag words of the barrier.

    With it the exchanges are FUSED into their producer kernels: ``qst_select_candidates_scatter``,
    ``qst_select_requests_scatter`` and ``qst_rescore_requests_scatter`` store every finished row straight
    into the receiving rank's buffer over NVLink, and what is left of each all-to-all is
    ``qst_peer_barrier`` (every rank signals every rank, release/acquire at system scope).  The receive
    layout is the one the all-to-all would have produced.  Generation g serves step g mod 2: a rank can be
    at most one barrier ahead of another, and a region written in step i was last read in step i-2."""

    LISTS, REQ, EXACT = 0, 1, 2

    def __init__(self, q_own: int, m: int, comm: Comm, device: torch.device):
        G = comm.world
        self.q_own, self.m, self.comm, self.device = q_own, m, comm, device
        sizes = (G * q_own * (m + 1) * 8, G * q_own * m * 4, G * q_own * m * 4)
        self.off, off = [], 0
        for b in sizes:
            self.off.append(off)
            off += ((b + 255) // 256) * 256
        self.gen_bytes = off
        self.buffers = comm.shared_buffers(2 * self.gen_bytes + 256, device) if G <= _lib.QST_MAX_WORLD else None
        self.ok = self.buffers is not None
        self.step, self.epoch = 0, 0
        self._flags = comm.scatter_descriptor(self.buffers, 2 * self.gen_bytes, 1) if self.ok else None

    def _gen(self) -> int:
        return (self.step % 2) * self.gen_bytes

    def dst(self, region: int):
        """Scatter descriptor of `region` in the current generation (rows per block: q_own)."""
        return self.comm.scatter_descriptor(self.buffers, self._gen() + self.off[region], self.q_own)

    def local(self, region: int) -> int:
        return self.buffers.local + self._gen() + self.off[region]

    def barrier(self):
        self.epoch += 1
        self.comm.peer_barrier(self._flags, self.epoch & 0xffffffff, self.device)

    def advance(self):
        self.step += 1

    def close(self):
        if self.buffers is not None:
            self.buffers.close()
            self.buffers = None


class ShardedCorpus:
    """This rank's shard of an N-row corpus + the collective top-k over all shards.

    ``full_master=None`` (default, the partition of BASELINE.json's north star): this rank keeps ITS rows
    only -- bf16 tensor-core operand AND fp32 master.  ``full_master=<[N, D] tensor>``: the fp32 master is
    replicated on every rank (3 GB per 1M x 768 rows), only the bf16 operand is sharded.
    ``comm``: a ``qst_b200.comm.Comm`` (default: ``torch.distributed``'s default group, or a world of one).
    """

    def __init__(self, shard_embeddings: torch.Tensor, n_total: int, score: str = "cos_sim", group=None,
                 query_tile: int = 16384, full_master: Optional[torch.Tensor] = None, comm: Optional[Comm] = None):
        self.comm = comm if comm is not None else default_comm(group)
        self.group = group
        self.world, self.rank = self.comm.world, self.comm.rank
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
        self._peer_gather: Optional[PeerGather] = None
        self._peer_gather_off = bool(os.environ.get("QST_NO_PEER_GATHER"))
        self._arena: Optional[ExchangeArena] = None
        self._arena_off = bool(os.environ.get("QST_NO_FUSED_EXCHANGE"))
        self._timing = [] if os.environ.get("QST_SHARD_TIMING") else None   # debug: per-stage CUDA events
        self.master = None
        self.last_rescanned = 0      # queries repaired by the distributed exact re-scan in the last call
        self._pending = None         # deferred certificate check of the last call (exact="deferred")
        self._prefetched = None      # next batch, already prepared and distributed (topk_owned(prefetch=))
        if full_master is not None:
            if full_master.shape[0] != n_total:
                raise ValueError(f"full_master must have {n_total} rows, got {full_master.shape[0]}")
            # norms / residual statistics of the whole corpus for rescoring and the certificate; the
            # bf16 operand is NOT built for it
            self.master = scoring.prepare_rows(full_master, scoring.CORPUS_PREP[score], want_bf16=False)
            self.global_stats = self.master.stats
        else:
            # certificate statistics (max rounding residual, max norm) over ALL shards
            self.global_stats = self.comm.all_reduce_max(self.index.rows.stats) if self.world > 1 \
                else self.index.rows.stats

    def close(self):
        """Releases the peer-mapped buffers (threshold hints, query gather).  Collective in effect: call it
        on every rank, after the last ``topk`` / ``topk_owned`` / ``finish_exact``."""
        self.finish_exact()
        self._prefetched = None
        dev = self.index.device
        if self._peer_hints is not None or self._peer_gather is not None or self._arena is not None:
            torch.cuda.synchronize(dev)
            if self.world > 1:
                self.comm.barrier()          # nobody may still be writing into a buffer that goes away
        for holder in ("_peer_hints", "_peer_gather", "_arena"):
            obj = getattr(self, holder)
            if obj is not None:
                obj.close()
                setattr(self, holder, None)

    @property
    def master_mode(self) -> str:
        return "replicated" if self.master is not None else "sharded"

    def _ws(self, nbytes: int, tag: str) -> torch.Tensor:
        # per-rank tags: the ranks of an emulated node (LocalComm) share one device and one stream
        return scoring._workspace(nbytes, self.index.device, f"{tag}.r{self.rank}")

    # ---------------------------------------------------------------------------------------------
    def topk(self, queries: torch.Tensor, k: int, kprime: int = 0, exact: bool = True,
             strategy: str = "owners"):
        """Global exact top-k for every query, on every rank: (values [Q, k], global ids [Q, k],
        margins [Q]).  ``strategy="owners"``: the batch is cut into one slice per rank, every rank
        finalises its slice (``topk_owned``) and the rankings are all-gathered.
        ``strategy="allgather_merge"``: every rank computes its exact LOCAL top-k of all queries, the
        lists are all-gathered per query tile and merged (``qst_merge_topk``)."""
        dev = self.index.device
        queries = queries.to(dev)
        if self.world == 1:
            r = scoring.topk(queries, self.index, k, kprime, exact)
            return r.values, r.indices, r.margin
        if strategy == "allgather_merge":
            return self._topk_list_exchange(queries, k, kprime, exact)
        if strategy != "owners":
            raise ValueError(f"strategy must be 'owners' or 'allgather_merge', {strategy!r} given")
        G, Q = self.world, queries.shape[0]
        q_own = -(-Q // G)
        q_pad = q_own * G
        if q_pad != Q:     # pad with copies of the last query so that every rank owns q_own rows
            queries = torch.cat([queries, queries[-1:].expand(q_pad - Q, -1)])
        vals, idx, margin = self.topk_owned(queries[self.rank * q_own:(self.rank + 1) * q_own], k, kprime, exact)
        with torch.cuda.device(dev):
            gv, gi, gm = self.comm.all_gather(vals), self.comm.all_gather(idx), self.comm.all_gather(margin)
            if self._timing:
                self._mark(self._timing[-1], "all_gather")
        return gv[:Q], gi[:Q], gm[:Q]

    def _hints_for(self, rows: int, dev) -> Optional[PeerHints]:
        """Peer-visible hint arrays for `rows` query rows (collective on first use / growth)."""
        if self._peer_hints_off:
            return None
        if self._peer_hints is None or self._peer_hints.rows < rows:
            if self._peer_hints is not None:
                torch.cuda.synchronize(dev)
                self.comm.barrier()               # nobody may still be pushing into the old buffers
                self._peer_hints.close()
            self._peer_hints = PeerHints(rows, self.comm, dev)
            if not self._peer_hints.ok:
                self._peer_hints_off = True       # no peer-writable memory here: per-shard thresholds only
                return None
        return self._peer_hints

    def _gather_for(self, q_own: int, dev) -> Optional[PeerGather]:
        """Copy-engine gather buffers for `q_own` query rows per rank: regions fp32 rows / bf16 operand /
        inverse norms (collective on first use / resize)."""
        if self._peer_gather_off or self.world == 1:
            return None
        D, d_pad = self.index.d, self.index.rows.bf16.shape[1]
        regions = (q_own * D * 4, q_own * d_pad * 2, q_own * 4)
        if self._peer_gather is None or self._peer_gather.region_bytes != regions:
            if self._peer_gather is not None:
                torch.cuda.synchronize(dev)
                self.comm.barrier()
                self._peer_gather.close()
            from .comm import LocalComm
            self._peer_gather = PeerGather(regions, self.comm, dev, own_stream=not isinstance(self.comm, LocalComm))
            if not self._peer_gather.ok:
                self._peer_gather_off = True
                return None
        return self._peer_gather

    def _arena_for(self, q_own: int, m: int, dev) -> Optional[ExchangeArena]:
        """Peer-mapped receive buffers for the fused exchanges (collective on first use / resize)."""
        if self._arena_off or self.world == 1:
            return None
        if self._arena is None or (self._arena.q_own, self._arena.m) != (q_own, m):
            if self._arena is not None:
                torch.cuda.synchronize(dev)
                self.comm.barrier()
                self._arena.close()
            self._arena = ExchangeArena(q_own, m, self.comm, dev)
            if not self._arena.ok:
                self._arena_off = True
                self._arena = None
        return self._arena

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

    # ---- the data-parallel entry: every rank brings the queries it owns ---------------------------
    def topk_owned(self, own_queries: torch.Tensor, k: int, kprime: int = 0, exact: bool = True,
                   prefetch: Optional[torch.Tensor] = None):
        """Collective: every rank passes ITS OWN slice of the query batch (same number of rows on
        every rank; global query id = rank * rows + i) and gets the exact global top-k of that slice
        back: (values [q_own, k], global ids [q_own, k], margins [q_own]).

        Every rank scores all G*q queries against its shard on the tensor cores (K2, thresholds shared
        between the shards through peer memory), lists its m best candidates per query by bf16 key, and
        ONE all-to-all routes each query's G lists to the rank that owns the query.  Then

        * sharded master (default): the owner picks the k' best overall and asks each shard for the
          exact fp32 scores of the rows it holds (``qst_select_requests`` -> all-to-all ->
          ``qst_rescore_requests`` on the shard, from its own fp32 rows -> all-to-all ->
          ``qst_finalize_exact``).  The fp32 queries are all-gathered once per call; rescoring work per
          rank is the same q*k' rows as on one GPU.
        * replicated master: the owner rescoring-finalises the lists itself from its copy of the whole
          fp32 corpus (``qst_finalize_lists``); only bf16 query operands travel.

        ``exact=True``: queries whose certificate failed are re-scanned exactly before the call returns;
        with a sharded master that costs ONE host read per call ("is anything flagged, anywhere?").
        ``exact="deferred"`` (sharded master) moves that read to the next ``topk_owned`` call or to
        ``finish_exact()``, whichever comes first -- by then the flag has long arrived, nothing stalls --
        and patches the returned tensors in place in the rare case a re-scan is needed.  A stream of
        batches should use it and call ``finish_exact()`` after the last one.

        ``prefetch`` (sharded master, peer-mapped buffers available): this rank's slice of the NEXT batch
        -- same shape on every rank, and every rank passes one or none.  Its K1 runs in front of this
        call's K2 and the distribution of its fp32 rows / bf16 operand / inverse norms to all ranks (copy
        engines, no SM) underneath it, so the next call
        -- when it is given that very tensor, unmodified -- starts K2 at once instead of after an
        all-gather (any other tensor: the prefetched batch is discarded; like every collective, all
        ranks must make the same sequence of calls).  A query set processed in tiles passes tile t+1
        while asking for tile t.
        """
        if exact == "deferred" and (self.master is not None or self.world == 1):
            exact = True        # the replicated-master / single-GPU re-scans are device-driven: nothing to defer
        if self.master is not None:
            return self._topk_owned_replicated(own_queries, k, kprime, exact)
        return self._topk_owned_sharded(own_queries, k, kprime, exact, prefetch)

    def _select_pass(self, q_bf16, q_pad, k, kprime, marks, fused=False):
        """K2 on the local shard for all q_pad queries + the per-query candidate lists by bf16 key.
        Returns (lists [q_pad, m+1, 2] int32, m, k' of the whole corpus) -- or, with ``fused`` and
        peer-mapped receive buffers, (the ExchangeArena the lists were scattered into, m, k')."""
        lib = _lib.load()
        dev = self.index.device
        G, score = self.world, self.score
        st = _lib.stream_ptr(dev)
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
        ws = self._ws(plan.ws_bytes, "select")
        if hints is not None:
            local, peers, n_peers = hints.launch_args()
            _lib.check(lib.qst_score_select_peers(C.byref(plan), _lib.ptr(q_bf16), self.index.rows.bf16.data_ptr(),
                                                  ws.data_ptr(), local, peers, n_peers, st))
            hints.advance(st)
        else:
            _lib.check(lib.qst_score_select(C.byref(plan), _lib.ptr(q_bf16), self.index.rows.bf16.data_ptr(),
                                            ws.data_ptr(), st))
        self._mark(marks, "K2")
        arena = self._arena_for(q_pad // G, m, dev) if fused else None
        if arena is not None:
            # every list goes straight into its owner's receive buffer, written by the kernel that makes it
            dst = arena.dst(ExchangeArena.LISTS)
            _lib.check(lib.qst_select_candidates_scatter(C.byref(plan), ws.data_ptr(), m, self.start, C.byref(dst), st))
            self._mark(marks, "select")
            return arena, m, kprime_all
        lists = torch.empty((q_pad, m + 1, 2), dtype=torch.int32, device=dev)
        _lib.check(lib.qst_select_candidates(C.byref(plan), ws.data_ptr(), m, self.start, lists.data_ptr(), st))
        self._mark(marks, "select")
        return lists, m, kprime_all

    # ---- fp32 master sharded: requests to the shards, exact scores back ----------------------------
    @staticmethod
    def _tensor_key(t: torch.Tensor):
        return (t.data_ptr(), tuple(t.shape), t.dtype, t._version)

    def _start_prefetch(self, nxt: torch.Tensor, gather: PeerGather):
        """K1 of the next batch's slice on the current stream (50 us, in front of this step's K2), then the
        copy-engine pushes of all three regions on the gather's side stream: they run underneath K2 (the
        8-GPU case moves 7 x 46 MB per rank, ~1.5 ms of copy-engine time -- started after K2 it would land
        on the critical path of the exchanges)."""
        dev = self.index.device
        score = self.score
        pq = scoring.prepare_rows(nxt.to(dev).float().contiguous(), scoring.QUERY_PREP[score])
        ptrs, pushed = gather.push([pq.f32, pq.bf16, pq.inv_norm if score == "cos_sim" else None])
        self._prefetched = (self._tensor_key(nxt), nxt.shape[0], pq, ptrs, pushed)

    def _topk_owned_sharded(self, own_queries, k, kprime, exact, prefetch=None):
        lib = _lib.load()
        dev = self.index.device
        comm, G, r = self.comm, self.world, self.rank
        pre, self._prefetched = getattr(self, "_prefetched", None), None
        if pre is not None and pre[0] != self._tensor_key(own_queries):
            pre = None      # not the batch that was announced (or modified since): discarded, plain path
        if pre is None:
            self.finish_exact()             # the previous call's deferred certificate check, if any
        # (a prefetched batch: that check waits until this call's K2 has been queued -- the host read of the
        #  previous step's flag then costs the GPU nothing; the previous step's gathered queries stay intact,
        #  this call's prefetch writes the third generation)
        own_queries = own_queries.to(dev)
        q_own = own_queries.shape[0]
        q_pad = q_own * G
        score = self.score
        cos = score == "cos_sim"
        code = scoring.SCORE_CODES[score]
        D = self.index.d
        marks = [] if self._timing is not None else None
        with torch.cuda.device(dev):
            st = _lib.stream_ptr(dev)
            self._mark(marks, "start")
            # Every rank rescoring rows of ITS shard needs the fp32 queries of all ranks -- but only
            # AFTER K2.  With peer-mapped buffers they are pushed by the copy engines underneath K2 and
            # only the bf16 operands (+ inverse norms) are all-gathered in front of it; without, the fp32
            # queries are all-gathered and K1 runs on all of them locally.  A batch the previous call
            # prefetched is already here in full.
            gather = self._gather_for(q_own, dev) if G > 1 else None
            if prefetch is not None and (gather is None or prefetch.shape[0] != q_own):
                prefetch = None             # no peer-mapped buffers (or a last, shorter tile): plain path next time
            if pre is not None:
                _, _, pq_own, (q_all_ptr, q_bf16, q_inv_all), _ = pre   # published by the previous call's last exchange
                if not cos:
                    q_inv_all = None
                pushed = None
                own_f32_ptr, own_err_ptr = pq_own.f32.data_ptr(), pq_own.err.data_ptr()
                keep = (pq_own,)
            elif gather is not None:
                own_f32 = own_queries.float().contiguous()
                pq_own = scoring.prepare_rows(own_f32, scoring.QUERY_PREP[score])
                q_bf16 = comm.all_gather(pq_own.bf16)
                q_inv_all = comm.all_gather(pq_own.inv_norm) if cos else None
                # pushed AFTER the all-gathers are queued (the side stream waits for them): the pushes then
                # share NVLink with nothing and run entirely underneath K2
                (q_all_ptr, _, _), pushed = gather.push([pq_own.f32, None, None])
                own_f32_ptr, own_err_ptr = pq_own.f32.data_ptr(), pq_own.err.data_ptr()
                keep = (pq_own, q_bf16, q_inv_all)
            else:
                own_f32 = own_queries.float().contiguous()
                q_all = comm.all_gather(own_f32) if G > 1 else own_f32
                pq = scoring.prepare_rows(q_all, scoring.QUERY_PREP[score])
                own = slice(r * q_own, (r + 1) * q_own)
                q_all_ptr, pushed, q_bf16, q_inv_all = pq.f32.data_ptr(), None, pq.bf16, (pq.inv_norm if cos else None)
                own_f32_ptr, own_err_ptr = pq.f32[own].data_ptr(), pq.err[own].data_ptr()
                keep = (pq,)
            if prefetch is not None:
                # safe with respect to the other ranks' reads of the buffers being overwritten: this call's
                # finish_exact() above was the last (collective) reader of the previous generation
                self._start_prefetch(prefetch, gather)
            self._mark(marks, "gather_q+prep")
            lists, m, kprime_all = self._select_pass(q_bf16, q_pad, k, kprime, marks, fused=G > 1)
            arena = lists if isinstance(lists, ExchangeArena) else None
            if pre is not None:
                self.finish_exact()         # previous call's deferred check, underneath this call's K2
            if pushed is not None:
                # this rank's pushes are done before it enters the exchange; the exchange completes only
                # after every rank has entered it, i.e. after every rank's pushes are done
                torch.cuda.current_stream(dev).wait_event(pushed)
            req = torch.empty((G * q_own, m), dtype=torch.int32, device=dev)      # [G shards, q_own, m]
            bound = torch.empty(q_own, dtype=torch.int32, device=dev)
            scratch = self._ws(lib.qst_finalize_lists_scratch_bytes(q_own, G), "lists")
            exact_out = torch.empty((G * q_own, m), dtype=torch.float32, device=dev)
            c = self.index.rows
            if arena is not None:
                # Exchanges fused into their producers: the rows were / are stored straight into the
                # receiving rank's buffer by the kernel that makes them; a flag barrier over peer memory
                # is all that is left of each all-to-all.
                arena.barrier()                                                   # lists have landed everywhere
                self._mark(marks, "all_to_all")
                dst = arena.dst(ExchangeArena.REQ)
                _lib.check(lib.qst_select_requests_scatter(q_own, G, m, kprime_all, self.n_total,
                                                           arena.local(ExchangeArena.LISTS), req.data_ptr(),
                                                           bound.data_ptr(), scratch.data_ptr(), C.byref(dst), st))
                self._mark(marks, "requests")
                arena.barrier()                                                   # requests have landed
                dst = arena.dst(ExchangeArena.EXACT)
                _lib.check(lib.qst_rescore_requests_scatter(G * q_own, m, D, code, arena.local(ExchangeArena.REQ),
                                                            q_all_ptr, _lib.ptr(q_inv_all), c.f32.data_ptr(),
                                                            c.inv_norm.data_ptr() if cos else None,
                                                            exact_out.data_ptr(), C.byref(dst), st))
                self._mark(marks, "rescore")
                if prefetch is not None:
                    # the next batch's pushes (underneath K2, long done) are published by this barrier
                    torch.cuda.current_stream(dev).wait_event(self._prefetched[4])
                arena.barrier()                                                   # exact scores have landed
                exact_in_ptr = arena.local(ExchangeArena.EXACT)
                arena.advance()
            else:
                recv = comm.all_to_all(lists) if G > 1 else lists                 # [G, q_own, m+1, 2]
                self._mark(marks, "all_to_all")
                _lib.check(lib.qst_select_requests(q_own, G, m, kprime_all, self.n_total, recv.data_ptr(),
                                                   req.data_ptr(), bound.data_ptr(), scratch.data_ptr(), st))
                self._mark(marks, "requests")
                req_in = comm.all_to_all(req) if G > 1 else req                   # [G owners, q_own, m]
                _lib.check(lib.qst_rescore_requests(G * q_own, m, D, code, req_in.data_ptr(), q_all_ptr,
                                                    _lib.ptr(q_inv_all), c.f32.data_ptr(),
                                                    c.inv_norm.data_ptr() if cos else None, exact_out.data_ptr(), st))
                self._mark(marks, "rescore")
                if prefetch is not None:
                    # the next batch's pushes (underneath K2, long done) are published by this exchange
                    torch.cuda.current_stream(dev).wait_event(self._prefetched[4])
                exact_in = comm.all_to_all(exact_out) if G > 1 else exact_out     # [G shards, q_own, m]
                exact_in_ptr = exact_in.data_ptr()
            vals = torch.empty((q_own, k), dtype=torch.float32, device=dev)
            idx = torch.empty((q_own, k), dtype=torch.int64, device=dev)
            margin = torch.empty(q_own, dtype=torch.float32, device=dev)
            _lib.check(lib.qst_finalize_exact(q_own, G, m, k, code, D, self.n_total, req.data_ptr(),
                                              exact_in_ptr, bound.data_ptr(), own_f32_ptr,
                                              own_err_ptr, self.global_stats.data_ptr(),
                                              vals.data_ptr(), idx.data_ptr(), margin.data_ptr(), st))
            self._mark(marks, "replies+finalize")
            if exact == "deferred":
                # decide about the re-scan one call later (or in finish_exact()): nothing stalls here
                state = torch.stack([margin, vals[:, k - 1]], dim=1).contiguous()
                all_state = comm.all_gather(state) if G > 1 else state
                flag = torch.zeros(1, dtype=torch.int32, pin_memory=True)
                flag.copy_((~(all_state[:, 0] > 0)).any().to(torch.int32).view(1), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                self._pending = (ev, flag, all_state, q_all_ptr, q_inv_all, q_own, k, vals, idx, margin, keep)
            elif exact:
                self._distributed_rescan(q_all_ptr, q_inv_all, q_own, k, vals, idx, margin)
            del keep
            self._mark(marks, "rescan")
            if marks is not None:
                self._timing.append(marks)
        return vals, idx, margin

    def finish_exact(self) -> int:
        """Completes a call made with ``exact="deferred"``: waits for that call's certificate flag (long
        since on the host when this runs one call later) and, if any query anywhere was left uncertified,
        runs the distributed exact re-scan and patches that call's result tensors IN PLACE.  Collective
        (every rank sees the same flag).  Returns the number of queries that were re-scanned."""
        pend, self._pending = getattr(self, "_pending", None), None
        if pend is None:
            return 0
        ev, flag, all_state, q_all_ptr, q_inv_all, q_own, k, vals, idx, margin, keep = pend
        ev.synchronize()
        if int(flag[0]) == 0:
            self.last_rescanned = 0
            return 0
        with torch.cuda.device(self.index.device):
            self._distributed_rescan(q_all_ptr, q_inv_all, q_own, k, vals, idx, margin, all_state=all_state)
        return self.last_rescanned

    def _distributed_rescan(self, q_all_ptr, q_inv_all, q_own, k, vals, idx, margin, all_state=None):
        """Backstop of the sharded-master path: queries whose certificate failed are re-scanned in fp32
        by EVERY shard against its own rows (``qst_exact_rescan_lists``: rows scoring at least the
        owner's current k-th exact score), the G lists go back to the owner and are merged.  Needs one
        host read of "is anything flagged" per call (the single-GPU re-scan is device-driven)."""
        lib = _lib.load()
        dev = self.index.device
        comm, G, r = self.comm, self.world, self.rank
        cos = self.score == "cos_sim"
        code = scoring.SCORE_CODES[self.score]
        st = _lib.stream_ptr(dev)
        if all_state is None:
            state = torch.stack([margin, vals[:, k - 1]], dim=1).contiguous()         # [q_own, 2]
            all_state = comm.all_gather(state) if G > 1 else state                     # [G*q_own, 2]
        flagged = ~(all_state[:, 0] > 0)
        n_flagged = int(flagged.sum())                                                 # host sync, same on all ranks
        self.last_rescanned = n_flagged
        if n_flagged == 0:
            return
        kth = all_state[:, 1].contiguous()
        rank_of = torch.cumsum(flagged.to(torch.int64), 0) - 1                         # 0-based rank among flagged
        n_rows = G * q_own
        c = self.index.rows
        scratch = self._ws(lib.qst_exact_rescan_workspace_bytes(n_rows, k), "rescan")
        lists_v = torch.full((n_rows, k), float("-inf"), dtype=torch.float32, device=dev)
        lists_i = torch.full((n_rows, k), -1, dtype=torch.int64, device=dev)
        overflow = torch.zeros(n_rows, dtype=torch.int32, device=dev)
        for lo in range(0, n_flagged, 8192):                                           # one pass serves 8192 queries
            sel = flagged & (rank_of >= lo) & (rank_of < lo + 8192)
            m_in = torch.where(sel, torch.full_like(kth, -1.0), torch.ones_like(kth))
            _lib.check(lib.qst_exact_rescan_lists(n_rows, self.index.n, self.index.d, k, code, q_all_ptr,
                                                  _lib.ptr(q_inv_all) if cos else None, c.f32.data_ptr(),
                                                  c.inv_norm.data_ptr() if cos else None, self.start, kth.data_ptr(),
                                                  m_in.data_ptr(), lists_v.data_ptr(), lists_i.data_ptr(),
                                                  overflow.data_ptr(), scratch.data_ptr(), st))
        if G > 1:
            lists_v, lists_i = comm.all_to_all(lists_v), comm.all_to_all(lists_i)      # [G shards, q_own, k]
            overflow = comm.all_to_all(overflow.view(n_rows, 1)).view(G, q_own)
        else:
            overflow = overflow.view(1, q_own)
        mv, mi = merge_topk(lists_v.view(G, q_own, k), lists_i.view(G, q_own, k))
        mine = flagged[r * q_own:(r + 1) * q_own]
        vals[mine] = mv[mine]
        idx[mine] = mi[mine]
        repaired = torch.where(overflow.amax(dim=0) > 0, torch.zeros_like(margin), torch.full_like(margin, float("inf")))
        margin[mine] = repaired[mine]

    # ---- fp32 master replicated: lists of bf16 candidates go to the owner of each query -----------
    def _topk_owned_replicated(self, own_queries, k, kprime, exact):
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
            q_bf16 = self.comm.all_gather(pq.bf16) if G > 1 else pq.bf16
            self._mark(marks, "prep+gather_q")
            lists, m, kprime_all = self._select_pass(q_bf16, q_pad, k, kprime, marks, fused=G > 1)
            if isinstance(lists, ExchangeArena):
                lists.barrier()                     # the lists were stored into their owners' buffers by the kernel
                recv_ptr = lists.local(ExchangeArena.LISTS)
                lists.advance()
            else:
                recv = self.comm.all_to_all(lists) if G > 1 else lists            # [G, q_own, m+1, 2]
                recv_ptr = recv.data_ptr()
            self._mark(marks, "all_to_all")
            vals = torch.empty((q_own, k), dtype=torch.float32, device=dev)
            idx = torch.empty((q_own, k), dtype=torch.int64, device=dev)
            margin = torch.empty(q_own, dtype=torch.float32, device=dev)
            scratch = self._ws(lib.qst_finalize_lists_scratch_bytes(q_own, G), "lists")
            mst = self.master
            q_inv = pq.inv_norm if cos else None
            _lib.check(lib.qst_finalize_lists(q_own, G, m, k, kprime_all, code, self.index.d, recv_ptr,
                                              pq.f32.data_ptr(), _lib.ptr(q_inv), pq.err.data_ptr(),
                                              mst.f32.data_ptr(), mst.inv_norm.data_ptr() if cos else None,
                                              mst.stats.data_ptr(), vals.data_ptr(), idx.data_ptr(),
                                              margin.data_ptr(), scratch.data_ptr(), st))
            self._mark(marks, "finalize_lists")
            if exact:
                rs = self._ws(lib.qst_exact_rescan_workspace_bytes(q_own, k), "rescan")
                _lib.check(lib.qst_exact_rescan(q_own, self.n_total, self.index.d, k, code, pq.f32.data_ptr(),
                                                _lib.ptr(q_inv), mst.f32.data_ptr(),
                                                mst.inv_norm.data_ptr() if cos else None, 0, vals.data_ptr(),
                                                idx.data_ptr(), margin.data_ptr(), rs.data_ptr(), st))
            self._mark(marks, "rescan")
            if marks is not None:
                self._timing.append(marks)
        return vals, idx, margin

    # ---- literal "local exact top-k, all-gather, merge" (north-star wording; more rescoring work) ----
    def _topk_list_exchange(self, queries, k, kprime, exact):
        dev = self.index.device
        Q = queries.shape[0]
        main = torch.cuda.current_stream(dev)
        side = self._side if isinstance(self.comm, TorchComm) else main    # emulated ranks share one stream
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
            with torch.cuda.stream(side):
                side.wait_event(done)
                gv = self.comm.all_gather(r.values).view(self.world, -1, k)
                gi = self.comm.all_gather(r.indices).view(self.world, -1, k)
                mv, mi = merge_topk(gv, gi)
                fin = torch.cuda.Event()
                fin.record(side)
            for t in (r.values, r.indices, gv, gi, mv, mi):
                t.record_stream(side)
            pending = (mv, mi, fin)
        out_v.append(pending[0]); out_i.append(pending[1])
        main.wait_event(pending[2])
        # a query is certified when every shard certified its local list
        margin = torch.cat(margins)
        margin = -self.comm.all_reduce_max(-margin)
        return torch.cat(out_v), torch.cat(out_i), margin


class ShardedHostPipeline:
    """Back-to-back ``topk_owned`` steps with HOST buffers on every rank, double-buffered (the sharded
    counterpart of ``scoring.HostTopkPipeline``).

    Per step and rank: pinned fp32 queries of the rank's slice in, K1 / K2 / exchanges / exact
    rescoring, pinned (values, indices) of that slice out.  Every slot owns a CUDA stream and pinned
    result buffers, so the H2D copy of step i+1 and the D2H copy of step i-1 run on the copy engines
    while the kernels and collectives of step i occupy the SMs and NVLink.  The kernels of consecutive
    steps are chained by an event (K2 fills the machine anyway, and the two generations of the
    peer-shared threshold hints assume steps in order).  Collective: all ranks call ``submit`` the same
    number of times.
    """

    def __init__(self, corp: ShardedCorpus, k: int, kprime: int = 0, exact: bool = True, depth: int = 2):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.corp, self.k, self.kprime, self.exact = corp, k, kprime, exact
        dev = corp.index.device
        self._streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]
        for st in self._streams:
            st.wait_stream(torch.cuda.current_stream(dev))
        self._done = [None] * depth        # slot fully finished (results on the host)
        self._out = [None] * depth
        self._keep = [None] * depth
        self._kernels = None               # event: kernels of the most recent step have been queued and run
        self._announced = None             # (host tensor, its version, device copy) of the batch announced last
        self._next = 0

    def submit(self, own_queries_host: torch.Tensor, next_queries_host: Optional[torch.Tensor] = None) -> int:
        """``next_queries_host``: the pinned slice the NEXT submit will bring (same rows on every rank, all
        ranks or none).  It is copied in now and announced to ``topk_owned(prefetch=)``, so its K1 and
        its distribution over the ranks run underneath this step's exchanges."""
        slot = self._next % len(self._streams)
        if self._done[slot] is not None:
            self._done[slot].synchronize()
        dev = self.corp.index.device
        st = self._streams[slot]
        with torch.cuda.stream(st):
            ann, self._announced = self._announced, None
            if ann is not None and ann[0] is own_queries_host and ann[1] == own_queries_host._version:
                q_dev = ann[2]                                       # copied in by the previous submit
                q_dev.record_stream(st)
            else:
                q_dev = own_queries_host.to(dev, non_blocking=True)  # overlaps the previous step's kernels
            nxt_dev = None
            if next_queries_host is not None and self.corp.master is None and self.corp.world > 1:
                nxt_dev = next_queries_host.to(dev, non_blocking=True)
                self._announced = (next_queries_host, next_queries_host._version, nxt_dev)
            if self._kernels is not None:
                st.wait_event(self._kernels)
            # the distributed exact re-scan needs a host decision ("is anything flagged, anywhere?"),
            # which would stall the pipeline every step: run the first pass only, ship the global flag
            # to the host with the results, and repair in result() in the (rare) case it is set
            deferred = self.exact and self.corp.master is None and self.corp.world > 1
            vals, idx, margin = self.corp.topk_owned(q_dev, self.k, self.kprime, self.exact and not deferred,
                                                     **({"prefetch": nxt_dev} if nxt_dev is not None else {}))
            flag = None
            if deferred:
                flag = self.corp.comm.all_reduce_max((~(margin > 0)).any().to(torch.float32).view(1))
            ev_k = torch.cuda.Event()
            ev_k.record(st)
            shape = tuple(vals.shape)
            if self._out[slot] is None or tuple(self._out[slot][0].shape) != shape:
                self._out[slot] = (torch.empty(shape, dtype=torch.float32, pin_memory=True),
                                   torch.empty(shape, dtype=torch.int64, pin_memory=True),
                                   torch.zeros(1, dtype=torch.float32, pin_memory=True))
            self._out[slot][0].copy_(vals, non_blocking=True)
            self._out[slot][1].copy_(idx, non_blocking=True)
            if flag is not None:
                self._out[slot][2].copy_(flag, non_blocking=True)
            else:
                self._out[slot][2].zero_()
            ev = torch.cuda.Event()
            ev.record(st)
        self._kernels, self._done[slot] = ev_k, ev
        self._keep[slot] = (q_dev, vals, idx, margin)
        ticket = self._next
        self._next += 1
        return ticket

    def result(self, ticket: int) -> Tuple[torch.Tensor, torch.Tensor]:
        if not (self._next - len(self._streams) <= ticket < self._next):
            raise ValueError(f"ticket {ticket} is not in flight (next {self._next}, depth {len(self._streams)})")
        slot = ticket % len(self._streams)
        self._done[slot].synchronize()
        out = self._out[slot]
        if float(out[2][0]) > 0:
            # some rank holds an uncertified query: every rank sees the same flag and repeats the step
            # with the exact re-scan, synchronously (collective)
            st = self._streams[slot]
            with torch.cuda.stream(st):
                if self._kernels is not None:
                    st.wait_event(self._kernels)
                vals, idx, _ = self.corp.topk_owned(self._keep[slot][0], self.k, self.kprime, True)
                out[0].copy_(vals, non_blocking=True)
                out[1].copy_(idx, non_blocking=True)
                self._kernels = torch.cuda.Event()
                self._kernels.record(st)
            st.synchronize()
            out[2].zero_()
        return out[0], out[1]

    @property
    def streams(self):
        return list(self._streams)

    def drain(self):
        for ev in self._done:
            if ev is not None:
                ev.synchronize()
