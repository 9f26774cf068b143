"""Collectives of the corpus-sharded path behind one small interface.

The sharded retrieval needs four things from "the other ranks": an all-gather, an all-to-all (equal
blocks), a MAX all-reduce of a few floats, and device buffers every rank of the node can write
(threshold hints pushed with remote atomics while K2 runs).  Three implementations:

* ``TorchComm``  -- ``torch.distributed`` (backend ``nccl`` on GPUs over NVLink/NVSwitch, ``gloo`` in the
  CPU tests of the plumbing); peer-writable buffers through CUDA IPC.
* ``NcclComm``   -- NCCL called directly through the C ABI (``qst_comm_*`` in ``include/qst.h``), no
  ``torch.distributed`` on the data path; the unique id is exchanged by the caller.
* ``LocalComm``  -- G ranks as G THREADS of one process on one GPU.  Exists so that the sharded code path
  (not a re-statement of it) runs in the single-GPU test tier: collectives are tensor copies behind a
  ``threading.Barrier``, "peer" buffers are plain allocations every thread can address.

All tensor collectives are stream-ordered on the caller's current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import List, Optional, Sequence

import torch

from . import _lib


class SharedBuffers:
    """One zero-filled device buffer per rank, writable by every rank: ``local`` is this rank's own,
    ``peers`` the other ranks' (mapped) pointers in rank order."""

    def __init__(self, local: int, peers: Sequence[int], closer=None, keep=None):
        self.local, self.peers, self._closer, self._keep = int(local), [int(p) for p in peers], closer, keep

    def close(self):
        if self._closer is not None:
            self._closer()
            self._closer = None
        self._keep = None


class Comm:
    world: int
    rank: int

    def all_gather(self, t: torch.Tensor) -> torch.Tensor:
        """[n, ...] per rank -> [world * n, ...] (rank-major) on every rank."""
        raise NotImplementedError

    def all_to_all(self, t: torch.Tensor) -> torch.Tensor:
        """[world * n, ...] (block r goes to rank r) -> [world * n, ...] (block r came from rank r)."""
        raise NotImplementedError

    def all_reduce_max(self, t: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def barrier(self) -> None:
        raise NotImplementedError

    def shared_buffers(self, nbytes: int, device: torch.device) -> Optional[SharedBuffers]:
        """Collective.  None when peer-writable memory is not available (the caller then does without)."""
        return None

    def scatter_descriptor(self, buffers: SharedBuffers, offset: int, rows_per_block: int) -> "_lib.Scatter":
        """``qst_scatter`` over `buffers` (+ byte offset) of every rank, in rank order."""
        sc = _lib.Scatter()
        others = iter(buffers.peers)
        for r in range(self.world):
            sc.base[r] = (buffers.local if r == self.rank else next(others)) + offset
        sc.world, sc.rank, sc.rows_per_block = self.world, self.rank, rows_per_block
        return sc

    def peer_barrier(self, flags: "_lib.Scatter", epoch: int, device: torch.device) -> None:
        """Barrier of the node's ranks over peer memory, ordered on the current stream
        (``qst_peer_barrier``): everything a rank's stream wrote before it -- into ANY rank's buffers --
        is visible to every rank's stream after it."""
        with torch.cuda.device(device):
            _lib.check(_lib.load().qst_peer_barrier(C.byref(flags), epoch, torch.cuda.current_stream(device).cuda_stream))


class SingleComm(Comm):
    """World of one: every collective is the identity."""
    world, rank = 1, 0

    def all_gather(self, t):
        return t

    def all_to_all(self, t):
        return t

    def all_reduce_max(self, t):
        return t

    def barrier(self):
        pass


class TorchComm(Comm):
    def __init__(self, group=None):
        import torch.distributed as dist
        self._dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)

    def all_gather(self, t):
        t = t.contiguous()
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        self._dist.all_gather_into_tensor(out, t, group=self.group)
        return out

    def all_to_all(self, t):
        t = t.contiguous()
        out = torch.empty_like(t)
        self._dist.all_to_all_single(out, t, group=self.group)
        return out

    def all_reduce_max(self, t):
        t = t.clone()
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX, group=self.group)
        return t

    def barrier(self):
        self._dist.barrier(self.group)

    def shared_buffers(self, nbytes, device):
        return _ipc_shared_buffers(self, nbytes, device)


def _ipc_shared_buffers(cm: "Comm", nbytes: int, device) -> Optional[SharedBuffers]:
    """cudaMalloc + CUDA IPC handles exchanged through ``cm``'s own all-gather; all ranks or none."""
    if torch.device(device).type != "cuda":
        return None
    lib = _lib.load()
    local, peers, ok = C.c_void_p(), [], 1
    handle = C.create_string_buffer(64)
    with torch.cuda.device(device):
        if lib.qst_peer_buffer_create(nbytes, C.byref(local), handle) != 0:
            ok, local = 0, C.c_void_p()
        mine = torch.tensor(list(handle.raw), dtype=torch.uint8, device=device)
        every = cm.all_gather(mine).cpu().view(cm.world, 64)
        if ok:
            for r in range(cm.world):
                if r == cm.rank:
                    continue
                p = C.c_void_p()
                if lib.qst_peer_buffer_open(bytes(every[r].tolist()), C.byref(p)) != 0:
                    ok = 0
                    break
                peers.append(p)
        failed = cm.all_reduce_max(torch.tensor([1.0 - ok], dtype=torch.float32, device=device))

    def closer():
        for p in peers:
            lib.qst_peer_buffer_close(p)
        if local:
            lib.qst_peer_buffer_destroy(local)

    if float(failed) > 0:
        closer()
        return None
    return SharedBuffers(local.value, [p.value for p in peers], closer)


class NcclComm(Comm):
    """NCCL through the C ABI (``qst_comm_*``): no ``torch.distributed`` on the data path.  The 128-byte
    unique id is made on one rank (``NcclComm.unique_id()``) and handed to all of them by the caller;
    ``NcclComm.from_torch`` borrows an initialised ``torch.distributed`` group just for that hand-over."""

    def __init__(self, world: int, rank: int, unique_id: bytes, device: torch.device):
        lib = _lib.load()
        if not lib.qst_comm_available():
            raise _lib.QstError("libnccl.so.2 could not be loaded: the C-ABI communicator is unavailable")
        if len(unique_id) != 128:
            raise ValueError("unique_id must be the 128 bytes of NcclComm.unique_id()")
        self.world, self.rank, self.device = world, rank, torch.device(device)
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(lib.qst_comm_init(bytes(unique_id), world, rank, C.byref(self._h)))

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _lib.check(_lib.load().qst_comm_unique_id(buf))
        return buf.raw

    @classmethod
    def from_torch(cls, device: torch.device, group=None) -> "NcclComm":
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        box = [cls.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls(world, rank, box[0], device)

    def _st(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def all_gather(self, t):
        t = t.contiguous()
        out = torch.empty((self.world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().qst_comm_allgather(self._h, t.data_ptr(), out.data_ptr(), t.numel() * t.element_size(),
                                                      self._st()))
        return out

    def all_to_all(self, t):
        t = t.contiguous()
        out = torch.empty_like(t)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().qst_comm_alltoall(self._h, t.data_ptr(), out.data_ptr(),
                                                     t.numel() * t.element_size() // self.world, self._st()))
        return out

    def all_reduce_max(self, t):
        src = t.to(torch.float32).contiguous()
        out = torch.empty_like(src)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().qst_comm_allreduce_max_f32(self._h, src.data_ptr(), out.data_ptr(), src.numel(),
                                                              self._st()))
        return out.to(t.dtype)

    def barrier(self):
        self.all_reduce_max(torch.zeros(1, device=self.device))
        torch.cuda.current_stream(self.device).synchronize()

    def shared_buffers(self, nbytes, device):
        return _ipc_shared_buffers(self, nbytes, device)

    def close(self):
        if self._h:
            _lib.load().qst_comm_destroy(self._h)
            self._h = C.c_void_p()


class _LocalWorld:
    """State shared by the G ``LocalComm`` threads of one emulated node."""

    def __init__(self, world: int):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots: List[object] = [None] * world
        self.failed = threading.Event()


class LocalComm(Comm):
    """Rank of a node emulated by threads of ONE process on ONE device.  A collective is "deposit,
    host barrier, read, host barrier"; the depositor's stream is drained before the deposit and the
    reader's stream after the read, so the threads may work on any streams (the double-buffered host
    pipeline uses one per slot).  A test shim: correctness of the sharded protocol, not speed."""

    def __init__(self, shared: _LocalWorld, rank: int):
        self._w, self.world, self.rank = shared, shared.world, rank

    @staticmethod
    def make_world(world: int) -> List["LocalComm"]:
        shared = _LocalWorld(world)
        return [LocalComm(shared, r) for r in range(world)]

    @staticmethod
    def _drain(obj):
        if isinstance(obj, torch.Tensor) and obj.is_cuda:
            torch.cuda.current_stream(obj.device).synchronize()

    def _exchange(self, obj, read=list):
        w = self._w
        self._drain(obj)          # what is deposited (and every peer write queued before it) is complete
        w.slots[self.rank] = obj
        w.barrier.wait()
        out = read(list(w.slots))
        self._drain(obj)          # the reads have happened before a depositor may reuse its memory
        w.barrier.wait()          # nobody overwrites a slot before everybody has read it
        return out

    def all_gather(self, t):
        return self._exchange(t.contiguous(), lambda got: torch.cat(got))

    def all_to_all(self, t):
        t = t.contiguous()
        n = t.shape[0] // self.world
        return self._exchange(t, lambda got: torch.cat([x[self.rank * n:(self.rank + 1) * n] for x in got]))

    def all_reduce_max(self, t):
        return self._exchange(t, lambda got: torch.stack(got).amax(dim=0))

    def barrier(self):
        self._w.barrier.wait()

    def peer_barrier(self, flags, epoch, device):
        # the "ranks" share one device (and often one stream): a spinning kernel would wait for work
        # queued behind it -- drain this thread's stream and meet on the host instead
        torch.cuda.current_stream(device).synchronize()
        self._w.barrier.wait()

    def shared_buffers(self, nbytes, device):
        buf = torch.zeros(max(nbytes, 4), dtype=torch.uint8, device=device)
        ptrs = self._exchange(buf.data_ptr())
        return SharedBuffers(ptrs[self.rank], [p for r, p in enumerate(ptrs) if r != self.rank], None, keep=buf)


def run_local_world(world: int, fn, *args):
    """Runs ``fn(comm, *args)`` on ``world`` threads (one ``LocalComm`` each); returns the results in
    rank order, re-raising the first exception of any rank."""
    comms = LocalComm.make_world(world)
    out: List[object] = [None] * world
    err: List[BaseException] = []

    def body(r):
        try:
            out[r] = fn(comms[r], *args)
        except BaseException as e:   # noqa: BLE001 -- re-raised below
            err.append(e)
            comms[r]._w.barrier.abort()
        finally:
            from . import scoring
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            scoring.release_workspaces()     # this thread's scratch buffers die with the thread

    threads = [threading.Thread(target=body, args=(r,), daemon=True) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if err:
        real = [e for e in err if not isinstance(e, threading.BrokenBarrierError)]
        raise (real or err)[0]
    return out


def default_comm(group=None) -> Comm:
    """``TorchComm`` over ``group`` when ``torch.distributed`` is initialised, else a world of one."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return TorchComm(group)
    return SingleComm()
