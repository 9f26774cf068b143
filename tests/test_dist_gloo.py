"""world_size-2 gloo run (CPU) of the plumbing of the corpus-sharded path: shard bounds, rank-major
all-gather of the per-shard top-k lists, and the merge semantics (checked here with a torch
restatement of qst_merge_topk, which itself is GPU-only and covered by the -m gpu tests)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _merge_reference(gv, gi, k):
    G, Q, _ = gv.shape
    flat_v = gv.permute(1, 0, 2).reshape(Q, -1)
    flat_i = gi.permute(1, 0, 2).reshape(Q, -1)
    key_i = torch.where(flat_i < 0, torch.full_like(flat_i, 2 ** 62), flat_i)
    order = torch.argsort(key_i, dim=1, stable=True)
    flat_v, flat_i = flat_v.gather(1, order), flat_i.gather(1, order)
    order = torch.argsort(flat_v, dim=1, descending=True, stable=True)[:, :k]
    return flat_v.gather(1, order), flat_i.gather(1, order)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from qst_b200 import sharded
        from oracle import ir_oracle
        g = torch.Generator().manual_seed(14)
        q = torch.randn(9, 16, generator=g)
        c = torch.randn(101, 16, generator=g)          # identical on both ranks
        k = 7
        s, e = sharded.shard_bounds(c.shape[0], world, rank)
        v, i = ir_oracle.topk_dense(q, c[s:e], k)       # this rank's exact local top-k (stands in for K2/K3)
        gv, gi = sharded.all_gather_topk(v, i + s)      # local row -> global id, then ONE exchange
        assert gv.shape == (world, 9, k) and gi.shape == (world, 9, k)
        assert torch.equal(gv[rank], v) and torch.equal(gi[rank], i + s)   # rank-major layout
        mv, mi = _merge_reference(gv, gi, k)
        wv, wi = ir_oracle.topk_dense(q, c, k)
        assert torch.equal(mi, wi), (rank, mi, wi)
        torch.testing.assert_close(mv, wv, rtol=0, atol=1e-6)
        # candidate exchange: one all-to-all routes every query's list to its owner, source-rank major
        q_own, m = 5, 3
        lists = torch.arange(world * q_own * (m + 1) * 2, dtype=torch.int32).view(world * q_own, m + 1, 2) + 1000 * rank
        recv = sharded.exchange_candidate_lists(lists)
        assert recv.shape == (world, q_own, m + 1, 2)
        for src in range(world):
            want = (torch.arange(world * q_own * (m + 1) * 2, dtype=torch.int32).view(world * q_own, m + 1, 2)
                    + 1000 * src)[rank * q_own:(rank + 1) * q_own]
            assert torch.equal(recv[src], want)
        assert sharded.candidates_per_shard(192, 8) == 96 and sharded.candidates_per_shard(192, 1) == 192
        # the communicator ShardedCorpus talks to (same layouts, torch.distributed underneath)
        from qst_b200 import comm
        cm = comm.default_comm()
        assert isinstance(cm, comm.TorchComm) and cm.world == world and cm.rank == rank
        x = torch.arange(6, dtype=torch.float32).view(3, 2) + 100 * rank
        ga = cm.all_gather(x)
        assert ga.shape == (world * 3, 2) and torch.equal(ga[rank * 3:(rank + 1) * 3], x)
        assert torch.equal(ga[(1 - rank) * 3:(2 - rank) * 3], x - 100 * rank + 100 * (1 - rank))
        y = torch.arange(world * 2, dtype=torch.int32).view(world * 2, 1) + 10 * rank      # block r -> rank r
        a2a = cm.all_to_all(y)
        for src in range(world):
            assert torch.equal(a2a[src * 2:(src + 1) * 2, 0], torch.arange(rank * 2, rank * 2 + 2, dtype=torch.int32) + 10 * src)
        assert torch.equal(cm.all_reduce_max(torch.tensor([float(rank), 5.0 - rank])), torch.tensor([world - 1.0, 5.0]))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shard_gather_merge(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
