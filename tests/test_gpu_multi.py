"""Real multi-GPU run of the corpus-sharded path (NCCL): sharded and replicated fp32 master, the
owner and the all-gather/merge strategies, against the CPU oracle.  Needs >= 2 visible GPUs (`gpurun --gpus 2`), skipped otherwise."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import qst_b200
        from qst_b200 import sharded
        from oracle import ir_oracle
        g = torch.Generator().manual_seed(14)
        Q, N, D, k = 333, 40_003, 128, 100
        q = torch.randn(Q, D, generator=g)
        c = torch.randn(N, D, generator=g)
        want_val, want_idx = ir_oracle.topk_dense(q, c, k)
        s, e = sharded.shard_bounds(N, world, rank)
        for full in (None, c.to(dev)):
            corp = sharded.ShardedCorpus(c[s:e].to(dev), N, "cos_sim", query_tile=128, full_master=full)
            assert corp.master_mode == ("sharded" if full is None else "replicated")
            strategies = ("owners", "allgather_merge") if full is None else ("owners",)
            for strategy in strategies:
                vals, idx, margin = corp.topk(q.to(dev), k, strategy=strategy)
                torch.cuda.synchronize()
                assert vals.shape == (Q, k) and bool((margin > 0).all())
                torch.testing.assert_close(vals.cpu(), want_val, rtol=0, atol=2e-6)
                mism = (idx.cpu() != want_idx)
                # only tie swaps (scores within 1e-6) may differ
                assert float((vals.cpu()[mism] - want_val[mism]).abs().max() if mism.any() else 0.0) <= 2e-6
                assert mism.float().mean() < 1e-3
            # owner-sliced entry: every rank passes its slice, gets its slice of the answer
            q_own = -(-Q // world)
            qp = torch.cat([q, q[-1:].expand(q_own * world - Q, -1)])
            lo, hi = rank * q_own, (rank + 1) * q_own
            v2, i2, m2 = corp.topk_owned(qp[lo:hi].to(dev), k)
            hi_real = min(hi, Q)
            torch.testing.assert_close(v2.cpu()[: hi_real - lo], want_val[lo:hi_real], rtol=0, atol=2e-6)
            assert (i2.cpu()[: hi_real - lo] != want_idx[lo:hi_real]).float().mean() < 1e-3
            assert bool((m2 > 0).all())
            assert corp._arena is not None, "one node: the exchanges must run fused over peer memory"
            if full is None:
                # a stream of batches, each announced one call ahead (K1 + copy-engine distribution of
                # batch t+1 underneath the exchanges of batch t): same rankings, and the path was taken
                own = qp[lo:hi].to(dev)
                for t in range(3):
                    v3, i3, m3 = corp.topk_owned(own, k, exact="deferred", prefetch=own if t < 2 else None)
                    assert (corp._prefetched is not None) == (t < 2), "peer-mapped buffers must exist on one node"
                    corp.finish_exact()
                    assert torch.equal(v3, v2) and torch.equal(i3, i2) and bool((m3 > 0).all())
        # the same retrieval with NCCL called through the C ABI (qst_comm_*), no torch.distributed on the data path
        from qst_b200 import comm
        nc = comm.NcclComm.from_torch(dev)
        x = torch.arange(6, dtype=torch.float32, device=dev).view(3, 2) + 100 * rank
        assert torch.equal(nc.all_gather(x), comm.TorchComm().all_gather(x))
        y = torch.arange(world * 4, dtype=torch.int32, device=dev).view(world * 2, 2) + 10 * rank
        assert torch.equal(nc.all_to_all(y), comm.TorchComm().all_to_all(y))
        assert torch.equal(nc.all_reduce_max(torch.tensor([float(rank), 3.0 - rank], device=dev)).cpu(),
                           torch.tensor([world - 1.0, 3.0]))
        corp = sharded.ShardedCorpus(c[s:e].to(dev), N, "cos_sim", comm=nc)
        vals, idx, margin = corp.topk(q.to(dev), k)
        torch.cuda.synchronize()
        torch.testing.assert_close(vals.cpu(), want_val, rtol=0, atol=2e-6)
        assert (idx.cpu() != want_idx).float().mean() < 1e-3 and bool((margin > 0).all())
        nc.close()
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_two_gpu_sharded_retrieval_matches_oracle(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
