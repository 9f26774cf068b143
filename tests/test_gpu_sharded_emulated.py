"""The corpus-sharded code path itself (``ShardedCorpus.topk`` / ``topk_owned``, both master modes and
the all-gather/merge strategy) on ONE GPU: the G ranks are G threads of this process, the collectives
go through ``qst_b200.comm.LocalComm`` (tensor copies behind a barrier), the peer-writable threshold
hints are plain allocations.  Everything else -- K1, K2 with shared thresholds, candidate lists, the
request / rescore / finalize kernels, the distributed exact re-scan -- is the code the multi-GPU run
executes.  Compared with the CPU oracle of the UNSHARDED corpus."""
import pytest
import torch

from test_gpu_scoring import _oracle_topk, assert_same_ranking

pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _run(G, q, c, k, score, master, strategy="owners", owned=False, query_tile=16384):
    """All ranks of an emulated node; returns rank-ordered (vals, idx, margin, rescanned)."""
    import qst_b200
    from qst_b200 import comm, sharded
    N = c.shape[0]
    dev = _dev()
    q_dev, c_dev = q.to(dev), c.to(dev)

    def body(cm):
        s, e = sharded.shard_bounds(N, cm.world, cm.rank)
        corp = sharded.ShardedCorpus(c_dev[s:e], N, score, comm=cm, query_tile=query_tile,
                                     full_master=c_dev if master == "replicated" else None)
        assert corp.master_mode == master and corp.world == G and corp.rank == cm.rank
        if owned:
            q_own = -(-q.shape[0] // G)
            qp = torch.cat([q_dev, q_dev[-1:].expand(q_own * G - q.shape[0], -1)])
            v, i, m = corp.topk_owned(qp[cm.rank * q_own:(cm.rank + 1) * q_own], k)
        else:
            v, i, m = corp.topk(q_dev, k, strategy=strategy)
        torch.cuda.synchronize()
        if G > 1 and strategy == "owners":
            assert corp._arena is not None, "the exchanges must have been fused into their producer kernels"
        return v.cpu(), i.cpu(), m.cpu(), corp.last_rescanned

    return comm.run_local_world(G, body)


@pytest.mark.parametrize("score", ["cos_sim", "dot_score", "euclid_score"])
@pytest.mark.parametrize("G,master", [(1, "sharded"), (2, "sharded"), (4, "sharded"), (3, "sharded"), (4, "replicated")])
def test_sharded_topk_matches_unsharded_oracle(G, master, score):
    g = torch.Generator().manual_seed(100 + G)
    Q, N, D, k = 150, 30011, 96, 50
    q = torch.randn(Q, D, generator=g)
    c = torch.randn(N, D, generator=g) * (1 + torch.rand(N, 1, generator=g))
    want_val, want_idx = _oracle_topk(q, c, k, score)
    out = _run(G, q, c, k, score, master)
    for rank, (v, i, m, _) in enumerate(out):
        what = f"G={G} {master} {score} rank {rank}"
        assert v.shape == (Q, k) and i.shape == (Q, k)
        assert bool((m > 0).all()), what
        if score == "dot_score":
            scale = float(want_val.abs().max())
            assert_same_ranking(i, v / scale, want_idx, want_val / scale, what)
        else:
            assert_same_ranking(i, v, want_idx, want_val, what, truth=(q, c, score) if score == "euclid_score" else None)
    # every rank holds the same answer
    for v, i, m, _ in out[1:]:
        assert torch.equal(v, out[0][0]) and torch.equal(i, out[0][1])


def test_owned_slices_and_allgather_merge_strategy():
    """``topk_owned`` returns exactly the owner's slice; the literal local-top-k / all-gather / merge
    strategy gives the same ranking."""
    g = torch.Generator().manual_seed(7)
    Q, N, D, k, G = 101, 20000, 64, 20, 4
    q = torch.randn(Q, D, generator=g)
    c = torch.randn(N, D, generator=g)
    want_val, want_idx = _oracle_topk(q, c, k, "cos_sim")
    q_own = -(-Q // G)
    for master in ("sharded", "replicated"):
        out = _run(G, q, c, k, "cos_sim", master, owned=True)
        for rank, (v, i, m, _) in enumerate(out):
            lo, hi = rank * q_own, min((rank + 1) * q_own, Q)
            assert v.shape == (q_own, k)
            assert_same_ranking(i[:hi - lo], v[:hi - lo], want_idx[lo:hi], want_val[lo:hi], f"{master} owned rank {rank}")
            assert bool((m > 0).all())
    out = _run(G, q, c, k, "cos_sim", "sharded", strategy="allgather_merge", query_tile=40)
    for rank, (v, i, m, _) in enumerate(out):
        assert_same_ranking(i, v, want_idx, want_val, f"allgather_merge rank {rank}")
        assert bool((m > 0).all())


@pytest.mark.parametrize("master", ["sharded", "replicated"])
def test_sharded_near_ties_go_through_the_exact_rescan(master):
    """Clustered corpus (centroid + tiny noise): the bf16 pass cannot separate neighbours, certificates
    fail, and the exact re-scan -- distributed over the shards when the master is sharded -- must
    still deliver the oracle's ranking."""
    g = torch.Generator().manual_seed(3)
    Q, N, D, k, G = 40, 6000, 64, 10, 3
    cent = torch.randn(30, D, generator=g)
    c = cent[torch.randint(0, 30, (N,), generator=g)] + 1e-4 * torch.randn(N, D, generator=g)
    q = cent[torch.randint(0, 30, (Q,), generator=g)] + 1e-3 * torch.randn(Q, D, generator=g)
    want_val, want_idx = _oracle_topk(q, c, k, "cos_sim")
    out = _run(G, q, c, k, "cos_sim", master)
    for rank, (v, i, m, rescanned) in enumerate(out):
        assert_same_ranking(i, v, want_idx, want_val, f"near ties {master} rank {rank}")
        assert bool((m > 0).all())
    if master == "sharded":
        assert out[0][3] > 0, "this data is meant to fail certificates and exercise the distributed re-scan"


def test_deferred_certificate_check_patches_results_in_place():
    """``exact="deferred"``: the call returns without a host read; the next call (or ``finish_exact``)
    re-scans what was left uncertified and patches the tensors of the earlier call in place."""
    import qst_b200
    from qst_b200 import comm, sharded
    g = torch.Generator().manual_seed(3)
    Q, N, D, k, G = 40, 6000, 64, 10, 2
    cent = torch.randn(30, D, generator=g)
    c = cent[torch.randint(0, 30, (N,), generator=g)] + 1e-4 * torch.randn(N, D, generator=g)
    q = cent[torch.randint(0, 30, (Q,), generator=g)] + 1e-3 * torch.randn(Q, D, generator=g)
    want_val, want_idx = _oracle_topk(q, c, k, "cos_sim")
    dev = _dev()
    q_dev, c_dev = q.to(dev), c.to(dev)
    q_own = Q // G

    def body(cm):
        s, e = sharded.shard_bounds(N, cm.world, cm.rank)
        corp = sharded.ShardedCorpus(c_dev[s:e], N, "cos_sim", comm=cm)
        own = q_dev[cm.rank * q_own:(cm.rank + 1) * q_own]
        v1, i1, m1 = corp.topk_owned(own, k, exact="deferred")
        torch.cuda.synchronize()
        flagged_before = int((~(m1 > 0)).sum())
        v2, i2, m2 = corp.topk_owned(own, k, exact="deferred")      # checks + repairs the first call
        torch.cuda.synchronize()
        first_ok = bool((m1 > 0).all())
        n2 = corp.finish_exact()                                    # ... and this one the second
        torch.cuda.synchronize()
        return v1.cpu(), i1.cpu(), m1.cpu(), v2.cpu(), i2.cpu(), m2.cpu(), flagged_before, first_ok, n2

    out = comm.run_local_world(G, body)
    assert sum(o[6] for o in out) > 0, "the data is meant to leave queries uncertified after the first pass"
    for rank, (v1, i1, m1, v2, i2, m2, _, first_ok, n2) in enumerate(out):
        lo, hi = rank * q_own, (rank + 1) * q_own
        assert first_ok and n2 > 0
        for v, i, m in ((v1, i1, m1), (v2, i2, m2)):
            assert_same_ranking(i, v, want_idx[lo:hi], want_val[lo:hi], f"deferred rank {rank}")
            assert bool((m > 0).all())


@pytest.mark.parametrize("master", ["sharded", "replicated"])
def test_sharded_all_zero_query_and_duplicated_documents(master):
    """An all-zero query scores 0 against every document of every shard (a corpus-wide exact tie, margin
    exactly 0 in the first pass) and duplicated documents straddle the shard boundary: the distributed
    re-scan must settle both towards the lower global position."""
    g = torch.Generator().manual_seed(11)
    Q, N, D, k, G = 12, 900, 48, 10, 3
    q = torch.randn(Q, D, generator=g)
    c = torch.randn(N, D, generator=g)
    q[4] = 0
    c[299] = c[300] = c[650]            # shard boundaries at 300 and 600
    q[5] = c[300] + 0.01 * torch.randn(D, generator=g)
    want_val, _ = _oracle_topk(q, c, k, "cos_sim")
    out = _run(G, q, c, k, "cos_sim", master)
    for rank, (v, i, m, _) in enumerate(out):
        torch.testing.assert_close(v, want_val, rtol=0, atol=2e-6)
        assert bool((m > 0).all()), f"{master} rank {rank}: {m}"
        assert i[4].tolist() == list(range(k)), "all scores tie at 0: the k lowest positions, in order"
        assert i[5, :3].tolist() == [299, 300, 650]


@pytest.mark.parametrize("exact", [True, "deferred"])
def test_prefetched_batches_give_the_same_rankings(exact):
    """``topk_owned(..., prefetch=next)``: K1 and the distribution of the next batch happen during the
    current call; the next call starts from the gathered buffers.  A stream of three batches, then a
    prefetched batch that is NOT the next one asked for (discarded)."""
    import qst_b200
    from qst_b200 import comm, sharded
    g = torch.Generator().manual_seed(21)
    N, D, k, G, q_own = 20011, 96, 20, 3, 30
    c = torch.randn(N, D, generator=g)
    batches = [torch.randn(G * q_own, D, generator=g) for _ in range(4)]
    want = [_oracle_topk(b, c, k, "cos_sim") for b in batches]
    dev = _dev()
    c_dev = c.to(dev)
    b_dev = [b.to(dev) for b in batches]

    def body(cm):
        s, e = sharded.shard_bounds(N, cm.world, cm.rank)
        corp = sharded.ShardedCorpus(c_dev[s:e], N, "cos_sim", comm=cm)
        own = [b[cm.rank * q_own:(cm.rank + 1) * q_own] for b in b_dev]
        out = []
        out.append(corp.topk_owned(own[0], k, exact=exact, prefetch=own[1]))
        used = corp._prefetched is not None
        out.append(corp.topk_owned(own[1], k, exact=exact, prefetch=own[2]))
        out.append(corp.topk_owned(own[2], k, exact=exact, prefetch=own[0]))     # announced: batch 0 ...
        out.append(corp.topk_owned(own[3], k, exact=exact))                       # ... asked for: batch 3
        corp.finish_exact()
        torch.cuda.synchronize()
        return [(v.cpu(), i.cpu(), m.cpu()) for v, i, m in out], used

    res = comm.run_local_world(G, body)
    for rank, (out, used) in enumerate(res):
        assert used, "peer-mapped buffers exist on an emulated node: the prefetch path must have been taken"
        for b, (v, i, m) in enumerate(out):
            lo, hi = rank * q_own, (rank + 1) * q_own
            assert_same_ranking(i, v, want[b][1][lo:hi], want[b][0][lo:hi], f"batch {b} rank {rank} exact={exact}")
            assert bool((m > 0).all())


@pytest.mark.parametrize("lookahead", [False, True])
def test_sharded_host_pipeline_with_and_without_lookahead(lookahead):
    """``ShardedHostPipeline``: pinned host slices in, pinned rankings out, double-buffered; with
    ``submit(cur, next)`` the next batch is copied in and distributed one step ahead."""
    import qst_b200
    from qst_b200 import comm, sharded
    g = torch.Generator().manual_seed(33)
    N, D, k, G, q_own = 15013, 64, 10, 2, 40
    c = torch.randn(N, D, generator=g)
    batches = [torch.randn(G * q_own, D, generator=g) for _ in range(5)]
    want = [_oracle_topk(b, c, k, "cos_sim") for b in batches]
    c_dev = c.to(_dev())

    def body(cm):
        s, e = sharded.shard_bounds(N, cm.world, cm.rank)
        corp = sharded.ShardedCorpus(c_dev[s:e], N, "cos_sim", comm=cm)
        own = [b[cm.rank * q_own:(cm.rank + 1) * q_own].contiguous().pin_memory() for b in batches]
        pipe = sharded.ShardedHostPipeline(corp, k)
        out, tickets = [], []
        for t in range(len(own)):
            nxt = own[t + 1] if (lookahead and t + 1 < len(own)) else None
            tickets.append(pipe.submit(own[t], nxt) if nxt is not None else pipe.submit(own[t]))
            if t >= 1:
                v, i = pipe.result(tickets[t - 1])
                out.append((v.clone(), i.clone()))
        v, i = pipe.result(tickets[-1])
        out.append((v.clone(), i.clone()))
        pipe.drain()
        return out

    res = comm.run_local_world(G, body)
    for rank, out in enumerate(res):
        lo, hi = rank * q_own, (rank + 1) * q_own
        for b, (v, i) in enumerate(out):
            assert_same_ranking(i, v, want[b][1][lo:hi], want[b][0][lo:hi], f"pipeline batch {b} rank {rank}")


def test_peer_barrier_and_scatter_descriptor_through_the_c_abi():
    """``qst_peer_barrier`` itself (the emulated ranks above meet on the host instead): two "ranks" = two
    streams of this process with one flag array each; each barrier completes only when both have signalled.
    And ``qst_rescore_requests_scatter`` against the plain call + the all-to-all it replaces."""
    import ctypes as C
    import qst_b200
    from qst_b200 import _lib
    lib = _lib.load()
    dev = _dev()
    flags = [torch.zeros(_lib.QST_MAX_WORLD, dtype=torch.int32, device=dev) for _ in range(2)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    import time
    torch.cuda.synchronize()

    def launch(r, epoch):
        sc = _lib.Scatter()
        sc.base[0], sc.base[1] = flags[0].data_ptr(), flags[1].data_ptr()
        sc.world, sc.rank, sc.rows_per_block = 2, r, 1
        _lib.check(lib.qst_peer_barrier(C.byref(sc), epoch, streams[r].cuda_stream))    # no torch call in between:

    for epoch in (1, 2, 3):                                    # anything that drains stream 0 would wait for ever
        launch(0, epoch)
        time.sleep(0.05)
        assert not streams[0].query(), "rank 0 must still be waiting: rank 1 has not signalled this epoch"
        launch(1, epoch)
        streams[0].synchronize()
        streams[1].synchronize()
    assert flags[0][:2].tolist() == [3, 3] and flags[1][:2].tolist() == [3, 3]
    launch(1, 3)                                               # a rank already signalled for: returns at once
    streams[1].synchronize()

    # scatter: G = 2 ranks' rescoring outputs land where an all-to-all would have put them
    g = torch.Generator().manual_seed(2)
    G, q_own, m, D, n = 2, 37, 24, 64, 500
    c = torch.randn(n, D, generator=g).to(dev)
    q_all = torch.randn(G * q_own, D, generator=g).to(dev)
    recv = [torch.full((G * q_own, m), float("nan"), device=dev) for _ in range(G)]
    want = [torch.empty((G * q_own, m), device=dev) for _ in range(G)]
    for r in range(G):
        req = torch.randint(-1, n, (G * q_own, m), generator=g, dtype=torch.int32).to(dev)
        _lib.check(lib.qst_rescore_requests(G * q_own, m, D, _lib.QST_SCORE_DOT, req.data_ptr(), q_all.data_ptr(), None,
                                            c.data_ptr(), None, want[r].data_ptr(), _lib.stream_ptr(dev)))
        sc = _lib.Scatter()
        for o in range(G):
            sc.base[o] = recv[o].data_ptr()
        sc.world, sc.rank, sc.rows_per_block = G, r, q_own
        staging = torch.empty((G * q_own, m), device=dev)
        _lib.check(lib.qst_rescore_requests_scatter(G * q_own, m, D, _lib.QST_SCORE_DOT, req.data_ptr(), q_all.data_ptr(),
                                                    None, c.data_ptr(), None, staging.data_ptr(), C.byref(sc),
                                                    _lib.stream_ptr(dev)))
    torch.cuda.synchronize()
    for o in range(G):       # owner o received block o of every rank r, at block position r
        for r in range(G):
            assert torch.equal(recv[o][r * q_own:(r + 1) * q_own], want[r][o * q_own:(o + 1) * q_own])
    # a descriptor that does not cover the rows is refused
    sc.rows_per_block = q_own + 1
    assert lib.qst_rescore_requests_scatter(G * q_own, m, D, _lib.QST_SCORE_DOT, req.data_ptr(), q_all.data_ptr(), None,
                                            c.data_ptr(), None, staging.data_ptr(), C.byref(sc), _lib.stream_ptr(dev)) != 0


def test_deferred_rescan_of_a_prefetched_stream_happens_under_the_next_k2():
    """Near-tie data (certificates fail), ``exact="deferred"`` and ``prefetch=`` together: the re-scan of
    batch t runs inside call t+1 AFTER that call's K2 was queued and after its prefetch has pushed batch
    t+2 -- the gathered queries of batch t must still be intact (third generation of the gather buffers)."""
    import qst_b200
    from qst_b200 import comm, sharded
    g = torch.Generator().manual_seed(3)
    N, D, k, G, q_own = 6000, 64, 10, 2, 20
    cent = torch.randn(30, D, generator=g)
    c = cent[torch.randint(0, 30, (N,), generator=g)] + 1e-4 * torch.randn(N, D, generator=g)
    batches = [cent[torch.randint(0, 30, (G * q_own,), generator=g)] + 1e-3 * torch.randn(G * q_own, D, generator=g)
               for _ in range(4)]
    want = [_oracle_topk(b, c, k, "cos_sim") for b in batches]
    dev = _dev()
    c_dev = c.to(dev)
    b_dev = [b.to(dev) for b in batches]

    def body(cm):
        s, e = sharded.shard_bounds(N, cm.world, cm.rank)
        corp = sharded.ShardedCorpus(c_dev[s:e], N, "cos_sim", comm=cm)
        own = [b[cm.rank * q_own:(cm.rank + 1) * q_own] for b in b_dev]
        out, rescanned = [], 0
        for t in range(4):
            out.append(corp.topk_owned(own[t], k, exact="deferred", prefetch=own[t + 1] if t < 3 else None))
            rescanned += corp.last_rescanned
        rescanned += corp.finish_exact()
        torch.cuda.synchronize()
        return [(v.cpu(), i.cpu(), m.cpu()) for v, i, m in out], rescanned

    res = comm.run_local_world(G, body)
    for rank, (out, rescanned) in enumerate(res):
        assert rescanned > 0, "the data is meant to leave queries uncertified after the first pass"
        for b, (v, i, m) in enumerate(out):
            lo, hi = rank * q_own, (rank + 1) * q_own
            assert_same_ranking(i, v, want[b][1][lo:hi], want[b][0][lo:hi], f"batch {b} rank {rank}")
            assert bool((m > 0).all())
