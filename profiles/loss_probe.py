"""Probe for config 2: what does a plain 50 MB -> 50 MB device copy of rotating buffers cost per
launch (the practical HBM floor for a 100 MB kernel), next to the fused loss kernel."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qst_b200  # noqa: E402
from qst_b200 import _lib, quad_loss  # noqa: E402

dev = torch.device("cuda:0")
B, D, sets = 4096, 768, 12
lib = _lib.load()
g = torch.Generator(device=dev).manual_seed(1)
data = [[torch.randn(B, D, generator=g, device=dev) for _ in range(4)] for _ in range(sets)]
grads = [[torch.empty(B, D, device=dev) for _ in range(4)] for _ in range(sets)]
src = [torch.randn(4 * B * D, generator=g, device=dev) for _ in range(sets)]
dst = [torch.empty(4 * B * D, device=dev) for _ in range(sets)]
loss = torch.empty(sets * B, device=dev)
ws = torch.zeros(lib.qst_quadruplet_workspace_bytes(), dtype=torch.uint8, device=dev)
prm = quad_loss._params(0.6, 1.0, 0.5, 0.5, 2.0, False)


def fused(st, red, upstream=1.0):
    for i in range(sets):
        x, gr = data[i], grads[i]
        _lib.check(lib.qst_quadruplet_fwd_bwd(x[0].data_ptr(), x[1].data_ptr(), x[2].data_ptr(), x[3].data_ptr(),
                                              _lib.QST_F32, B, D, C.byref(prm), red, upstream, loss[i * B:].data_ptr(),
                                              gr[0].data_ptr(), gr[1].data_ptr(), gr[2].data_ptr(),
                                              gr[3].data_ptr(), ws.data_ptr(), st))


def copies(st):
    for i in range(sets):
        dst[i].copy_(src[i])


def timeit(name, fn):
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        fn(side.cuda_stream)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            fn(torch.cuda.current_stream(dev).cuda_stream)
    torch.cuda.current_stream(dev).wait_stream(side)
    import pynvml
    blocks = []
    for blk in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            graph.replay()
        b.record()
        torch.cuda.synchronize()
        mhz = pynvml.nvmlDeviceGetClockInfo(NV, pynvml.NVML_CLOCK_SM)
        blocks.append((a.elapsed_time(b) * 1e3 / (20 * sets), mhz))
    us = sorted(x for x, _ in blocks)[len(blocks) // 2]
    print(f"{name:34s} median {us:7.2f} us/launch  {8 * B * D * 4 / us / 1e3:7.0f} GB/s   blocks: "
          + " ".join(f"{x:.2f}@{m}" for x, m in blocks))


import pynvml
pynvml.nvmlInit()
NV = pynvml.nvmlDeviceGetHandleByIndex(0)
timeit("torch copy 50MB->50MB", copies)
# QST_LOSS_REDUCE is read at every call, i.e. when the graph is captured
modes = os.environ.get("QST_PROBE_MODES", "default").split(",")
for rep in range(2):
    for mode in modes:
        if mode == "default":
            os.environ.pop("QST_LOSS_REDUCE", None)
        else:
            os.environ["QST_LOSS_REDUCE"] = mode
        timeit(f"fused loss mean [{mode}]", lambda st: fused(st, _lib.QST_RED_MEAN))
    timeit("fused loss none (no reduction)", lambda st: fused(st, _lib.QST_RED_NONE))
    if os.environ.get("QST_PROBE_EXTRA"):   # is it the reduction or the VALUES of the gradients (x 1/B under 'mean')?
        timeit("fused loss sum", lambda st: fused(st, _lib.QST_RED_SUM))
        timeit("fused loss mean, upstream = B", lambda st: fused(st, _lib.QST_RED_MEAN, float(B)))
        timeit("fused loss none, upstream = 1/B", lambda st: fused(st, _lib.QST_RED_NONE, 1.0 / B))
