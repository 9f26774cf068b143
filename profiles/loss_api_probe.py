"""Wall time per training step of the loss through the Python API (forward + backward through autograd,
gradients reset to None between steps as `optimizer.zero_grad()` does), config 2 (4096 x 768 fp32):

  floor      a custom autograd Function that launches nothing (what torch's engine costs by itself)
  ours       qst_b200.GammaQuadrupletLoss (one fused kernel in forward, one scale in backward)
  torch_ops  the reference's formulation: three F.triplet_margin_loss calls (models/losses/losses.py:35-69)
"""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qst_b200  # noqa: E402

dev = torch.device("cuda:0")
B, D = 4096, 768
g = torch.Generator(device=dev).manual_seed(14)
xs = [torch.randn(B, D, generator=g, device=dev).requires_grad_(True) for _ in range(4)]


class Floor(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, p, q, n, buf, loss):
        ctx.buf = buf
        return loss

    @staticmethod
    def backward(ctx, go):
        b = ctx.buf.unbind(0)
        return b[0], b[1], b[2], b[3], None, None


buf = torch.zeros(4, B, D, device=dev)
loss0 = torch.zeros((), device=dev)
mod = qst_b200.GammaQuadrupletLoss(gamma=0.6, margin_pos_neg=1.0, margin_pos_part=0.5, margin_part_neg=0.5)


def floor():
    return Floor.apply(*xs, buf, loss0)


def ours():
    return mod(x_anchor=xs[0], x_pos=xs[1], x_part=xs[2], x_neg=xs[3])


def torch_ops():
    a, p, q, n = xs
    return (F.triplet_margin_loss(a, p, n, margin=1.0) + 0.6 * F.triplet_margin_loss(a, q, n, margin=0.5)
            + 0.4 * F.triplet_margin_loss(a, p, q, margin=0.5))


def bench(fn, iters=300):
    for _ in range(20):
        for x in xs:
            x.grad = None
        fn().backward()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(iters):
        for x in xs:
            x.grad = None
        fn().backward()
    e1.record()
    t_submit = time.perf_counter() - t0
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    return wall / iters * 1e6, t_submit / iters * 1e6, e0.elapsed_time(e1) / iters * 1e3


def fwd_only(fn, iters=300):
    with torch.no_grad():
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e6


for name, fn in (("floor", floor), ("ours", ours), ("torch_ops", torch_ops)):
    w, s, d = bench(fn)
    print(f"{name:10s} fwd+bwd: wall {w:7.1f} us/step   host submit {s:7.1f} us/step   device {d:7.1f} us/step")
print(f"ours       fwd only (no_grad): {fwd_only(ours):7.1f} us/step;  torch_ops fwd only: {fwd_only(torch_ops):7.1f} us/step")
import cProfile
import pstats
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    for x in xs:
        x.grad = None
    ours().backward()
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(14)
