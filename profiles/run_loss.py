"""ncu target for config 2: fused forward+backward of the quadruplet loss, 4096 x 768 fp32, over
rotating input sets larger than L2."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qst_b200  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(14)
sets = [[torch.randn(4096, 768, generator=g, device=dev) for _ in range(4)] for _ in range(8)]
for i in range(16):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    loss, grads = qst_b200.gamma_quadruplet_loss_and_grads(*sets[i % 8])
    b.record()
    torch.cuda.synchronize()
print(f"last call {a.elapsed_time(b) * 1e3:.1f} us (includes Python launch overhead), loss {float(loss):.6f}")
