"""Generate tests/golden/loss_golden.npz from the REFERENCE's own loss module.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_loss_golden.py

Imports ``/root/reference/models/losses/losses.py`` by file path (it only needs
torch), evaluates ``gamma_quadruplet_loss`` and ``GammaQuadrupletLoss`` on seeded
inputs over a sweep of (B, D, p, swap, reduction, gamma, margins) and stores
inputs, outputs and autograd gradients of ``out.sum()``.  Seed 14 is the
reference's RANDOM_SEED (``dataset/constants.py:5``).
"""
import itertools
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.loss_oracle import load_reference_losses  # noqa: E402


def main():
    ref = load_reference_losses()
    if ref is None:
        raise SystemExit("/root/reference not present: golden vectors can only be made in the authoring container")
    torch.manual_seed(14)
    torch.set_num_threads(1)
    shapes = [(5, 10), (7, 19), (9, 64), (3, 300)]
    ps = [2.0, 1.0, 3.0, 0.5, float("inf")]
    cases = []
    arrays = {}
    cid = 0
    for (B, D), p, swap in itertools.product(shapes, ps, (False, True)):
        g = torch.Generator().manual_seed(14 + cid)
        xs = [torch.randn(B, D, generator=g) for _ in range(4)]
        # make some rows easy (zero loss terms) and some hard, so clamps are exercised both ways
        xs[1] = xs[0] + 0.05 * xs[1]          # positives close to the anchor
        xs[2] = xs[0] + 0.6 * xs[2]           # partial positives mid-way
        if B >= 7:
            xs[3][:3] = xs[0][:3] + 0.01 * xs[3][:3]   # a few negatives nearly on the anchor
        param_sets = ((0.6, 1.0, 0.5, 0.5), (0.0, 0.3, 2.0, 1.0), (1.0, 2.5, 0.1, 0.7))
        for r_i, reduction in enumerate(("none", "sum", "mean")):
            gamma, m_pn, m_pp, m_partn = param_sets[(cid // 3 + r_i) % 3]
            leaves = [x.clone().requires_grad_(True) for x in xs]
            out = ref.gamma_quadruplet_loss(*leaves, gamma=gamma, margin_pos_neg=m_pn,
                                            margin_pos_part=m_pp, margin_part_neg=m_partn,
                                            p=p, swap=swap, reduction=reduction)
            out.sum().backward()
            mod = ref.GammaQuadrupletLoss(gamma=gamma, margin_pos_neg=m_pn, margin_pos_part=m_pp,
                                          margin_part_neg=m_partn, p=p, swap=swap, reduction="mean")
            out_mod = mod(x_anchor=xs[0], x_pos=xs[1], x_part=xs[2], x_neg=xs[3], reduction=reduction)
            assert torch.equal(out.detach(), out_mod), "module != functional in the reference"
            key = f"c{cid:04d}"
            for name, x in zip(("a", "p", "pp", "n"), xs):
                arrays[f"{key}_x_{name}"] = x.numpy()
            arrays[f"{key}_out"] = out.detach().numpy()
            for name, leaf in zip(("a", "p", "pp", "n"), leaves):
                arrays[f"{key}_g_{name}"] = leaf.grad.numpy()
            cases.append(dict(key=key, B=B, D=D, p=("inf" if p == float("inf") else p), swap=swap,
                              reduction=reduction, gamma=gamma, margin_pos_neg=m_pn,
                              margin_pos_part=m_pp, margin_part_neg=m_partn))
            cid += 1
    arrays["cases_json"] = np.frombuffer(json.dumps(cases).encode(), dtype=np.uint8)
    out_path = os.path.join(HERE, "loss_golden.npz")
    np.savez_compressed(out_path, **arrays)
    print(f"wrote {out_path}: {len(cases)} cases, {os.path.getsize(out_path)/1e6:.2f} MB, torch {torch.__version__}")


if __name__ == "__main__":
    main()
