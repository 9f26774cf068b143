"""Parity of the fused quadruplet-loss kernels (K5) with the reference, through the C ABI.

Tolerance (BASELINE.json north_star): loss values and gradients within 1e-5 relative.
Golden vectors come from the reference's own module (tests/golden/make_loss_golden.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _dev():
    return torch.device("cuda:0")


def _close(got, want, what, atol=2e-6):
    # 1e-5 relative to the magnitude of the tensor (elementwise for losses, which are O(1))
    got = got.detach().float().cpu()
    want = torch.as_tensor(want).float()
    scale = max(float(want.abs().max()), 1e-3) if want.numel() else 1.0
    err = float((got - want).abs().max()) if want.numel() else 0.0
    assert got.shape == want.shape, (what, got.shape, want.shape)
    assert err <= RTOL * scale + atol, f"{what}: max abs err {err:.3e} vs scale {scale:.3e}"


def _kw(c):
    return dict(gamma=c["gamma"], margin_pos_neg=c["margin_pos_neg"], margin_pos_part=c["margin_pos_part"],
                margin_part_neg=c["margin_part_neg"], p=c["p"], swap=c["swap"], reduction=c["reduction"])


def test_golden_autograd_path(loss_golden):
    import qst_b200
    z, cases = loss_golden
    for c in cases:
        xs = [torch.from_numpy(z[f"{c['key']}_x_{n}"]).to(_dev()).requires_grad_(True) for n in ("a", "p", "pp", "n")]
        out = qst_b200.gamma_quadruplet_loss(*xs, **_kw(c))
        out.sum().backward()
        _close(out, z[f"{c['key']}_out"], f"loss {c}")
        for n, x in zip(("a", "p", "pp", "n"), xs):
            _close(x.grad, z[f"{c['key']}_g_{n}"], f"grad_{n} {c}", atol=1e-7)


def test_golden_fused_path(loss_golden):
    import qst_b200
    z, cases = loss_golden
    for c in cases:
        xs = [torch.from_numpy(z[f"{c['key']}_x_{n}"]).to(_dev()) for n in ("a", "p", "pp", "n")]
        out, grads = qst_b200.gamma_quadruplet_loss_and_grads(*xs, **_kw(c))
        _close(out, z[f"{c['key']}_out"], f"fused loss {c}")
        for n, g in zip(("a", "p", "pp", "n"), grads):
            _close(g, z[f"{c['key']}_g_{n}"], f"fused grad_{n} {c}", atol=1e-7)


@pytest.mark.parametrize("swap", [False, True])
@pytest.mark.parametrize("reduction", ["mean", "sum", "none"])
def test_config2_shape_vs_oracle(swap, reduction):
    """BASELINE.json config 2: 4096 quadruplets x 768-d."""
    import qst_b200
    from oracle import loss_oracle
    xs = qst_b200.synth.quadruplet_batch(4096, 768)
    kw = dict(gamma=0.6, margin_pos_neg=1.0, margin_pos_part=0.5, margin_part_neg=0.5, p=2.0, swap=swap,
              reduction=reduction)
    want, want_g = loss_oracle.loss_and_grads(*xs, **kw)
    leaves = [x.to(_dev()).requires_grad_(True) for x in xs]
    out = qst_b200.gamma_quadruplet_loss(*leaves, **kw)
    out.sum().backward()
    _close(out, want, "loss")
    for leaf, g in zip(leaves, want_g):
        _close(leaf.grad, g, "grad", atol=1e-9)
    out2, grads2 = qst_b200.gamma_quadruplet_loss_and_grads(*[x.to(_dev()) for x in xs], **kw)
    _close(out2, want, "fused loss")
    for g2, g in zip(grads2, want_g):
        _close(g2, g, "fused grad", atol=1e-9)


def test_module_protocol_and_identities():
    """The notebook identities (quadruplet_loss_test.ipynb cells 9/13) and the kwargs call of
    models/quadruplet_sentence_transformer.py:69-75."""
    import qst_b200
    g = torch.Generator().manual_seed(14)
    xs = [torch.randn(5, 10, generator=g).to(_dev()) for _ in range(4)]
    mod = qst_b200.GammaQuadrupletLoss(gamma=0.6, margin_pos_neg=1.0, margin_pos_part=0.5, margin_part_neg=0.5)
    none = mod(x_anchor=xs[0], x_pos=xs[1], x_part=xs[2], x_neg=xs[3], reduction="none")
    mean = mod(x_anchor=xs[0], x_pos=xs[1], x_part=xs[2], x_neg=xs[3])
    total = mod(x_anchor=xs[0], x_pos=xs[1], x_part=xs[2], x_neg=xs[3], reduction="sum", unused_kwarg=1)
    fn = qst_b200.gamma_quadruplet_loss(*xs)
    assert none.shape == (5,)
    torch.testing.assert_close(none.mean(), mean, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(total / 5, mean, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(fn, mean, rtol=0, atol=0)
    assert mod.gamma == 0.6


def test_no_grad_autocast_and_half_inputs():
    """models/evaluators.py:82,92-96: the loss runs under no_grad and autocast."""
    import qst_b200
    from oracle import loss_oracle
    xs = qst_b200.synth.quadruplet_batch(64, 128)
    want = loss_oracle.gamma_quadruplet_loss(*xs)
    mod = qst_b200.GammaQuadrupletLoss(margin_pos_part=0.5, margin_part_neg=0.5)
    with torch.no_grad():
        out = mod(*[x.to(_dev()) for x in xs])
    assert not out.requires_grad
    _close(out, want, "no_grad loss")
    with torch.autocast("cuda", dtype=torch.float16):
        out16 = mod(*[x.to(_dev()).half() for x in xs])
    assert out16.dtype == torch.float32
    want16 = loss_oracle.gamma_quadruplet_loss(*[x.half().float() for x in xs])
    _close(out16, want16, "autocast loss")
    # native half / bf16 inputs: fp32 math inside, gradients come back in the input dtype
    for dt in (torch.float16, torch.bfloat16):
        leaves = [x.to(_dev()).to(dt).requires_grad_(True) for x in xs]
        out = mod(*leaves)
        out.backward()
        ref_in = [x.to(dt).float() for x in xs]
        want_dt, want_g = loss_oracle.loss_and_grads(*ref_in, margin_pos_part=0.5, margin_part_neg=0.5,
                                                     margin_pos_neg=1.0, gamma=0.6)
        assert out.dtype == dt and leaves[0].grad.dtype == dt
        assert abs(float(out) - float(want_dt)) <= 1e-2 * abs(float(want_dt))
        for leaf, g in zip(leaves, want_g):
            assert float((leaf.grad.float().cpu() - g).abs().max()) <= 1e-2 * float(g.abs().max()) + 1e-6


def test_edge_shapes():
    import qst_b200
    from oracle import loss_oracle
    for B, D in [(1, 1), (3, 7), (2, 1025), (130, 33)]:
        g = torch.Generator().manual_seed(B * 1000 + D)
        xs = [torch.randn(B, D, generator=g) for _ in range(4)]
        for p in (2.0, 1.0, 3.0, float("inf")):
            want, want_g = loss_oracle.loss_and_grads(*xs, p=p, swap=True, reduction="sum")
            leaves = [x.to(_dev()).requires_grad_(True) for x in xs]
            out = qst_b200.gamma_quadruplet_loss(*leaves, p=p, swap=True, reduction="sum")
            out.backward()
            _close(out, want, f"loss B={B} D={D} p={p}")
            for leaf, gg in zip(leaves, want_g):
                _close(leaf.grad, gg, f"grad B={B} D={D} p={p}", atol=1e-7)
    # unaligned views fall back to the scalar loader
    base = torch.randn(4, 8, 65, device=_dev())
    xs = [base[i, :, 1:] for i in range(4)]
    want = loss_oracle.gamma_quadruplet_loss(*[x.cpu() for x in xs])
    _close(qst_b200.gamma_quadruplet_loss(*xs), want, "unaligned")
    # empty batch
    e = torch.empty(0, 16, device=_dev())
    assert qst_b200.gamma_quadruplet_loss(e, e, e, e, reduction="none").shape == (0,)
    assert float(qst_b200.gamma_quadruplet_loss(e, e, e, e, reduction="sum")) == 0.0


def test_validation_matches_reference():
    import qst_b200
    xs = [torch.zeros(2, 3, device=_dev()) for _ in range(4)]
    for kw in (dict(gamma=-0.1), dict(gamma=1.5), dict(margin_pos_neg=0), dict(margin_pos_part=-1),
               dict(margin_part_neg=0), dict(p=0), dict(reduction="avg")):
        with pytest.raises(ValueError):
            qst_b200.gamma_quadruplet_loss(*xs, **kw)
        if "reduction" not in kw or True:
            with pytest.raises(ValueError):
                qst_b200.GammaQuadrupletLoss(**kw)
    m = qst_b200.GammaQuadrupletLoss()
    with pytest.raises(ValueError):
        m.margin_part_neg = 0
    with pytest.raises(ValueError):
        m.reduction = "x"
    with pytest.raises(qst_b200.QstError):
        m(*[x.cpu() for x in xs])   # no CPU fallback


def test_cross_cta_reduction_sequences_reproducibility_and_special_partials():
    """The fused kernel's 'mean' / 'sum' travel between CTAs as fixed-point limbs inside one atomic
    instruction per CTA (no fence, no ticket): launches of different grid sizes on ONE workspace, the same
    input twice (bitwise equal), sums far beyond 2^63, negative partial sums (gamma outside [0, 1], reachable
    through the C ABI only), NaN and inf rows -- which must come out as the reference's float arithmetic
    gives them -- followed by an ordinary launch again."""
    import ctypes as C
    import qst_b200
    from qst_b200 import _lib, quad_loss
    from oracle import loss_oracle
    dev = _dev()
    kw = dict(gamma=0.6, margin_pos_neg=1.0, margin_pos_part=0.5, margin_part_neg=0.5, swap=True)

    def batch(B, D, seed):
        g = torch.Generator().manual_seed(seed)
        return [torch.randn(B, D, generator=g) for _ in range(4)]

    def run(xs, **over):
        k = dict(kw, **over)
        want = loss_oracle.gamma_quadruplet_loss(*xs, **k)
        got, _ = qst_b200.gamma_quadruplet_loss_and_grads(*[x.to(dev) for x in xs], **k)
        return got, want

    for B, red in [(4096, "mean"), (5, "sum"), (1000, "mean"), (1, "mean"), (1777, "sum"), (4096, "sum")]:
        xs = batch(B, 256, B)
        got, want = run(xs, reduction=red, p=2.0)
        _close(got, want, f"B={B} {red}")
        again, _ = run(xs, reduction=red, p=2.0)
        assert torch.equal(got, again), "same input, same workspace: bitwise equal"

    xs = batch(600, 256, 7)
    for scale in (1e17, 1e30):                         # p=1 distances ~1e20 / ~1e33 per row: far beyond 2^63, finite in fp32
        big = [x * scale for x in xs]
        got, want = run(big, reduction="mean", p=1.0)
        assert torch.isfinite(want) and float(want) > 1e18
        _close(got, want, f"partials of magnitude {scale:g}")
    for bad in (float("nan"), float("inf")):
        ys = [x.clone() for x in xs]
        ys[1][17, 3] = bad
        got, want = run(ys, reduction="sum", p=2.0)
        torch.testing.assert_close(got.cpu(), want, rtol=1e-5, atol=1e-6, equal_nan=True)
    got, want = run(xs, reduction="mean", p=2.0)     # the words were left clean
    _close(got, want, "after the special partials")

    # negative partial sums: gamma = 3 weighs the third hinge term with -2 (rejected by the Python API, not by the C ABI)
    lib = _lib.load()
    xd = [x.to(dev) for x in xs]
    prm = quad_loss._params(3.0, 1.0, 0.5, 0.5, 2.0, True)
    ws = torch.zeros(lib.qst_quadruplet_workspace_bytes(), dtype=torch.uint8, device=dev)
    grads = [torch.empty_like(x) for x in xd]
    out = {}
    for name, red, n in (("none", _lib.QST_RED_NONE, 600), ("sum", _lib.QST_RED_SUM, 1), ("mean", _lib.QST_RED_MEAN, 1)):
        out[name] = torch.empty(n, device=dev)
        _lib.check(lib.qst_quadruplet_fwd_bwd(xd[0].data_ptr(), xd[1].data_ptr(), xd[2].data_ptr(), xd[3].data_ptr(),
                                              _lib.QST_F32, 600, 256, C.byref(prm), red, 1.0, out[name].data_ptr(),
                                              grads[0].data_ptr(), grads[1].data_ptr(), grads[2].data_ptr(),
                                              grads[3].data_ptr(), ws.data_ptr(), _lib.stream_ptr(dev)))
    rows = out["none"].double().cpu()
    assert int((rows < 0).sum()) > 50 and int((rows > 0).sum()) > 50
    assert float(out["sum"]) == float(rows.sum().float()) and float(out["mean"]) == float((rows.sum() / 600).float())
