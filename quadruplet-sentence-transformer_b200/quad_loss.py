"""Drop-in quadruplet losses backed by the fused sm_100a kernels (K5).

Mirrors ``/root/reference/models/losses/losses.py``:

* ``gamma_quadruplet_loss``  (``:9-69``)   -- same signature, defaults and ``ValueError``s
* ``QuadrupletLoss``         (``:157-238``) -- ABC with validated properties
* ``GammaQuadrupletLoss``    (``:241-303``) -- ``forward(x_anchor, x_pos, x_part, x_neg,
  reduction=None, **kwargs)`` as called at ``models/quadruplet_sentence_transformer.py:69-75``;
  ``.gamma`` is read by ``models/evaluators.py:593``.

The arithmetic runs in ``libqst.so`` (``qst_quadruplet_fwd`` / ``_bwd`` / ``_fwd_bwd``);
there is no PyTorch fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from abc import ABC, abstractmethod
from typing import Optional

import torch

from . import _lib

DEFAULT_GAMMA = 0.6
REDUCTIONS = frozenset(["mean", "sum", "none"])
EPS = 1e-6  # torch's triplet_margin_loss default, which the reference never overrides

_workspaces = {}


def _workspace(device: torch.device) -> torch.Tensor:
    """Per-(device, stream) reduction scratch, zero-filled once (the kernels re-zero it)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None:
        n = _lib.load().qst_quadruplet_workspace_bytes()
        ws = torch.zeros(n, dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _validate(gamma, margin_pos_neg, margin_pos_part, margin_part_neg, p, reduction):
    """Conditions, order and messages of losses.py:20-32."""
    if not 0 <= gamma <= 1:
        raise ValueError(f"gamma must be between 0 and 1, {gamma} given")
    for name, v in (("margin_pos_neg", margin_pos_neg), ("margin_pos_part", margin_pos_part),
                    ("margin_part_neg", margin_part_neg)):
        if v <= 0:
            raise ValueError(f"{name} must be positive, {v} given")
    if reduction not in REDUCTIONS:
        raise ValueError(f"reduction must be one of: {REDUCTIONS}, {reduction} given")
    if p <= 0:
        raise ValueError(f"p must be positive, {p} given")


_param_cache = {}


def _params(gamma, margin_pos_neg, margin_pos_part, margin_part_neg, p, swap) -> _lib.QuadParams:
    """``qst_quad_params`` for these hyper-parameters (cached: a training loop passes the same ones
    every step and building the ctypes struct is a measurable part of a 20 us call)."""
    key = (gamma, margin_pos_neg, margin_pos_part, margin_part_neg, p, bool(swap))
    prm = _param_cache.get(key)
    if prm is None:
        prm = _lib.QuadParams(float(gamma), float(1.0 - float(gamma)), float(margin_pos_neg),
                              float(margin_pos_part), float(margin_part_neg), float(p), EPS, int(bool(swap)))
        if len(_param_cache) < 256:
            _param_cache[key] = prm
    return prm


_KERNEL_DTYPES = (torch.float32, torch.float16, torch.bfloat16)
_DTYPE_CODE = {torch.float32: _lib.QST_F32, torch.float16: _lib.QST_F16, torch.bfloat16: _lib.QST_BF16}
_fwd_bwd_fn = None


def _fwd_bwd():
    """``qst_quadruplet_fwd_bwd`` bound once (the library lookup is off the per-step path)."""
    global _fwd_bwd_fn
    if _fwd_bwd_fn is None:
        _fwd_bwd_fn = _lib.load().qst_quadruplet_fwd_bwd
    return _fwd_bwd_fn


def _same_layout(xs) -> bool:
    """True when the four inputs can go to the kernel as they are (the case inside
    ``SentenceTransformer.fit``: four [B, D] embeddings of one dtype on one device)."""
    a = xs[0]
    if not (a.is_cuda and a.dim() >= 1 and a.dtype in _KERNEL_DTYPES and a.is_contiguous()):
        return False
    shape, dtype, device = a.shape, a.dtype, a.device
    for x in xs[1:]:
        if not isinstance(x, torch.Tensor) or x.shape != shape or x.dtype is not dtype or x.device != device \
                or not x.is_contiguous():
            return False
    return not torch.is_autocast_enabled()


def _prepare(x_anchor, x_pos, x_part, x_neg):
    xs = [x_anchor, x_pos, x_part, x_neg]
    _lib.require_cuda(*xs)
    if torch.is_autocast_enabled():
        # torch runs triplet_margin_loss in float32 under autocast (models/evaluators.py:92-96)
        xs = [x.float() for x in xs]
    dt = torch.result_type(torch.result_type(xs[0], xs[1]), torch.result_type(xs[2], xs[3])) \
        if len({x.dtype for x in xs}) > 1 else xs[0].dtype
    if dt not in (torch.float32, torch.float16, torch.bfloat16):
        dt = torch.float32
    shape = torch.broadcast_shapes(*[x.shape for x in xs])
    if len(shape) == 0:
        raise ValueError("inputs must have at least one dimension")
    out = []
    for x in xs:
        if x.dtype != dt:
            x = x.to(dt)
        if x.shape != shape:
            x = x.expand(shape)
        out.append(x.contiguous())
    D = shape[-1]
    B = 1
    for s in shape[:-1]:
        B *= s
    return out, shape, B, D, dt


class _FusedQuadrupletFn(torch.autograd.Function):
    """Training-step path for inputs that already agree in shape / dtype / device: ONE launch of
    ``qst_quadruplet_fwd_bwd`` in ``forward`` produces the loss and the four gradients for an upstream
    gradient of 1 (each input read once, each gradient written once: 8*B*D*itemsize bytes);
    ``backward`` only scales them by the upstream gradient (one launch over the stacked gradient
    buffer).  Nothing is saved for backward but the gradients themselves.  Gradient slots of inputs that
    do not require grad are never written by the kernel and never returned."""

    @staticmethod
    def forward(ctx, x_anchor, x_pos, x_part, x_neg, prm, red):
        shape = x_anchor.shape
        D = shape[-1]
        n = x_anchor.numel()
        B = n // D if D else 0
        dev, dt = x_anchor.device, x_anchor.dtype
        need = ctx.needs_input_grad
        guard = torch.cuda.device(dev) if dev.index != torch.cuda.current_device() else None
        if guard is not None:
            guard.__enter__()
        try:
            loss = torch.empty(shape[:-1] if red == _lib.QST_RED_NONE else (), dtype=torch.float32, device=dev)
            buf = torch.empty((4,) + tuple(shape), dtype=dt, device=dev)
            g0 = buf.data_ptr()
            step = n * buf.element_size()
            stream = torch.cuda.current_stream(dev).cuda_stream
            ws = _workspaces.get((dev.index, stream))
            if ws is None:
                ws = _workspace(dev)
            rc = _fwd_bwd()(
                x_anchor.data_ptr(), x_pos.data_ptr(), x_part.data_ptr(), x_neg.data_ptr(), _DTYPE_CODE[dt], B, D,
                prm, red, 1.0, loss.data_ptr(),
                g0 if need[0] else None, g0 + step if need[1] else None,
                g0 + 2 * step if need[2] else None, g0 + 3 * step if need[3] else None,
                ws.data_ptr(), stream)
            if rc != 0:
                _lib.check(rc)
        finally:
            if guard is not None:
                guard.__exit__(None, None, None)
        ctx.buf, ctx.per_row = buf, red == _lib.QST_RED_NONE
        return loss if dt == torch.float32 else loss.to(dt)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        buf = ctx.buf
        if grad_out.dtype != buf.dtype:
            grad_out = grad_out.to(buf.dtype)
        # out of place: the unit-upstream gradients stay intact, so backward(retain_graph=True) can be
        # called again like on the reference's autograd graph
        g = torch.mul(buf, grad_out.reshape((1,) + tuple(grad_out.shape) + (1,)) if ctx.per_row else grad_out).unbind(0)
        need = ctx.needs_input_grad
        return (g[0] if need[0] else None, g[1] if need[1] else None, g[2] if need[2] else None,
                g[3] if need[3] else None, None, None)


class _QuadrupletFn(torch.autograd.Function):
    """General path (broadcasting, mixed dtypes, autocast): forward kernel, then the backward kernel
    from the saved distances."""

    @staticmethod
    def forward(ctx, x_anchor, x_pos, x_part, x_neg, prm, reduction):
        lib = _lib.load()
        xs, shape, B, D, dt = _prepare(x_anchor, x_pos, x_part, x_neg)
        dev = xs[0].device
        red = _lib.REDUCTION_CODES[reduction]
        needs_grad = any(ctx.needs_input_grad[:4])
        with torch.cuda.device(dev):
            loss = torch.empty(shape[:-1] if red == _lib.QST_RED_NONE else (), dtype=torch.float32, device=dev)
            saved = torch.empty((B, _lib.QST_QUAD_SAVED_PER_ROW), dtype=torch.float32, device=dev) if needs_grad else None
            ws = _workspace(dev)
            _lib.check(lib.qst_quadruplet_fwd(xs[0].data_ptr(), xs[1].data_ptr(), xs[2].data_ptr(), xs[3].data_ptr(),
                                              _lib.dtype_code(dt), B, D, C.byref(prm), red, loss.data_ptr(),
                                              _lib.ptr(saved), ws.data_ptr(), _lib.stream_ptr(dev)))
        if needs_grad:
            ctx.save_for_backward(*xs, saved)
            ctx.prm, ctx.red, ctx.B, ctx.D, ctx.dt = prm, red, B, D, dt
            ctx.in_shapes = [t.shape for t in (x_anchor, x_pos, x_part, x_neg)]
            ctx.in_dtypes = [t.dtype for t in (x_anchor, x_pos, x_part, x_neg)]
            ctx.bshape = shape
        return loss if dt == torch.float32 else loss.to(dt)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        lib = _lib.load()
        *xs, saved = ctx.saved_tensors
        dev = xs[0].device
        grad_out = grad_out.to(torch.float32).contiguous()
        if ctx.red == _lib.QST_RED_NONE:
            grad_out = grad_out.expand(ctx.bshape[:-1]).contiguous()
        grads = []
        with torch.cuda.device(dev):
            for i in range(4):
                grads.append(torch.empty((ctx.B, ctx.D), dtype=ctx.dt, device=dev)
                             if ctx.needs_input_grad[i] else None)
            _lib.check(lib.qst_quadruplet_bwd(xs[0].data_ptr(), xs[1].data_ptr(), xs[2].data_ptr(), xs[3].data_ptr(),
                                              _lib.dtype_code(ctx.dt), ctx.B, ctx.D, C.byref(ctx.prm), ctx.red,
                                              saved.data_ptr(), grad_out.data_ptr(),
                                              _lib.ptr(grads[0]), _lib.ptr(grads[1]), _lib.ptr(grads[2]),
                                              _lib.ptr(grads[3]), _lib.stream_ptr(dev)))
        outs = []
        for g, shp, dt in zip(grads, ctx.in_shapes, ctx.in_dtypes):
            if g is None:
                outs.append(None)
                continue
            g = g.view(ctx.bshape)
            if tuple(shp) != tuple(ctx.bshape):  # undo broadcasting
                g = g.sum_to_size(shp)
            outs.append(g.to(dt) if g.dtype != dt else g)
        return (*outs, None, None)


def gamma_quadruplet_loss(x_anchor: torch.Tensor, x_pos: torch.Tensor, x_part: torch.Tensor,
                          x_neg: torch.Tensor, gamma: float = DEFAULT_GAMMA, margin_pos_neg: float = 1.0,
                          margin_pos_part: float = 0.5, margin_part_neg: float = 0.5, p: float = 2.0,
                          swap: bool = False, reduction: str = "mean") -> torch.Tensor:
    """Same contract as ``models/losses/losses.py:9-69``."""
    # hyper-parameters seen (and validated) before are served from the cache: the checks and the ctypes
    # struct are a measurable share of a 20 us call
    key = (gamma, margin_pos_neg, margin_pos_part, margin_part_neg, p, bool(swap))
    prm = _param_cache.get(key)
    red = _lib.REDUCTION_CODES.get(reduction)
    if prm is None or red is None:
        _validate(gamma, margin_pos_neg, margin_pos_part, margin_part_neg, p, reduction)
        prm = _params(gamma, margin_pos_neg, margin_pos_part, margin_part_neg, p, swap)
        red = _lib.REDUCTION_CODES[reduction]
    if (torch.is_grad_enabled() and isinstance(x_anchor, torch.Tensor)
            and (x_anchor.requires_grad or x_pos.requires_grad or x_part.requires_grad or x_neg.requires_grad)
            and _same_layout((x_anchor, x_pos, x_part, x_neg))):
        return _FusedQuadrupletFn.apply(x_anchor, x_pos, x_part, x_neg, prm, red)
    return _QuadrupletFn.apply(x_anchor, x_pos, x_part, x_neg, prm, reduction)


def gamma_quadruplet_loss_and_grads(x_anchor, x_pos, x_part, x_neg, gamma=DEFAULT_GAMMA, margin_pos_neg=1.0,
                                    margin_pos_part=0.5, margin_part_neg=0.5, p=2.0, swap=False,
                                    reduction="mean", upstream: float = 1.0):
    """One-launch training step: loss and d(loss*upstream)/d(inputs) (``qst_quadruplet_fwd_bwd``).

    Equivalent to ``loss = gamma_quadruplet_loss(...); (loss.sum()*upstream).backward()`` but each
    input is read once and each gradient written once (8*B*D*itemsize bytes of HBM traffic).
    """
    _validate(gamma, margin_pos_neg, margin_pos_part, margin_part_neg, p, reduction)
    lib = _lib.load()
    prm = _params(gamma, margin_pos_neg, margin_pos_part, margin_part_neg, p, swap)
    if isinstance(x_anchor, torch.Tensor) and _same_layout((x_anchor, x_pos, x_part, x_neg)):
        xs, shape, dt = [x_anchor, x_pos, x_part, x_neg], x_anchor.shape, x_anchor.dtype   # nothing to convert
        D = shape[-1]
        B = x_anchor.numel() // D if D else 0
    else:
        xs, shape, B, D, dt = _prepare(x_anchor.detach(), x_pos.detach(), x_part.detach(), x_neg.detach())
    dev = xs[0].device
    red = _lib.REDUCTION_CODES[reduction]
    with torch.cuda.device(dev):
        loss = torch.empty(shape[:-1] if red == _lib.QST_RED_NONE else (), dtype=torch.float32, device=dev)
        grads = [torch.empty(shape, dtype=dt, device=dev) for _ in range(4)]
        ws = _workspace(dev)
        _lib.check(lib.qst_quadruplet_fwd_bwd(xs[0].data_ptr(), xs[1].data_ptr(), xs[2].data_ptr(), xs[3].data_ptr(),
                                              _lib.dtype_code(dt), B, D, C.byref(prm), red, float(upstream),
                                              loss.data_ptr(), grads[0].data_ptr(), grads[1].data_ptr(),
                                              grads[2].data_ptr(), grads[3].data_ptr(), ws.data_ptr(),
                                              _lib.stream_ptr(dev)))
    return loss, grads


class _Validated:
    """Attribute whose assignment is checked with the reference's condition and message."""

    def __init__(self, check, message):
        self._check, self._message = check, message

    def __set_name__(self, owner, name):
        self._name, self._slot = name, "_v_" + name

    def __get__(self, obj, objtype=None):
        return self if obj is None else getattr(obj, self._slot)

    def __set__(self, obj, value):
        if not self._check(value):
            raise ValueError(self._message.format(name=self._name, value=value))
        object.__setattr__(obj, self._slot, value)


_positive = lambda: _Validated(lambda v: v > 0, "{name} must be positive, {value} given")  # noqa: E731


class QuadrupletLoss(torch.nn.Module, ABC):
    """Base class with the constructor, validation and attributes of
    ``models/losses/losses.py:157-238`` (``margin_pos_neg``, ``margin_pos_part``, ``p``, ``swap``,
    ``reduction`` are readable and writable; bad values raise ``ValueError`` on assignment)."""

    margin_pos_neg = _positive()
    margin_pos_part = _positive()
    p = _positive()
    reduction = _Validated(lambda v: v in REDUCTIONS,
                           "{name} must be one of: " + str(REDUCTIONS).replace("{", "{{").replace("}", "}}")
                           + ", {value} given")
    swap = _Validated(lambda v: True, "")

    def __init__(self, margin_pos_neg: float = 1.0, margin_pos_part: float = 1.0, p: float = 2.0,
                 swap: bool = False, reduction: str = "mean"):
        super().__init__()
        self.margin_pos_neg, self.margin_pos_part = margin_pos_neg, margin_pos_part
        self.reduction, self.p, self.swap = reduction, p, swap

    @abstractmethod
    def forward(self, x_anchor: torch.Tensor, x_pos: torch.Tensor, x_part: torch.Tensor,
                x_neg: torch.Tensor, reduction: Optional[str] = None, **kwargs) -> torch.Tensor:
        raise NotImplementedError()


class GammaQuadrupletLoss(QuadrupletLoss):
    """``models/losses/losses.py:241-303``: adds ``gamma`` (read by ``models/evaluators.py:593``)
    and ``margin_part_neg``; ``forward`` lets a per-call ``reduction`` override the constructor's
    (``:291``) and runs the fused kernel."""

    gamma = _Validated(lambda v: 0 <= v <= 1, "{name} must be between 0 and 1, {value} given")
    margin_part_neg = _positive()

    def __init__(self, gamma: float = DEFAULT_GAMMA, margin_pos_neg: float = 1.0,
                 margin_pos_part: float = 1.0, margin_part_neg: float = 1.0, p: float = 2.0,
                 swap: bool = False, reduction: str = "mean"):
        super().__init__(margin_pos_neg=margin_pos_neg, margin_pos_part=margin_pos_part, p=p,
                         swap=swap, reduction=reduction)
        self.gamma, self.margin_part_neg = gamma, margin_part_neg

    def forward(self, x_anchor: torch.Tensor, x_pos: torch.Tensor, x_part: torch.Tensor,
                x_neg: torch.Tensor, reduction: Optional[str] = None, **kwargs) -> torch.Tensor:
        return gamma_quadruplet_loss(
            x_anchor, x_pos, x_part, x_neg, gamma=self.gamma, margin_pos_neg=self.margin_pos_neg,
            margin_pos_part=self.margin_pos_part, margin_part_neg=self.margin_part_neg, p=self.p,
            swap=self.swap, reduction=self.reduction if reduction is None else reduction)

    def loss_and_grads(self, x_anchor, x_pos, x_part, x_neg, reduction: Optional[str] = None,
                       upstream: float = 1.0):
        """One-launch fused forward+backward (see ``gamma_quadruplet_loss_and_grads``)."""
        return gamma_quadruplet_loss_and_grads(
            x_anchor, x_pos, x_part, x_neg, gamma=self.gamma, margin_pos_neg=self.margin_pos_neg,
            margin_pos_part=self.margin_pos_part, margin_part_neg=self.margin_part_neg, p=self.p,
            swap=self.swap, reduction=self.reduction if reduction is None else reduction,
            upstream=upstream)
