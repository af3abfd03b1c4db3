"""Backward building blocks shared by the decoders' autograd Functions (all compute on libccx).

Linear backward re-uses the forward tcgen05 GEMM (``ccx_linear``):
    dX[M,K] = dY[M,N] . W[N,K]        -> A = dY (row-major, contraction over N), B = W^T [K, N]
    dW[N,K] = dY^T[N,M] . X[M,K]      -> A = dY^T [N, M], B = X^T [K, M]   (+= through the `residual` epilogue)
    db[N]   = column sums of dY
bf16: the tensor core reads an operand's transpose in place (``a_mn`` / ``w_mn`` of ccx_linear: MN-major UMMA
descriptors), so B = W as the forward has it and A = dY, B = X as they are — one bf16 cast of dY is the only copy.
fp32 (3xTF32, K-major only): transposed / converted operands from ``ccx_convert_operand`` (optionally fused with a
dropout multiplier or the ReLU mask).
"""
import os
import torch

from . import _lib
from ._lib import Operand, ptr


def zeros_many(shapes, device):
    """fp32 zero tensors of the given shapes carved out of ONE allocation and ONE fill kernel (a backward pass
    otherwise issues ~100 tiny memset launches).  Views start on 256-byte boundaries."""
    sizes = [max(1, int(torch.Size(sh).numel())) for sh in shapes]
    offs, total = [], 0
    for n in sizes:
        offs.append(total)
        total += (n + 63) // 64 * 64
    flat = torch.zeros(total, dtype=torch.float32, device=device)
    return [flat[o:o + torch.Size(sh).numel()].view(sh) for o, sh in zip(offs, shapes)]


# ---- direct gradient accumulation --------------------------------------------------------------------------------------
# By default a backward pass accumulates into fresh zero buffers and RETURNS them; autograd's AccumulateGrad then adds
# each one into ``p.grad`` (one more kernel per parameter, ~50 per train step, plus the zero fills).  Inside
# ``with direct_grads():`` the accumulators ARE the existing ``p.grad`` tensors (the weight-gradient GEMMs accumulate
# in place through their residual epilogue) and the Function returns None for those parameters.  The caller owns the
# zeroing of ``p.grad`` (``optimizer.zero_grad(set_to_none=False)`` / the flat buckets of CapturedTrainStep).
_DIRECT = [False]


class direct_grads:
    def __enter__(self):
        self._old = _DIRECT[0]
        _DIRECT[0] = True

    def __exit__(self, *exc):
        _DIRECT[0] = self._old


def _direct_target(p):
    g = p.grad
    if _DIRECT[0] and g is not None and g.dtype == torch.float32 and g.is_contiguous() and g.shape == p.shape \
            and g.device == p.device:
        return g
    return None


def zero_grads_like(named_params, extra_shapes=None):
    """{name: fp32 gradient accumulator} for the parameters that require grad: zero buffers carved from one allocation
    (see zeros_many), or — under direct_grads() — the parameters' own ``.grad`` tensors.
    extra_shapes: further zero buffers carved from the same allocation -> returns (dict, [extras])."""
    items = [(n, p) for n, p in named_params if p.requires_grad]
    if not items:
        return {} if extra_shapes is None else ({}, zeros_many(list(extra_shapes), named_params[0][1].device))
    grads = {}
    fresh = []
    for n, p in items:
        t = _direct_target(p)
        if t is None:
            fresh.append((n, p))
        else:
            grads[n] = t
    shapes = [tuple(p.shape) for _, p in fresh] + list(extra_shapes or ())
    bufs = zeros_many(shapes, items[0][1].device) if shapes else []
    for (n, _), b in zip(fresh, bufs):
        grads[n] = b
    return grads if extra_shapes is None else (grads, bufs[len(fresh):])


def grads_out(grads, named_params):
    """What a Function's backward returns for its parameter inputs: the accumulator, or None when it IS the
    parameter's .grad (already accumulated in place)."""
    out = []
    for n, p in named_params:
        g = grads.get(n)
        out.append(None if (g is not None and p.grad is not None and g.data_ptr() == p.grad.data_ptr()
                            and g.shape == p.grad.shape) else g)
    return tuple(out)


def deliver(grads, name, param, value):
    """Hand a gradient that was computed in its own buffer (a slice of a combined GEMM result, a re-laid-out filter)
    to parameter `name`: added into .grad under direct_grads(), else returned through autograd."""
    t = grads.get(name)
    if t is not None and param.grad is not None and t.data_ptr() == param.grad.data_ptr():
        t.add_(value.view_as(t) if value.shape != t.shape else value)
    else:
        grads[name] = value.contiguous() if not value.is_contiguous() else value


def _pad8(n):
    return (n + 7) // 8 * 8


def _src(x):
    """(hi ptr, lo ptr, dtype code, ld) of a 2-D source: fp32 tensor or Operand."""
    if isinstance(x, Operand):
        return ptr(x.hi), x.lo_ptr, _lib.dt_code(x.dtype), x.hi.stride(0)
    return ptr(x), None, _lib.CCX_F32, x.stride(0)


def to_operand(x, cd, mul=None, mul_mode=0, mul_scale=1.0, transpose=False):
    """2-D fp32 tensor / Operand [R, C] -> Operand in compute dtype `cd`.
    Plain: [R, pad8(C)] storage viewed as [R, C].  Transposed: [C, pad8(R)] with zero padding, viewed [C, pad8(R)]
    (the padded contraction columns are zeros on both GEMM operands)."""
    hi, lo, code, ldx = _src(x)
    R, C = (x.hi.shape if isinstance(x, Operand) else x.shape)
    dev = x.hi.device if isinstance(x, Operand) else x.device
    if transpose:
        Rp = _pad8(R)
        out = Operand.empty((C, Rp), cd, dev)
        view = out
        ldo = Rp
    else:
        Cp = _pad8(C)
        out = Operand.empty((R, Cp), cd, dev)
        view = out.map(lambda t: t[:, :C]) if Cp != C else out
        ldo = Cp
        Rp = 0
    _lib.check(_lib.lib().ccx_convert_operand(hi, lo, code, ldx, ptr(mul), mul.stride(0) if mul is not None else 0,
                                              mul_mode, mul_scale, ptr(out.hi), out.lo_ptr, _lib.dt_code(cd), ldo, R, C,
                                              1 if transpose else 0, Rp, _lib.stream_ptr()), "convert_operand")
    return view


_MN = os.environ.get("CCX_GEMM_MN", "1") != "0"        # A/B switch: 0 = transposed copies as in the fp32 path


class MnWeight:
    """A Linear weight for the dgrad GEMM in bf16: the plain [N, K] operand, read MN-major (no W^T copy)."""

    def __init__(self, op):
        self.op = op


def _mn_ready(op):
    return (isinstance(op, Operand) and op.dtype == torch.bfloat16 and op.lo is None and op.hi.stride(1) == 1 and
            op.hi.stride(0) % 8 == 0 and op.hi.data_ptr() % 16 == 0)


def mn_operands(cd, *ops):
    """True when the bf16 operands `ops` can be read transposed in place (pitch / alignment) and the switch is on."""
    return _MN and cd == torch.bfloat16 and all(_mn_ready(o) for o in ops)


def linear_dgrad(dy_op, wt, residual=None, n=None):
    """dX = dY . W for wt = weight_t(W): MnWeight (bf16, W read in place) or the transposed Operand W^T [K, pad8(N)]."""
    if isinstance(wt, MnWeight):
        return _lib.linear(dy_op, wt.op, residual=residual, w_mn=True)
    return _lib.linear(dy_op, wt, residual=residual, k=n)


def colsum_acc(dy, out, mul=None, mul_mode=0, mul_scale=1.0):
    R, C = dy.shape
    _lib.check(_lib.lib().ccx_colsum_acc(ptr(dy), dy.stride(0), ptr(mul), mul.stride(0) if mul is not None else 0,
                                         mul_mode, mul_scale, ptr(out), R, C, _lib.stream_ptr()), "colsum_acc")


_FUSE_CAST_COLSUM = os.environ.get("CCX_FUSE_CAST_COLSUM", "1") != "0"      # A/B switch


def _fusable_cast_colsum(dy, cd, mul):
    return (_FUSE_CAST_COLSUM and cd == torch.bfloat16 and torch.is_tensor(dy) and dy.dtype == torch.float32 and
            dy.dim() == 2 and dy.stride(1) == 1 and dy.shape[1] % 8 == 0 and dy.stride(0) % 4 == 0 and
            dy.data_ptr() % 16 == 0 and
            (mul is None or (mul.stride(1) == 1 and mul.stride(0) % 4 == 0 and mul.data_ptr() % 16 == 0)))


def linear_bwd(dy, x, wt, cd, w_grad=None, b_grad=None, need_dx=True, dx_residual=None, mul=None, mul_mode=0,
               mul_scale=1.0):
    """Backward of y = x . W^T + b.
    dy: fp32 [M, N] (upstream gradient, multiplied element-wise by `mul` per mul_mode before use);
    x: the forward's A operand (Operand or fp32 tensor) [M, K]; wt: Operand holding W^T [K, pad8(N)];
    w_grad [N, K] / b_grad [N]: fp32 accumulators (+=).  Returns dX [M, K] fp32 (+ dx_residual) or None."""
    M, N = dy.shape
    dx = None
    dy_op = None
    mn_wgrad = w_grad is not None and _MN and cd == torch.bfloat16 and (_mn_ready(x) or not isinstance(x, Operand))
    if (need_dx or mn_wgrad) and b_grad is not None and _fusable_cast_colsum(dy, cd, mul):
        # one pass over dY: its bf16 GEMM operand and the bias gradient (ccx_convert_colsum)
        dy_op = Operand.empty((M, N), cd, dy.device)
        _lib.check(_lib.lib().ccx_convert_colsum(ptr(dy), dy.stride(0), ptr(mul), mul.stride(0) if mul is not None else 0,
                                                 mul_mode, mul_scale, ptr(dy_op.hi), N, ptr(b_grad), M, N,
                                                 _lib.stream_ptr()), "convert_colsum")
        b_grad = None
    if need_dx:
        if dy_op is None:
            dy_op = to_operand(dy, cd, mul, mul_mode, mul_scale)
        dx = linear_dgrad(dy_op, wt, residual=dx_residual, n=N)
    if w_grad is not None:
        x_op = x if _mn_ready(x) else None
        if _MN and cd == torch.bfloat16 and (x_op is not None or not isinstance(x, Operand)):
            # dW += dY^T . X on the row-major bf16 dY and X themselves (both read MN-major, contraction over rows)
            if dy_op is None:
                dy_op = to_operand(dy, cd, mul, mul_mode, mul_scale)
            if x_op is None:
                x_op = to_operand(x, cd)
            _lib.linear(dy_op, x_op, residual=w_grad, out=w_grad, a_mn=True, w_mn=True)
        else:
            dy_t = to_operand(dy, cd, mul, mul_mode, mul_scale, transpose=True)      # [N, Mp]
            x_t = to_operand(x, cd, transpose=True)                                   # [K, Mp]
            _lib.linear(dy_t, x_t, residual=w_grad, out=w_grad)
    if b_grad is not None:
        colsum_acc(dy, b_grad, mul, mul_mode, mul_scale)
    return dx


def ln_bwd(dy, x_in, gamma, dgamma, dbeta, eps, merge_hw=None):
    """merge_hw=(H, W): dy is in the 2x2 patch-merged layout of the downsample LayerNorm (see ccx_ln_bwd)."""
    M, C = x_in.shape
    dx = torch.empty_like(x_in)
    mg, H, W = (0, 1, 1) if merge_hw is None else (1, merge_hw[0], merge_hw[1])
    _lib.check(_lib.lib().ccx_ln_bwd(ptr(dy), ptr(x_in), ptr(gamma), ptr(dx), ptr(dgamma), ptr(dbeta), M, C, eps,
                                     mg, H, W, _lib.stream_ptr()), "ln_bwd")
    return dx


def weight_t(w, cd, op=None):
    """W [N, K] fp32 parameter -> what the dgrad GEMM dX = dY . W takes as its B operand: in bf16 the plain [N, K]
    operand read MN-major (`op`: the forward's own bf16 copy of W if the caller has it — then nothing is launched),
    otherwise the transposed Operand W^T [K, pad8(N)]."""
    if _MN and cd == torch.bfloat16:
        if op is not None and _mn_ready(op) and tuple(op.hi.shape) == tuple(w.shape):
            return MnWeight(op)
        if w.shape[1] % 8 == 0:
            return MnWeight(to_operand(w.detach(), cd))
    return to_operand(w.detach(), cd, transpose=True)
