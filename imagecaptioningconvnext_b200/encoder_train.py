"""Encoder forward WITH autograd for ``Encoder.fine_tune(True, startingLayer)`` (models/encoder.py:29-34), for any
``startingLayer`` in 1..7 (trainMultiGPU.py:68 defaults to 7, train.py:63 to 5; 0 = stem is not built and raises).

Children [0, startingLayer) run frozen through ``ccx_encoder_run``; the trainable children run op by op so their
inputs can be kept — one autograd node per CNBlock / downsample, so that DistributedDataParallel can all-reduce a
block's gradients while the previous block still back-propagates — and the backward is explicit (libccx launches
only):
  CNBlock:   layer_scale / stochastic-depth / residual -> second Linear (dgrad + un-scaled wgrad; the layer_scale
             gradient is derived from it without recomputing the branch) -> GELU' (pre-activation recomputed by one
             GEMM) -> first Linear -> LayerNorm backward (its input recomputed by the conv kernel in plain mode) ->
             depthwise-conv data gradient (the same TMA/cluster kernel with flipped taps) and filter gradient;
  downsample: patch-merge GEMM dgrad/wgrad -> LayerNorm2d backward reading dy in the merged layout.
"""
import os

import torch

from . import _lib
from ._host import any_requires_grad, named_params
from ._lib import Operand, ptr
from .encoder import DIMS
from .train_ops import (colsum_acc, deliver, grads_out, linear_bwd, linear_dgrad, ln_bwd, mn_operands, to_operand,
                        weight_t, zero_grads_like)


def _first_trainable_child(enc):
    for i, c in enumerate(enc.convnext.children()):
        if any_requires_grad(c):
            return i
    return 8


def _block_index0(enc, child):
    """Global index (stage-major) of the first CNBlock of stage child `child` (1, 3, 5 or 7)."""
    return sum(len(enc.convnext[i]) for i in (1, 3, 5, 7) if i < child)


def _named(mod):
    return named_params(mod)


_FUSED_GELU_BWD = os.environ.get("CCX_FUSED_GELU_BWD", "0") != "0"     # 1: GELU backward inside the re-computing GEMM (measured: no gain)


class _BlockFn(torch.autograd.Function):
    """One CNBlock (torchvision/models/convnext.py:51-67) as its own autograd node: under DistributedDataParallel the
    gradients of block k are handed to the reducer as soon as block k's backward is done, so their all-reduce
    overlaps the backward of block k-1 instead of waiting for the whole stage."""

    @staticmethod
    def forward(ctx, enc, child, i, gi, x, rs, *params):
        L, st = _lib.lib(), _lib.stream_ptr()
        cd = enc.compute_dtype
        code = _lib.dt_code(cd)
        B, H, W, C = x.shape
        M = B * H * W
        blk = enc.convnext[child][i]
        ops = enc._block_ops[gi]
        y_op = Operand.empty((M, C), cd, x.device)
        _lib.check(L.ccx_dwconv7_ln(ptr(x), ptr(ops["dw_w"]), ptr(blk.block[0].bias.detach()),
                                    ptr(blk.block[2].weight.detach()), ptr(blk.block[2].bias.detach()),
                                    ptr(y_op.hi), y_op.lo_ptr, B, H, W, C, 1e-6, code, st), "dwconv7_ln")
        if cd == torch.bfloat16:
            h_op = Operand(_lib.linear(y_op, ops["w1"], bias=blk.block[3].bias.detach(), act=_lib.ACT_GELU,
                                       out_dtype=torch.bfloat16), None, torch.bfloat16)
        else:
            h_op = _lib.linear(y_op, ops["w1"], bias=blk.block[3].bias.detach(), act=_lib.ACT_GELU, split=True)
        x_out = _lib.linear(h_op, ops["w2"], bias=blk.block[5].bias.detach(),
                            colscale=blk.layer_scale.detach().view(C), rowscale=rs, rows_per_group=H * W,
                            residual=x.view(M, C)).view(B, H, W, C)
        ctx.enc, ctx.where, ctx.saved = enc, (child, i, gi), (x, y_op, h_op, rs)
        return x_out

    @staticmethod
    def backward(ctx, dout):
        enc = ctx.enc
        L, st = _lib.lib(), _lib.stream_ptr()
        cd = enc.compute_dtype
        child, i, gi = ctx.where
        x_in, y_op, h_op, rs = ctx.saved
        B, H, W, C = x_in.shape
        M, K4 = B * H * W, 4 * C
        dev = dout.device
        f32 = dict(dtype=torch.float32, device=dev)
        blk = enc.convnext[child][i]
        ops = enc._block_ops[gi]
        named = _named(blk)
        grads, (s, dw) = zero_grads_like(named, extra_shapes=[(C,), (49, C)])
        for n, p in named:                     # frozen parameters inside the tail (not a reference use case)
            if n not in grads:                 # (never `setdefault(n, zeros_like(p))`: that fills a buffer per parameter)
                grads[n] = torch.zeros_like(p, dtype=torch.float32)
        dout = dout.contiguous().view(M, C)
        gamma = blk.layer_scale.detach().view(C)
        W1, W2 = blk.block[3].weight.detach(), blk.block[5].weight.detach()
        # layer_scale * stochastic depth
        dz = torch.empty((M, C), **f32)
        _lib.check(L.ccx_scale_rows_cols(ptr(dout), ptr(gamma), ptr(rs), H * W, ptr(dz), M, C, st), "scale")
        if rs is not None:
            doutp = torch.empty((M, C), **f32)
            _lib.check(L.ccx_scale_rows_cols(ptr(dout), None, ptr(rs), H * W, ptr(doutp), M, C, st), "scale")
        else:
            doutp = dout
        # second Linear: dgrad, un-scaled wgrad G, and the layer_scale / W2 / b2 gradients from it
        dh = linear_dgrad(to_operand(dz, cd), weight_t(W2, cd, ops["w2"]), n=C)             # [M, 4C]
        if mn_operands(cd, h_op):
            # G = doutp^T . h on the row-major operands themselves (tensor core reads both transposed in place)
            G = _lib.linear(to_operand(doutp, cd), h_op, a_mn=True, w_mn=True)                       # [C, 4C]
        else:
            G = _lib.linear(to_operand(doutp, cd, transpose=True), to_operand(h_op, cd, transpose=True))
        colsum_acc(doutp, s)
        _lib.check(L.ccx_cnblock_param_grads(ptr(G), ptr(W2), ptr(blk.block[5].bias.detach()), ptr(gamma), ptr(s),
                                             ptr(grads["block.5.weight"]), ptr(grads["layer_scale"]),
                                             ptr(grads["block.5.bias"]), C, K4, st), "cnblock_param_grads")
        # GELU' on the recomputed pre-activation, first Linear
        if _FUSED_GELU_BWD:
            # dh *= GELU'(y . W1^T + b1): the pre-activation is recomputed and consumed inside the GEMM's epilogue
            _lib.linear(y_op, ops["w1"], bias=blk.block[3].bias.detach(), act=_lib.ACT_GELU_GRAD, residual=dh, out=dh,
                        res_mul=True)
        else:
            pre = _lib.linear(y_op, ops["w1"], bias=blk.block[3].bias.detach())
            _lib.check(L.ccx_gelu_bwd(ptr(pre), ptr(dh), M * K4, st), "gelu_bwd")
            del pre
        dy = linear_bwd(dh, y_op, weight_t(W1, cd, ops["w1"]), cd, grads["block.3.weight"], grads["block.3.bias"])
        del dh
        # LayerNorm backward on the recomputed conv output
        u = torch.empty((M, C), **f32)
        _lib.check(L.ccx_dwconv7_plain(ptr(x_in), ptr(ops["dw_w"]), ptr(blk.block[0].bias.detach()), None, ptr(u),
                                       B, H, W, C, st), "dwconv7_plain")
        du = ln_bwd(dy, u, blk.block[2].weight.detach(), grads["block.2.weight"], grads["block.2.bias"], 1e-6)
        # depthwise conv: bias, filter and data gradients; residual add fused into the data-gradient launch
        colsum_acc(du, grads["block.0.bias"])
        _lib.check(L.ccx_dwconv7_wgrad(ptr(x_in), ptr(du), ptr(dw), B, H, W, C, st), "dwconv7_wgrad")
        deliver(grads, "block.0.weight", blk.block[0].weight, dw.t().reshape(C, 1, 7, 7))     # tap-major -> (C,1,7,7)
        dprev = None
        if ctx.needs_input_grad[4]:
            w_flip = ops["dw_w"].flip(0).contiguous()                                  # tiny (49 x C) re-layout
            dprev = torch.empty((M, C), **f32)
            _lib.check(L.ccx_dwconv7_plain(ptr(du), ptr(w_flip), None, ptr(dout), ptr(dprev), B, H, W, C, st),
                       "dwconv7_dgrad")
            dprev = dprev.view(B, H, W, C)
        return (None, None, None, None, dprev, None) + grads_out(grads, named)


class _DownFn(torch.autograd.Function):
    """Downsample child: LayerNorm2d + 2x2/s2 conv as patch-merge GEMM (convnext.py:146-151)."""

    @staticmethod
    def forward(ctx, enc, child, x, *params):
        L, st = _lib.lib(), _lib.stream_ptr()
        cd = enc.compute_dtype
        B, H, W, C = x.shape
        M = B * H * W
        mod = enc.convnext[child]
        y_op = Operand.empty((M // 4, 4 * C), cd, x.device)
        _lib.check(L.ccx_ln_rows(ptr(x), ptr(mod[0].weight.detach()), ptr(mod[0].bias.detach()), ptr(y_op.hi),
                                 y_op.lo_ptr, None, M, C, 1e-6, _lib.dt_code(cd), 1, H, W, st), "ln_rows")
        Cout = DIMS[child // 2]
        x_out = _lib.linear(y_op, enc._down_ops[child], bias=mod[1].bias.detach()).view(B, H // 2, W // 2, Cout)
        ctx.enc, ctx.child, ctx.saved = enc, child, (x, y_op)
        return x_out

    @staticmethod
    def backward(ctx, dout):
        enc, child = ctx.enc, ctx.child
        cd = enc.compute_dtype
        x_in, y_op = ctx.saved
        B, H, W, C = x_in.shape
        M = B * H * W
        mod = enc.convnext[child]
        named = _named(mod)
        grads = zero_grads_like(named)
        for n, p in named:
            if n not in grads:
                grads[n] = torch.zeros_like(p, dtype=torch.float32)
        Cout = mod[1].weight.shape[0]
        dout = dout.contiguous().view(M // 4, Cout)
        gw = torch.zeros((Cout, 4 * C), dtype=torch.float32, device=dout.device)
        w_perm = mod[1].weight.detach().permute(0, 2, 3, 1).reshape(Cout, 4 * C)       # (Cout, kh, kw, Cin)
        dmerged = linear_bwd(dout, y_op, weight_t(w_perm, cd), cd, gw, grads["1.bias"])
        deliver(grads, "1.weight", mod[1].weight, gw.view(Cout, 2, 2, C).permute(0, 3, 1, 2))
        dx = ln_bwd(dmerged, x_in.view(M, C), mod[0].weight.detach(), grads["0.weight"], grads["0.bias"], 1e-6,
                    merge_hw=(H, W))
        dx = dx.view(B, H, W, C) if ctx.needs_input_grad[2] else None
        return (None, None, dx) + grads_out(grads, named)


class _PoolFn(torch.autograd.Function):
    """AdaptiveAvgPool2d + permute (models/encoder.py:25-26) on the NHWC stream."""

    @staticmethod
    def forward(ctx, enc, x):
        ctx.enc, ctx.shape = enc, x.shape
        return enc._pool(x)

    @staticmethod
    def backward(ctx, dpooled):
        B, H, W, C = ctx.shape
        dx = torch.empty((B, H, W, C), dtype=torch.float32, device=dpooled.device)
        _lib.check(_lib.lib().ccx_avgpool_nhwc_bwd(ptr(dpooled.contiguous()), ptr(dx), B, H, W, C,
                                                   ctx.enc.enc_image_size, _lib.stream_ptr()), "avgpool_bwd")
        return None, dx


def encoder_features_with_grad(enc, images, noise, begin_child=0, image_hw=None):
    """Returns the POOLED features (B,s,s,C); Encoder.forward must not pool again.
    begin_child=1: `images` is already the stem's NHWC output (uint8 input path), image_hw the original size."""
    first = _first_trainable_child(enc)
    if first < 1:
        raise NotImplementedError("fine-tuning the stem (startingLayer=0) is not built: its backward kernel is missing; "
                                  "startingLayer 1..7 are supported (reference defaults: 7 and 5)")
    with torch.no_grad():
        x = images if begin_child >= first else enc.run_children(images, begin_child, first, noise, image_hw=image_hw)
    pre = getattr(enc, "_before_trainable", None)
    if pre is not None:          # CapturedTrainStep: the refreshed weight copies come from its side stream
        pre()
    enc.prepared()
    # optional callback hook(child, i): the gradients of block i of child `child` (i = None: a downsample child) are
    # complete — fired from a tensor hook on the unit's INPUT, i.e. when its backward node has run and (AccumulateGrad
    # nodes have the highest priority in the autograd engine) its parameter gradients have been accumulated.
    # CapturedTrainStep uses it to all-reduce a unit's gradient slice while the previous unit still back-propagates.
    hook = getattr(enc, "_unit_grads_ready", None)

    def watch(x, child, i):
        if hook is not None and x.requires_grad:
            x.register_hook(lambda g, child=child, i=i: hook(child, i))
    for child in range(first, 8):
        if child % 2 == 0:
            watch(x, child, None)
            x = _DownFn.apply(enc, child, x, *[p for _, p in _named(enc.convnext[child])])
            continue
        nb0 = _block_index0(enc, child)
        for i, blk in enumerate(enc.convnext[child]):
            rs = None if noise is None else noise[nb0 + i]
            watch(x, child, i)
            x = _BlockFn.apply(enc, child, i, nb0 + i, x, rs, *[p for _, p in _named(blk)])
    return _PoolFn.apply(enc, x)
