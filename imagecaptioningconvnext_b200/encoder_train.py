"""Encoder forward WITH autograd for ``Encoder.fine_tune(True, startingLayer=7)`` — the reference's default
fine-tuning extent (trainMultiGPU.py:68: children()[7:] = the three C=1024 CNBlocks at 8x8; models/encoder.py:29-34).

Children [0, 7) run frozen through ``ccx_encoder_run``; the trainable blocks run op by op so their inputs can be
kept, and the backward is explicit (libccx launches only):
  layer_scale / stochastic-depth / residual -> second Linear (dgrad + un-scaled wgrad, layer_scale gradient derived
  from it without recomputing the branch) -> GELU' (pre-activation recomputed by one GEMM) -> first Linear ->
  LayerNorm backward (its input recomputed by the conv kernel in plain mode) -> depthwise conv data gradient (same
  TMA/cluster kernel with flipped taps) and filter gradient.
Fine-tuning from an earlier child (downsample / stem backward) is not built yet and raises.
"""
import torch

from . import _lib
from ._lib import Operand, ptr
from .train_ops import colsum_acc, linear_bwd, ln_bwd, to_operand, weight_t


def _first_trainable_child(enc):
    for i, c in enumerate(enc.convnext.children()):
        if any(p.requires_grad for p in c.parameters()):
            return i
    return 8


class _EncoderTail(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, x7, noise, *params):
        L, st = _lib.lib(), _lib.stream_ptr()
        cd = enc.compute_dtype
        code = _lib.dt_code(cd)
        enc.prepared()
        B, H, W, C = x7.shape
        M = B * H * W
        blocks = list(enc.convnext[7])
        nb0 = sum(len(enc.convnext[i]) for i in (1, 3, 5))       # index of the first stage-4 block
        saved = []
        x = x7
        for i, blk in enumerate(blocks):
            ops = enc._block_ops[nb0 + i]
            y_op = Operand.empty((M, C), cd, x.device)
            _lib.check(L.ccx_dwconv7_ln(ptr(x), ptr(ops["dw_w"]), ptr(blk.block[0].bias.detach()),
                                        ptr(blk.block[2].weight.detach()), ptr(blk.block[2].bias.detach()),
                                        ptr(y_op.hi), y_op.lo_ptr, B, H, W, C, 1e-6, code, st), "dwconv7_ln")
            if cd == torch.bfloat16:
                h_op = Operand(_lib.linear(y_op, ops["w1"], bias=blk.block[3].bias.detach(), act=_lib.ACT_GELU,
                                           out_dtype=torch.bfloat16), None, torch.bfloat16)
            else:
                h_op = _lib.linear(y_op, ops["w1"], bias=blk.block[3].bias.detach(), act=_lib.ACT_GELU, split=True)
            rs = None if noise is None else noise[nb0 + i]
            x_out = _lib.linear(h_op, ops["w2"], bias=blk.block[5].bias.detach(),
                                colscale=blk.layer_scale.detach().view(C), rowscale=rs, rows_per_group=H * W,
                                residual=x.view(M, C)).view(B, H, W, C)
            saved.append((x, y_op, h_op, rs))
            x = x_out
        out = enc._pool(x)
        ctx.enc, ctx.saved, ctx.dims, ctx.nb0 = enc, saved, (B, H, W, C), nb0
        ctx.x7_needs_grad = x7.requires_grad
        return out

    @staticmethod
    def backward(ctx, dpooled):
        enc = ctx.enc
        L, st = _lib.lib(), _lib.stream_ptr()
        cd = enc.compute_dtype
        B, H, W, C = ctx.dims
        M, K4 = B * H * W, 4 * C
        dev = dpooled.device
        f32 = dict(dtype=torch.float32, device=dev)
        blocks = list(enc.convnext[7])
        names = [n for n, _ in enc.convnext[7].named_parameters()]
        grads = {n: torch.zeros_like(p, dtype=torch.float32) for n, p in enc.convnext[7].named_parameters()}
        dout = torch.empty((M, C), **f32)
        _lib.check(L.ccx_avgpool_nhwc_bwd(ptr(dpooled.contiguous()), ptr(dout), B, H, W, C, enc.enc_image_size, st),
                   "avgpool_bwd")
        for i in reversed(range(len(blocks))):
            blk = blocks[i]
            x_in, y_op, h_op, rs = ctx.saved[i]
            ops = enc._block_ops[ctx.nb0 + i]
            gamma = blk.layer_scale.detach().view(C)
            W1, W2 = blk.block[3].weight.detach(), blk.block[5].weight.detach()
            pre_n = f"{i}."
            # layer_scale * stochastic depth
            dz = torch.empty((M, C), **f32)
            _lib.check(L.ccx_scale_rows_cols(ptr(dout), ptr(gamma), ptr(rs), H * W, ptr(dz), M, C, st), "scale")
            if rs is not None:
                doutp = torch.empty((M, C), **f32)
                _lib.check(L.ccx_scale_rows_cols(ptr(dout), None, ptr(rs), H * W, ptr(doutp), M, C, st), "scale")
            else:
                doutp = dout
            # second Linear: dgrad, un-scaled wgrad G, and the layer_scale / W2 / b2 gradients from it
            dh = _lib.linear(to_operand(dz, cd), weight_t(W2, cd), k=C)                       # [M, 4C]
            G = _lib.linear(to_operand(doutp, cd, transpose=True), to_operand(h_op, cd, transpose=True))  # [C, 4C]
            s = torch.zeros((C,), **f32)
            colsum_acc(doutp, s)
            _lib.check(L.ccx_cnblock_param_grads(ptr(G), ptr(W2), ptr(blk.block[5].bias.detach()), ptr(gamma), ptr(s),
                                                 ptr(grads[pre_n + "block.5.weight"]),
                                                 ptr(grads[pre_n + "layer_scale"]),
                                                 ptr(grads[pre_n + "block.5.bias"]), C, K4, st), "cnblock_param_grads")
            # GELU' on the recomputed pre-activation, first Linear
            pre = _lib.linear(y_op, ops["w1"], bias=blk.block[3].bias.detach())
            _lib.check(L.ccx_gelu_bwd(ptr(pre), ptr(dh), M * K4, st), "gelu_bwd")
            dy = linear_bwd(dh, y_op, weight_t(W1, cd), cd, grads[pre_n + "block.3.weight"],
                            grads[pre_n + "block.3.bias"])
            # LayerNorm backward on the recomputed conv output
            u = torch.empty((M, C), **f32)
            _lib.check(L.ccx_dwconv7_plain(ptr(x_in), ptr(ops["dw_w"]), ptr(blk.block[0].bias.detach()), None, ptr(u),
                                           B, H, W, C, st), "dwconv7_plain")
            du = ln_bwd(dy, u, blk.block[2].weight.detach(), grads[pre_n + "block.2.weight"],
                        grads[pre_n + "block.2.bias"], 1e-6)
            # depthwise conv: bias, filter and data gradients; residual add fused into the data-gradient launch
            colsum_acc(du, grads[pre_n + "block.0.bias"])
            dw = torch.zeros((49, C), **f32)
            _lib.check(L.ccx_dwconv7_wgrad(ptr(x_in), ptr(du), ptr(dw), B, H, W, C, st), "dwconv7_wgrad")
            grads[pre_n + "block.0.weight"] = dw.t().reshape(C, 1, 7, 7).contiguous()     # tap-major -> (C,1,7,7)
            w_flip = ops["dw_w"].flip(0).contiguous()                                      # tiny (49 x C) re-layout
            dprev = torch.empty((M, C), **f32)
            _lib.check(L.ccx_dwconv7_plain(ptr(du), ptr(w_flip), None, ptr(dout), ptr(dprev), B, H, W, C, st),
                       "dwconv7_dgrad")
            dout = dprev
        dx7 = dout.view(B, H, W, C) if ctx.x7_needs_grad else None
        return (None, dx7, None) + tuple(grads[n] for n in names)


def encoder_features_with_grad(enc, images, noise):
    """Returns the POOLED features (B,s,s,C); Encoder.forward must not pool again."""
    first = _first_trainable_child(enc)
    if first < 7:
        raise NotImplementedError(
            f"fine-tuning from child {first} needs the downsample/stem backward kernels, which are not built yet; "
            "the reference's default (trainMultiGPU.py --startingLayer 7) is supported")
    B, _, H, W = images.shape
    with torch.no_grad():
        x7 = enc.run_children(images, 0, 7, noise)
    params = [p for _, p in enc.convnext[7].named_parameters()]
    return _EncoderTail.apply(enc, x7, noise, *params)
