"""Encoder forward WITH autograd for ``Encoder.fine_tune(True, startingLayer)`` (models/encoder.py:29-34), for any
``startingLayer`` in 1..7 (trainMultiGPU.py:68 defaults to 7, train.py:63 to 5; 0 = stem is not built and raises).

Children [0, startingLayer) run frozen through ``ccx_encoder_run``; the trainable children run op by op so their
inputs can be kept, and the backward is explicit (libccx launches only):
  CNBlock:   layer_scale / stochastic-depth / residual -> second Linear (dgrad + un-scaled wgrad; the layer_scale
             gradient is derived from it without recomputing the branch) -> GELU' (pre-activation recomputed by one
             GEMM) -> first Linear -> LayerNorm backward (its input recomputed by the conv kernel in plain mode) ->
             depthwise-conv data gradient (the same TMA/cluster kernel with flipped taps) and filter gradient;
  downsample: patch-merge GEMM dgrad/wgrad -> LayerNorm2d backward reading dy in the merged layout.
"""
import torch

from . import _lib
from ._lib import Operand, ptr
from .encoder import DIMS
from .train_ops import colsum_acc, linear_bwd, ln_bwd, to_operand, weight_t, zero_grads_like


def _first_trainable_child(enc):
    for i, c in enumerate(enc.convnext.children()):
        if any(p.requires_grad for p in c.parameters()):
            return i
    return 8


def _block_index0(enc, child):
    """Global index (stage-major) of the first CNBlock of stage child `child` (1, 3, 5 or 7)."""
    return sum(len(enc.convnext[i]) for i in (1, 3, 5, 7) if i < child)


class _EncoderTail(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, x_in, first, noise, *params):
        L, st = _lib.lib(), _lib.stream_ptr()
        cd = enc.compute_dtype
        code = _lib.dt_code(cd)
        enc.prepared()
        saved = []          # one entry per op, in execution order
        x = x_in
        for child in range(first, 8):
            B, H, W, C = x.shape
            M = B * H * W
            if child % 2 == 0:
                # downsample: LayerNorm2d + 2x2/s2 conv as patch-merge GEMM (convnext.py:146-151)
                mod = enc.convnext[child]
                y_op = Operand.empty((M // 4, 4 * C), cd, x.device)
                _lib.check(L.ccx_ln_rows(ptr(x), ptr(mod[0].weight.detach()), ptr(mod[0].bias.detach()), ptr(y_op.hi),
                                         y_op.lo_ptr, None, M, C, 1e-6, code, 1, H, W, st), "ln_rows")
                Cout = DIMS[child // 2]
                x_out = _lib.linear(y_op, enc._down_ops[child], bias=mod[1].bias.detach()).view(B, H // 2, W // 2, Cout)
                saved.append(("down", child, x, y_op))
                x = x_out
                continue
            nb0 = _block_index0(enc, child)
            for i, blk in enumerate(enc.convnext[child]):
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
                saved.append(("block", (child, i, nb0 + i), x, y_op, h_op, rs))
                x = x_out
        out = enc._pool(x)
        ctx.enc, ctx.saved, ctx.first, ctx.feat_shape = enc, saved, first, x.shape
        ctx.x_needs_grad = x_in.requires_grad
        return out

    @staticmethod
    def backward(ctx, dpooled):
        enc = ctx.enc
        L, st = _lib.lib(), _lib.stream_ptr()
        cd = enc.compute_dtype
        dev = dpooled.device
        f32 = dict(dtype=torch.float32, device=dev)
        tail = [(f"{c}.{n}", p) for c in range(ctx.first, 8) for n, p in enc.convnext[c].named_parameters()]
        grads = zero_grads_like(tail)
        for n, p in tail:                      # frozen parameters inside the tail (not a reference use case)
            grads.setdefault(n, torch.zeros_like(p, dtype=torch.float32))
        B, H, W, C = ctx.feat_shape
        dout = torch.empty((B * H * W, C), **f32)
        _lib.check(L.ccx_avgpool_nhwc_bwd(ptr(dpooled.contiguous()), ptr(dout), B, H, W, C, enc.enc_image_size, st),
                   "avgpool_bwd")
        for entry in reversed(ctx.saved):
            if entry[0] == "down":
                _, child, x_in, y_op = entry
                B, H, W, C = x_in.shape
                M = B * H * W
                mod = enc.convnext[child]
                Cout = mod[1].weight.shape[0]
                gw = torch.zeros((Cout, 4 * C), **f32)
                w_perm = mod[1].weight.detach().permute(0, 2, 3, 1).reshape(Cout, 4 * C)       # (Cout, kh, kw, Cin)
                dmerged = linear_bwd(dout, y_op, weight_t(w_perm, cd), cd, gw, grads[f"{child}.1.bias"])
                grads[f"{child}.1.weight"] = gw.view(Cout, 2, 2, C).permute(0, 3, 1, 2).contiguous()
                dout = ln_bwd(dmerged, x_in.view(M, C), mod[0].weight.detach(), grads[f"{child}.0.weight"],
                              grads[f"{child}.0.bias"], 1e-6, merge_hw=(H, W))
                continue
            _, (child, i, gi), x_in, y_op, h_op, rs = entry
            B, H, W, C = x_in.shape
            M, K4 = B * H * W, 4 * C
            blk = enc.convnext[child][i]
            ops = enc._block_ops[gi]
            gamma = blk.layer_scale.detach().view(C)
            W1, W2 = blk.block[3].weight.detach(), blk.block[5].weight.detach()
            pre_n = f"{child}.{i}."
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
            del pre, dh
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
        dx = None
        if ctx.x_needs_grad:
            dx = dout.view(ctx.saved[0][2].shape)
        return (None, dx, None, None) + tuple(grads[n] for n, _ in tail)


def encoder_features_with_grad(enc, images, noise):
    """Returns the POOLED features (B,s,s,C); Encoder.forward must not pool again."""
    first = _first_trainable_child(enc)
    if first < 1:
        raise NotImplementedError("fine-tuning the stem (startingLayer=0) is not built: its backward kernel is missing; "
                                  "startingLayer 1..7 are supported (reference defaults: 7 and 5)")
    B, _, H, W = images.shape
    with torch.no_grad():
        x = enc.run_children(images, 0, first, noise)
    params = [p for c in range(first, 8) for _, p in enc.convnext[c].named_parameters()]
    return _EncoderTail.apply(enc, x, first, noise, *params)
