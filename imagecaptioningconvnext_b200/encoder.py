"""``Encoder`` — drop-in for the reference's ``models/encoder.py:14-34`` on hand-written sm_100a kernels.

Same constructor (``encoded_image_size=7``), same ``forward(images) -> (B, s, s, 1024)``, same
``fine_tune(fine_tune, startingLayer)``, same ``state_dict`` keys (``convnext.<child>...``: the 344 tensors of
``torchvision.models.convnext_base().features``), so ``load_state_dict(checkpoint['encoder'])`` works unchanged.
Differences a caller can see:
  * weights are NOT downloaded (reference: models/encoder.py:18 pulls IMAGENET1K_V1); initialisation follows
    torchvision's (trunc_normal std 0.02, zero bias, layer_scale 1e-6) and real weights come from load_state_dict;
  * the output is a contiguous NHWC tensor instead of a permuted NCHW view (same shape/values; ``.view(B,-1,C)``
    as the decoders do at models/decoder.py:75 is legal on both);
  * extra ctor kwarg ``compute_dtype`` (float32 = 3xTF32 tensor-core path for the 1e-3 parity bar, bfloat16 =
    fast path); input/output stay fp32 either way.
There is no eager/CPU path: tensors must be CUDA and libccx.so must be built.
"""
import ctypes as C

import torch
from torch import nn

from . import _lib
from ._host import any_requires_grad, no_gc, params_of
from ._lib import Operand

DEPTHS = (3, 3, 27, 3)
DIMS = (128, 256, 512, 1024)
IMAGENET_MEAN = (0.485, 0.456, 0.406)   # the Normalize constants of train.py:152-153 / caption.py:62
IMAGENET_STD = (0.229, 0.224, 0.225)
SD_PROB = 0.5  # torchvision convnext_base default stochastic_depth_prob (tv:models/convnext.py:356-382)


# ---- parameter containers that reproduce torchvision's module tree (and therefore its state_dict keys) ----
class _Affine(nn.Module):
    """weight/bias holder standing in for LayerNorm / LayerNorm2d."""

    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))


class _Kernel(nn.Module):
    """weight/bias holder standing in for Conv2d / Linear (torchvision init: trunc_normal std .02, zero bias)."""

    def __init__(self, *shape):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(*shape))
        self.bias = nn.Parameter(torch.zeros(shape[0]))
        nn.init.trunc_normal_(self.weight, std=0.02)


class _Hole(nn.Module):
    """parameter-less slot (Permute / GELU in torchvision's Sequential) so the indices 0,2,3,5 line up."""


class _CNBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.layer_scale = nn.Parameter(torch.ones(c, 1, 1) * 1e-6)  # registered first, as torchvision does
        self.block = nn.Sequential(_Kernel(c, 1, 7, 7), _Hole(), _Affine(c), _Kernel(4 * c, c), _Hole(),
                                   _Kernel(c, 4 * c), _Hole())


def _build_features():
    children = [nn.Sequential(_Kernel(128, 3, 4, 4), _Affine(128))]
    for s, (d, c) in enumerate(zip(DEPTHS, DIMS)):
        if s > 0:
            children.append(nn.Sequential(_Affine(DIMS[s - 1]), _Kernel(c, DIMS[s - 1], 2, 2)))
        children.append(nn.Sequential(*[_CNBlock(c) for _ in range(d)]))
    return nn.Sequential(*children)


def stochastic_depth_probs():
    """tv:models/convnext.py:137-141: p_i = 0.5 * i / 35 for block i of 36."""
    n = sum(DEPTHS)
    return [SD_PROB * i / (n - 1.0) for i in range(n)]


class Encoder(nn.Module):
    def __init__(self, encoded_image_size=7, compute_dtype=torch.float32):
        super().__init__()
        self.enc_image_size = encoded_image_size
        self.compute_dtype = compute_dtype
        _lib.dt_code(compute_dtype)
        self.convnext = _build_features()
        self.adaptive_pool = nn.AdaptiveAvgPool2d((encoded_image_size, encoded_image_size))  # attribute parity only
        self._prep_key = None
        self._prep = None      # (EncoderWeights struct, list of tensors kept alive)
        self._ws = None
        self.sd_noise = None   # tests may inject a (36, B) tensor of stochastic-depth row factors
        self._graphs = None    # {input shape/dtype: (CUDAGraph, static_in, static_out, weights key)} when enabled
        self.fine_tune()

    def enable_cuda_graph(self, enabled=True):
        """Inference-only option: capture the whole forward (117 launches, TMA descriptors included) into one CUDA
        graph per input shape and replay it, so a slow or busy host cannot starve the GPU.  Eval mode / no-grad calls
        only; train-mode or grad-enabled calls keep the eager path.  Results are bit-identical to the eager path."""
        self._graphs = {} if enabled else None
        return self

    # ---- reference API -------------------------------------------------------------------------------------
    def fine_tune(self, fine_tune=True, startingLayer=7):
        """models/encoder.py:29-34."""
        for p in self.convnext.parameters():
            p.requires_grad = False
        for c in list(self.convnext.children())[startingLayer:]:
            for p in c.parameters():
                p.requires_grad = fine_tune
        self._first_trainable = startingLayer if fine_tune else 8

    def forward(self, images):
        """models/encoder.py:23-27: (B,3,H,W) fp32 -> (B,s,s,1024) fp32."""
        _lib.require_cuda(images, "images")
        if images.dim() != 4 or images.shape[1] != 3 or images.dtype not in (torch.float32, torch.uint8):
            raise ValueError(f"images must be float32 (or raw uint8) (B,3,H,W); got {tuple(images.shape)} {images.dtype}")
        B, _, H, W = images.shape
        if H % 32 or W % 32:
            raise ValueError("image height/width must be multiples of 32 (ConvNeXt total stride)")
        images = images.contiguous()
        needs_grad = torch.is_grad_enabled() and any_requires_grad(self.convnext)
        noise = self._stochastic_depth_noise(B, images.device)
        if self._graphs is not None and not needs_grad and noise is None:
            return self._forward_graphed(images)
        if images.dtype == torch.uint8:
            # raw dataset pixels: /255 and Normalize(mean, std) (dataLoader.py:43-45) are fused into the stem kernel
            x1 = self._stem_u8(images)
            if needs_grad:    # the stem (child 0) is never trainable (encoder_train.py), so it runs frozen here too
                from .encoder_train import encoder_features_with_grad
                return encoder_features_with_grad(self, x1, noise, begin_child=1, image_hw=(H, W))
            return self._pool(self.run_children(x1, 1, 8, noise, image_hw=(H, W)))
        if needs_grad:
            from .encoder_train import encoder_features_with_grad  # backward kernels live there
            return encoder_features_with_grad(self, images, noise)   # pooled inside the autograd Function
        return self._pool(self.run_children(images, 0, 8, noise))

    # ---- implementation ------------------------------------------------------------------------------------
    def _stochastic_depth_noise(self, B, device):
        """tv:ops/stochastic_depth.py:8-44 (mode 'row'): noise = bernoulli(1-p) / (1-p), one value per sample."""
        if not self.training:
            return None
        if self.sd_noise is not None:
            return self.sd_noise.to(device=device, dtype=torch.float32).contiguous()
        if getattr(self, "_sd_keep", None) is None or self._sd_keep.device != device:
            p = torch.tensor(stochastic_depth_probs(), device=device, dtype=torch.float32).view(-1, 1)
            self._sd_keep = 1.0 - p          # built once: torch.tensor(..., device=cuda) is a blocking copy
        keep = self._sd_keep
        return (torch.bernoulli(keep.expand(-1, B)) / keep).contiguous()

    def _forward_eager_nograd(self, images):
        B, _, H, W = images.shape
        if images.dtype == torch.uint8:
            return self._pool(self.run_children(self._stem_u8(images), 1, 8, None, image_hw=(H, W)))
        return self._pool(self.run_children(images, 0, 8, None))

    @torch.no_grad()
    def _forward_graphed(self, images):
        self.prepared()
        key = (tuple(images.shape), images.dtype, images.device)
        entry = self._graphs.get(key)
        if entry is None or entry[3] != self._prep_key[:3]:
            static_in = images.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._forward_eager_nograd(static_in)          # warm-up: kernel attributes, workspace, weight prep
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with no_gc(), torch.cuda.graph(g):
                static_out = self._forward_eager_nograd(static_in)
            entry = (g, static_in, static_out, self._prep_key[:3])
            self._graphs[key] = entry
        g, static_in, static_out, _ = entry
        static_in.copy_(images)
        g.replay()
        return static_out.clone()

    def _stem_u8(self, images_u8):
        B, _, H, W = images_u8.shape
        w = self.prepared()
        dev = images_u8.device
        if getattr(self, "_norm_consts", None) is None or self._norm_consts[0].device != dev:
            mean = torch.tensor(IMAGENET_MEAN, dtype=torch.float32, device=dev)
            inv_std = (1.0 / torch.tensor(IMAGENET_STD, dtype=torch.float64)).to(torch.float32).to(dev)
            self._norm_consts = (mean, inv_std)
        out = torch.empty((B, H // 4, W // 4, DIMS[0]), dtype=torch.float32, device=dev)
        _lib.check(_lib.lib().ccx_stem_ln_u8(images_u8.data_ptr(), self._norm_consts[0].data_ptr(),
                                             self._norm_consts[1].data_ptr(), w.stem_w, w.stem_b, w.stem_ln_g,
                                             w.stem_ln_b, out.data_ptr(), B, H, W, 1e-6, _lib.stream_ptr()),
                   "stem_ln_u8")
        return out

    def _pool(self, feat):
        B, h, w, Cc = feat.shape
        s = self.enc_image_size
        out = torch.empty((B, s, s, Cc), dtype=torch.float32, device=feat.device)
        _lib.check(_lib.lib().ccx_avgpool_nhwc(feat.data_ptr(), out.data_ptr(), B, h, w, Cc, s, _lib.stream_ptr()),
                   "avgpool")
        return out

    def prepared(self):
        """Kernel-side weight table (re-laid-out / bf16 or tf32-split copies of the fp32 master weights).
        Built once; afterwards a changed parameter (optimizer step, load_state_dict) is re-converted IN PLACE into
        its existing buffer, so the table's pointers never change (fine-tuning touches 3 of the 36 blocks per step;
        rebuilding the table or re-casting all 88 M weights every step costs more than those blocks' forward)."""
        params = params_of(self.convnext)
        vsum = sum(p._version + getattr(p, "_ccx_epoch", 0) for p in params)
        key = (self.compute_dtype, params[0].data_ptr(), params[-1].data_ptr(), vsum)
        if self._prep_key == key:
            return self._prep[0]
        cd = self.compute_dtype
        if (self._prep is not None and self._prep_key is not None and self._prep_key[:3] == key[:3]):
            # same storage, same dtype: refresh only the derived copies whose source parameter changed
            for ent in self._derived:
                p = ent["param"]
                k = p._version + getattr(p, "_ccx_epoch", 0)
                if k != ent["key"]:
                    ent["refresh"](p.detach())
                    ent["key"] = k
            self._prep_key = key
            return self._prep[0]
        keep, derived = [], []

        def f32(p):
            t = p.detach()
            if t.dtype != torch.float32 or not t.is_cuda:
                raise ValueError("Encoder parameters must be float32 CUDA tensors (call .cuda())")
            if not t.is_contiguous():
                raise ValueError("Encoder parameters must be contiguous")
            keep.append(t)
            return t.data_ptr()

        def relayout(p, fn):
            """fp32 re-laid-out copy (depthwise filter tap-major, stem [48][128]) refreshed in place."""
            buf = fn(p.detach()).contiguous()
            derived.append({"param": p, "key": p._version + getattr(p, "_ccx_epoch", 0),
                            "refresh": lambda t, buf=buf, fn=fn: buf.copy_(fn(t))})
            keep.append(buf)
            return buf

        def operand(p, fn=lambda t: t):
            op = Operand.prepare(fn(p.detach()).contiguous(), cd)
            derived.append({"param": p, "key": p._version + getattr(p, "_ccx_epoch", 0),
                            "refresh": lambda t, op=op, fn=fn: op.refresh(fn(t))})
            keep.append(op)
            return op

        def op_ptrs(op):
            return op.hi.data_ptr(), (op.lo.data_ptr() if op.lo is not None else None)

        self._block_ops = []   # per CNBlock: dict(dw_w, w1, w2) python-side handles (used by encoder_train.py)
        self._down_ops = {}    # downsample child index (2, 4, 6) -> Operand of the re-ordered conv weight
        w = _lib.EncoderWeights()
        ch = list(self.convnext.children())
        stem_w = relayout(ch[0][0].weight, lambda t: t.reshape(128, 48).t())
        w.stem_w, w.stem_b = stem_w.data_ptr(), f32(ch[0][0].bias)
        w.stem_ln_g, w.stem_ln_b = f32(ch[0][1].weight), f32(ch[0][1].bias)
        bi = 0
        for s in range(4):
            Cc = DIMS[s]
            for blk in ch[1 + 2 * s]:
                bw = w.blocks[bi]
                dw = relayout(blk.block[0].weight, lambda t, Cc=Cc: t.reshape(Cc, 49).t())
                bw.dw_w, bw.dw_b = dw.data_ptr(), f32(blk.block[0].bias)
                bw.ln_g, bw.ln_b = f32(blk.block[2].weight), f32(blk.block[2].bias)
                w1, w2 = operand(blk.block[3].weight), operand(blk.block[5].weight)
                bw.w1, bw.w1_lo = op_ptrs(w1)
                bw.b1 = f32(blk.block[3].bias)
                bw.w2, bw.w2_lo = op_ptrs(w2)
                bw.b2 = f32(blk.block[5].bias)
                bw.layer_scale = f32(blk.layer_scale)          # (C,1,1) contiguous == [C]
                self._block_ops.append({"dw_w": dw, "w1": w1, "w2": w2})
                bi += 1
            if s > 0:
                d = ch[2 * s]
                dwn = w.down[s - 1]
                dwn.ln_g, dwn.ln_b = f32(d[0].weight), f32(d[0].bias)
                wd = operand(d[1].weight, lambda t, Cc=Cc, Ci=DIMS[s - 1]: t.permute(0, 2, 3, 1).reshape(Cc, 4 * Ci))
                dwn.w, dwn.w_lo = op_ptrs(wd)
                self._down_ops[2 * s] = wd
                dwn.b = f32(d[1].bias)
        for s in range(4):
            w.depths[s], w.dims[s] = DEPTHS[s], DIMS[s]
        w.compute_dtype = _lib.dt_code(cd)
        self._prep, self._prep_key, self._derived = (w, keep), key, derived
        return w

    def _workspace(self, B, H, W, device):
        need = _lib.lib().ccx_encoder_workspace_bytes(B, H, W, _lib.dt_code(self.compute_dtype))
        if self._ws is None or self._ws.numel() < need or self._ws.device != device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=device)
        return self._ws

    def run_children(self, x, child_begin, child_end, noise=None, image_hw=None):
        """Run ``convnext[child_begin:child_end]`` (no autograd).  x: NCHW images if child_begin == 0 else the NHWC
        fp32 stream; returns the NHWC fp32 stream.  image_hw: original image size when child_begin > 0."""
        if child_begin == 0:
            B, _, H, W = x.shape
        else:
            B = x.shape[0]
            H, W = image_hw
        stage = (child_end - 1) // 2  # stage the stream is in after the last child run (0 stem, odd = blocks, even = downsample)
        oh, ow, oc = H >> (2 + stage), W >> (2 + stage), DIMS[stage]
        out = torch.empty((B, oh, ow, oc), dtype=torch.float32, device=x.device)
        ws = self._workspace(B, H, W, x.device)
        w = self.prepared()
        rc = _lib.lib().ccx_encoder_run(C.byref(w), x.data_ptr(), out.data_ptr(), B, H, W, child_begin, child_end,
                                        _lib.ptr(noise), ws.data_ptr(), ws.numel(), _lib.stream_ptr())
        _lib.check(rc, f"encoder_run[{child_begin}:{child_end}]")
        return out
