"""CPU oracle for ``Encoder.forward`` (reference models/encoder.py:23-27).  Test infrastructure only.

The reference calls ``torchvision.models.convnext_base(...).features`` (models/encoder.py:18-19), so the
arithmetic restated here is torchvision 0.26.0's:
  * stem  Conv2d(3,128,k4,s4)+LayerNorm2d           tv:models/convnext.py:120-131, 31-36
  * CNBlock: dw7x7 -> LN(eps 1e-6) -> Linear(C,4C) -> GELU(erf) -> Linear(4C,C) -> layer_scale
             -> StochasticDepth(row) -> + input       tv:models/convnext.py:51-67, tv:ops/stochastic_depth.py:8-44
  * downsample LayerNorm2d + Conv2d(k2,s2)            tv:models/convnext.py:146-151
  * AdaptiveAvgPool2d((s,s)) + permute(0,2,3,1)       models/encoder.py:25-26

Works on a reference-format state dict (keys ``convnext.<child>...``), any float dtype.
"""
import torch
import torch.nn.functional as F

# (child index in convnext.features, channels, number of CNBlocks); children 0/2/4/6 are stem / downsamples
STAGES = ((1, 128, 3), (3, 256, 3), (5, 512, 27), (7, 1024, 3))
# torchvision convnext_base(stochastic_depth_prob=0.5): p ramps linearly over the 36 blocks
_TOTAL_BLOCKS = 36
_SD_PROB = 0.5


def stochastic_depth_probs():
    """tv:models/convnext.py:137-141 — sd_prob = 0.5 * block_id / (total_blocks - 1)."""
    out, bid = {}, 0
    for child, _, nblk in STAGES:
        for i in range(nblk):
            out[(child, i)] = _SD_PROB * bid / (_TOTAL_BLOCKS - 1.0)
            bid += 1
    return out


def _ln2d(x, w, b, eps=1e-6):
    # tv:models/convnext.py:31-36 LayerNorm2d: permute -> layer_norm over C -> permute back
    x = x.permute(0, 2, 3, 1)
    x = F.layer_norm(x, (x.shape[-1],), w, b, eps)
    return x.permute(0, 3, 1, 2)


def cnblock(x, sd, prefix, noise=None):
    """tv:models/convnext.py:63-67.  ``noise`` (B,) is the stochastic-depth row factor
    bernoulli(1-p)/(1-p) (tv:ops/stochastic_depth.py:33-39); None = eval mode."""
    C = x.shape[1]
    y = F.conv2d(x, sd[prefix + "block.0.weight"], sd[prefix + "block.0.bias"], padding=3, groups=C)
    y = y.permute(0, 2, 3, 1)
    y = F.layer_norm(y, (C,), sd[prefix + "block.2.weight"], sd[prefix + "block.2.bias"], 1e-6)
    y = F.linear(y, sd[prefix + "block.3.weight"], sd[prefix + "block.3.bias"])
    y = F.gelu(y)  # exact erf GELU (nn.GELU() default)
    y = F.linear(y, sd[prefix + "block.5.weight"], sd[prefix + "block.5.bias"])
    y = y.permute(0, 3, 1, 2)
    y = sd[prefix + "layer_scale"] * y
    if noise is not None:
        y = y * noise.view(-1, 1, 1, 1)
    return y + x


def encoder_forward(sd, images, encoded_image_size=7, noise=None, return_pre_pool=False):
    """models/encoder.py:23-27.  images (B,3,H,W) NCHW -> (B,s,s,1024).
    noise: optional dict {(child, block): (B,) tensor} of stochastic-depth row factors (train mode)."""
    p = "convnext."
    x = F.conv2d(images, sd[p + "0.0.weight"], sd[p + "0.0.bias"], stride=4)
    x = _ln2d(x, sd[p + "0.1.weight"], sd[p + "0.1.bias"])
    for child, _, nblk in STAGES:
        if child > 1:
            d = child - 1
            x = _ln2d(x, sd[f"{p}{d}.0.weight"], sd[f"{p}{d}.0.bias"])
            x = F.conv2d(x, sd[f"{p}{d}.1.weight"], sd[f"{p}{d}.1.bias"], stride=2)
        for i in range(nblk):
            nz = None if noise is None else noise.get((child, i))
            x = cnblock(x, sd, f"{p}{child}.{i}.", nz)
    if return_pre_pool:
        return x
    x = F.adaptive_avg_pool2d(x, (encoded_image_size, encoded_image_size))
    return x.permute(0, 2, 3, 1)


from synthetic import random_encoder_state  # noqa: F401,E402  (neutral module; re-exported for the tests)
