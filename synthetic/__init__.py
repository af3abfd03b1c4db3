"""synthetic/ — seeded synthetic inputs and random-init weights in the reference's ``state_dict`` layout.

Neutral ground: neither the product (``imagecaptioningconvnext_b200``) nor the oracle.  ``bench.py`` and the tests both
draw their weights / images / captions from here, so the benchmark's own arm never has to import ``oracle``.
Weights use the reference's module-construction order and initialisers (models/decoder.py:35-61,
models/transformerDecoder.py:54-86, torchvision's convnext_base initialiser behind models/encoder.py:18-20), so the
tensors equal what the reference modules hold under the same seed (checked by tests/golden/make_golden.py).
"""
import math

import torch


def positional_encoding(embed_dim, max_len, dtype=torch.float32):
    """The constant ``pos_encoding.pe`` buffer of the reference checkpoint layout (models/transformerDecoder.py:14-27)."""
    pe = torch.zeros(max_len, embed_dim)
    pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, embed_dim, 2).float() * (-math.log(10000.0) / embed_dim))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.to(dtype)


def _perturb(sd, seed, scale):
    """Add small noise to every tensor so that zero-initialised biases / identical cloned layers do not hide
    indexing bugs.  Deterministic in (seed, key order)."""
    g = torch.Generator().manual_seed(seed + 7919)
    for k in sd:
        if sd[k].is_floating_point() and k != "pos_encoding.pe":
            sd[k] = sd[k] + scale * sd[k].abs().mean().clamp_min(0.02) * torch.randn(sd[k].shape, generator=g)
    return sd


def random_lstm_decoder_state(seed=0, vocab=9490, attention_dim=512, embed_dim=512, decoder_dim=512,
                              encoder_dim=1024, end_bias=None, perturb=0.5):
    """Same module-construction order as models/decoder.py:35-61, so under the same seed the tensors equal
    ``DecoderWithAttention(...).state_dict()`` bit for bit (checked by tests/golden/make_golden.py) before the
    optional perturbation.  end_bias: value written to fc.bias[<end> = vocab-1]."""
    from torch import nn
    rng = torch.random.get_rng_state()
    torch.manual_seed(seed)
    mods = {}
    mods["attention.encoder_att"] = nn.Linear(encoder_dim, attention_dim)
    mods["attention.decoder_att"] = nn.Linear(decoder_dim, attention_dim)
    mods["attention.full_att"] = nn.Linear(attention_dim, 1)
    mods["embedding"] = nn.Embedding(vocab, embed_dim)
    mods["decode_step"] = nn.LSTMCell(embed_dim + encoder_dim, decoder_dim, bias=True)
    mods["init_h"] = nn.Linear(encoder_dim, decoder_dim)
    mods["init_c"] = nn.Linear(encoder_dim, decoder_dim)
    mods["f_beta"] = nn.Linear(decoder_dim, encoder_dim)
    mods["fc"] = nn.Linear(decoder_dim, vocab)
    mods["embedding"].weight.data.uniform_(-0.1, 0.1)
    mods["fc"].bias.data.fill_(0)
    mods["fc"].weight.data.uniform_(-0.1, 0.1)
    torch.random.set_rng_state(rng)
    sd = {}
    for name, m in mods.items():
        for k, v in m.state_dict().items():
            sd[f"{name}.{k}"] = v.detach().clone()
    if perturb:
        _perturb(sd, seed, perturb)
    if end_bias is not None:
        sd["fc.bias"][vocab - 1] = end_bias
    return sd


def random_transformer_decoder_state(seed=0, vocab=9490, embed_dim=512, decoder_dim=512, max_len=52,
                                     encoder_dim=1024, nheads=8, nlayers=6, end_bias=None, perturb=0.5):
    """Same construction order as models/transformerDecoder.py:54-86 (random embeddings branch)."""
    from torch import nn
    rng = torch.random.get_rng_state()
    torch.manual_seed(seed)
    emb = nn.Embedding(vocab, embed_dim)
    layer = nn.TransformerDecoderLayer(d_model=embed_dim, nhead=nheads, dim_feedforward=decoder_dim, dropout=0.5)
    dec = nn.TransformerDecoder(layer, num_layers=nlayers)
    fc_out = nn.Linear(embed_dim, vocab)
    proj = nn.Linear(encoder_dim, embed_dim)
    torch.random.set_rng_state(rng)
    sd = {"embedding.weight": emb.weight.detach().clone(),
          "pos_encoding.pe": positional_encoding(embed_dim, max_len).unsqueeze(0)}
    for k, v in dec.state_dict().items():
        sd["transformer_decoder." + k] = v.detach().clone()
    for k, v in fc_out.state_dict().items():
        sd["fc_out." + k] = v.detach().clone()
    for k, v in proj.state_dict().items():
        sd["encoder_proj." + k] = v.detach().clone()
    if perturb:
        _perturb(sd, seed, perturb)
    if end_bias is not None:
        sd["fc_out.bias"][vocab - 1] = end_bias
    return sd


def synthetic_features(B, seed, P=49, E=1024):
    """Encoder-output-like features (B, 7, 7, E): non-negative-ish, O(1) scale."""
    g = torch.Generator().manual_seed(seed)
    s = int(round(P ** 0.5))
    return torch.randn(B, s, s, E, generator=g) * 0.7


def synthetic_captions(B, seed, vocab=9490, T=52, min_len=7):
    """<start>, len-2 tokens in [1, V-4], <end>, then <pad>=0; lengths uniform in [min_len, T]."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(min_len, T + 1, (B, 1), generator=g)
    caps = torch.zeros(B, T, dtype=torch.long)
    for b in range(B):
        L = int(lens[b])
        caps[b, 0] = vocab - 2
        caps[b, 1:L - 1] = torch.randint(1, vocab - 3, (L - 2,), generator=g)
        caps[b, L - 1] = vocab - 1
    return caps, lens


def synthetic_images(batch, seed):
    """randn (B, 3, 256, 256) fp32 ~ the post-Normalize statistics of dataLoader.py:43-45."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(batch, 3, 256, 256, generator=g)


def random_encoder_state(seed=0, layer_scale=1.0, dtype=torch.float32):
    """Random-init ConvNeXt-Base weights in the reference's key layout (``convnext.*``), with layer_scale
    overwritten (the default 1e-6 hides CNBlock bugs).  Uses torchvision's own initialiser so
    the statistics match what ``Encoder()`` would hold before loading ImageNet weights."""
    import torchvision

    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    feats = torchvision.models.convnext_base(weights=None).features
    torch.random.set_rng_state(g)
    sd = {"convnext." + k: v.detach().clone().to(dtype) for k, v in feats.state_dict().items()}
    for k in sd:
        if k.endswith("layer_scale"):
            sd[k].fill_(layer_scale)
    return sd
