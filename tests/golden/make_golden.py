"""Generate the golden fixtures in tests/golden/ by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference; the GPU box never runs this):
    python tests/golden/make_golden.py
The reference ships no tests or known-answer vectors, so the oracle (oracle/*.py) is pinned against
outputs of the reference itself: reference nn.Modules imported from /root/reference with three shims
(no-download convnext_base, gensim stub, matplotlib/skimage stubs), random-init weights under fixed seeds,
synthetic inputs.  Only (seed, small inputs, outputs) are stored — weights are regenerated from the seed.
"""
import os
import sys
import types

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def install_shims():
    import torchvision

    real = torchvision.models.convnext_base

    def convnext_base_no_download(*a, **k):  # models/encoder.py:18 asks for IMAGENET1K_V1 (network)
        k["weights"] = None
        return real(*a[:0], **k)

    torchvision.models.convnext_base = convnext_base_no_download
    for name in ("gensim", "gensim.downloader", "gensim.models", "matplotlib", "matplotlib.pyplot",
                 "matplotlib.cm", "skimage", "skimage.transform"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["gensim.models"].KeyedVectors = type("KeyedVectors", (), {})
    sys.modules["gensim"].downloader = sys.modules["gensim.downloader"]
    sys.modules["gensim"].models = sys.modules["gensim.models"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].cm = sys.modules["matplotlib.cm"]
    sys.modules["skimage"].transform = sys.modules["skimage.transform"]
    if REF not in sys.path:
        sys.path.insert(0, REF)


def golden_encoder():
    """Reference Encoder (models/encoder.py) on 1 synthetic 256x256 image and 2 synthetic 64x64 images."""
    from models.encoder import Encoder
    from oracle.encoder_oracle import random_encoder_state

    out = {}
    for name, seed, shape, s in (("img256_s7", 0, (1, 3, 256, 256), 7), ("img64_s7", 1, (2, 3, 64, 64), 7),
                                 ("img256_s14", 0, (1, 3, 256, 256), 14)):
        torch.manual_seed(seed)
        enc = Encoder(encoded_image_size=s).eval()
        sd = random_encoder_state(seed=seed, layer_scale=1.0)
        enc.load_state_dict(sd)
        g = torch.Generator().manual_seed(1234 + seed)
        x = torch.randn(*shape, generator=g)
        with torch.no_grad():
            y = enc(x).contiguous()
        out[name] = {"weight_seed": seed, "input_seed": 1234 + seed, "shape": shape, "enc_size": s, "out": y}
        print(name, tuple(y.shape), float(y.abs().max()), float(y.std()))
    torch.save(out, os.path.join(HERE, "encoder.pt"))


V = 9490
WORDMAP = {"<pad>": 0, "<unk>": V - 3, "<start>": V - 2, "<end>": V - 1}
WORDMAP.update({f"w{i}": i for i in range(1, V - 3)})   # len(wordMap) == V, as caption.py:49,163 needs
assert len(WORDMAP) == V


def _sub(preds):
    """Keep the fixtures small: every 31st vocabulary column + per-row argmax and logsumexp."""
    return {"sub": preds[..., ::31].contiguous(), "argmax": preds.argmax(-1), "lse": preds.logsumexp(-1)}


def golden_lstm():
    """Reference DecoderWithAttention (models/decoder.py): teacher forcing (eval) and greedy."""
    from models.decoder import DecoderWithAttention
    from oracle import decoder_oracle as do

    out = {}
    seed = 0
    torch.manual_seed(seed)
    ref = DecoderWithAttention(512, 512, 512, V, torch.device("cpu")).eval()
    plain = do.random_lstm_decoder_state(seed, V, perturb=0)
    for k, v in ref.state_dict().items():          # the oracle's generator reproduces the reference init bit for bit
        assert torch.equal(v, plain[k]), k
    sd = do.random_lstm_decoder_state(seed, V, end_bias=0.21)
    ref.load_state_dict(sd)
    B = 5
    enc = do.synthetic_features(B, 100)
    caps, lens = do.synthetic_captions(B, 101, V)
    with torch.no_grad():
        preds, caps_s, dl, alphas, sort_ind = ref(teacherForcing=True, encoder_out=enc, encoded_captions=caps,
                                                  caption_lengths=lens)
        loss = do.train_loss_lstm(preds, caps_s, dl, alphas)
        gp, ga, gs = ref(teacherForcing=False, encoder_out=enc, wordMap=WORDMAP, maxDecodeLen=51)
    out["tf"] = {"B": B, "feat_seed": 100, "cap_seed": 101, "weight_seed": seed, "end_bias": 0.21,
                 "preds": _sub(preds), "alphas": alphas, "sort_ind": sort_ind, "decode_lengths": dl,
                 "caps_sorted": caps_s, "loss": loss}
    out["greedy"] = {"preds": _sub(gp), "alphas": ga, "sequences": gs}
    print("lstm tf loss", float(loss), "greedy lengths", [(int((r == V - 1).nonzero()[0]) if (r == V - 1).any() else -1) for r in gs])
    torch.save(out, os.path.join(HERE, "lstm_decoder.pt"))


def golden_transformer():
    """Reference TransformerDecoder (models/transformerDecoder.py): teacher forcing (eval) and greedy."""
    from models.transformerDecoder import TransformerDecoder
    from oracle import decoder_oracle as do

    out = {}
    seed = 0
    torch.manual_seed(seed)
    ref = TransformerDecoder(512, 512, V, 52, torch.device("cpu"), None, None, True).eval()
    plain = do.random_transformer_decoder_state(seed, V, perturb=0)
    for k, v in ref.state_dict().items():
        assert torch.equal(v, plain[k]), k
    sd = do.random_transformer_decoder_state(seed, V, end_bias=3.2)
    ref.load_state_dict(sd)
    B = 4
    enc = do.synthetic_features(B, 200)
    caps, lens = do.synthetic_captions(B, 201, V)
    kpm = caps == 0
    with torch.no_grad():
        preds, caps_o, dl = ref(teacherForcing=True, encoder_out=enc, encoded_captions=caps, caption_lengths=lens,
                                tgt_key_padding_mask=kpm)
        loss = do.train_loss_transformer(preds, caps_o, dl)
        gp, gs = ref(teacherForcing=False, encoder_out=enc, wordMap=WORDMAP, maxDecodeLen=51)
    out["tf"] = {"B": B, "feat_seed": 200, "cap_seed": 201, "weight_seed": seed, "end_bias": 3.2,
                 "preds": _sub(preds), "decode_lengths": dl, "loss": loss}
    out["greedy"] = {"preds": _sub(gp), "sequences": gs}
    print("transformer tf loss", float(loss), "greedy lengths", [(int((r == V - 1).nonzero()[0]) if (r == V - 1).any() else -1) for r in gs])
    torch.save(out, os.path.join(HERE, "transformer_decoder.pt"))


def _grad_digest(named_grads, stride=499):
    """Small fingerprint of a tensor set: per tensor its norm and every stride-th element."""
    return {k: {"norm": g.norm(), "sub": g.reshape(-1)[::stride].clone()} for k, g in named_grads}


def golden_free_running():
    """Reference free-running TRAINING step body (trainMultiGPU.py:444-460) on both decoders: greedy forward with
    autograd, preprocessDecoderOutputForMetrics, CrossEntropyLoss (+ alpha term), backward.  Modules in eval mode
    (dropout is random in train mode; the arithmetic is otherwise identical)."""
    import torch.nn as nn
    from models.decoder import DecoderWithAttention
    from models.transformerDecoder import TransformerDecoder
    from oracle import decoder_oracle as do
    from oracle.metrics_oracle import preprocess_decoder_output_for_metrics as restated

    # utils/utils.py cannot be imported (h5py); its preprocessDecoderOutputForMetrics is exec'd from the source text
    src = open(os.path.join(REF, "utils", "utils.py")).read()
    a = src.index("def preprocessDecoderOutputForMetrics")
    ns = {"torch": torch}
    exec(src[a:], ns)
    preprocess = ns["preprocessDecoderOutputForMetrics"]
    crit = nn.CrossEntropyLoss()
    out = {}
    B = 4
    for kind, end_bias, fs in (("lstm", 0.21, 300), ("transformer", 3.2, 310)):
        if kind == "lstm":
            ref = DecoderWithAttention(512, 512, 512, V, torch.device("cpu")).eval()
            sd = do.random_lstm_decoder_state(0, V, end_bias=end_bias)
        else:
            ref = TransformerDecoder(512, 512, V, 52, torch.device("cpu"), None, None, True).eval()
            sd = do.random_transformer_decoder_state(0, V, end_bias=end_bias)
        ref.load_state_dict(sd)
        enc = do.synthetic_features(B, fs).requires_grad_(True)
        caps, lens = do.synthetic_captions(B, fs + 1, V)
        res = ref(teacherForcing=False, encoder_out=enc, wordMap=WORDMAP, maxDecodeLen=51)
        scores, seqs = res[0], res[-1]
        su, tu, ntok, adl = preprocess(scores, seqs, caps, WORDMAP["<end>"], WORDMAP["<pad>"], 51)
        su2, tu2, ntok2, adl2 = restated(scores, seqs, caps, WORDMAP["<end>"], WORDMAP["<pad>"], 51)
        assert torch.equal(su, su2) and torch.equal(tu, tu2) and ntok == ntok2 and adl == adl2
        loss = crit(su, tu)
        if kind == "lstm":
            loss = loss + 1.0 * ((1.0 - res[1].sum(dim=1)) ** 2).mean()
        loss.backward()
        grads = [(k, p.grad) for k, p in ref.named_parameters() if p.grad is not None]
        out[kind] = {"B": B, "feat_seed": fs, "cap_seed": fs + 1, "weight_seed": 0, "end_bias": end_bias,
                     "loss": loss.detach(), "sequences": seqs, "decode_lengths": adl, "tokens": ntok,
                     "enc_grad_norm": enc.grad.norm(), "enc_grad_sub": enc.grad[..., ::16].clone(), "grads": _grad_digest(grads)}
        print(kind, "free-running loss", float(loss), "lengths", adl, "tokens", ntok)
    torch.save(out, os.path.join(HERE, "free_running.pt"))


def golden_attvis():
    """Reference TransformerDecoderForAttentionViz (models/transformerDecoderAttVis.py): teacher forcing and greedy,
    eval mode, same weights as the TransformerDecoder golden under the reference class's key names."""
    from models.transformerDecoderAttVis import TransformerDecoderForAttentionViz
    from oracle import decoder_oracle as do

    sd = do.random_transformer_decoder_state(0, V, end_bias=3.2)
    ref = TransformerDecoderForAttentionViz(512, 512, V, 52, torch.device("cpu")).eval()
    ref.load_state_dict({k.replace("transformer_decoder.layers.", "decoder_layers."): v for k, v in sd.items()})
    B = 4
    enc = do.synthetic_features(B, 200)
    caps, lens = do.synthetic_captions(B, 201, V)
    with torch.no_grad():
        preds, _, dl, alphas = ref(teacherForcing=True, encoder_out=enc, encoded_captions=caps, caption_lengths=lens,
                                   tgt_key_padding_mask=caps == 0)
        gp, gs, ga = ref(teacherForcing=False, encoder_out=enc, wordMap=WORDMAP, maxDecodeLen=51)
    out = {"B": B, "feat_seed": 200, "cap_seed": 201, "weight_seed": 0, "end_bias": 3.2,
           "state_dict_keys": sorted(ref.state_dict().keys()),
           "tf": {"preds": _sub(preds), "alphas": alphas, "decode_lengths": dl},
           "greedy": {"preds": _sub(gp), "sequences": gs, "alphas": ga}}
    print("attvis alphas", tuple(alphas.shape), float(alphas.sum(-1).mean()), tuple(ga.shape))
    torch.save(out, os.path.join(HERE, "attvis.pt"))


def golden_train_step():
    """The reference's train-step body, statement by statement (trainMultiGPU.py:361-394 = train.py:261-291) on the
    reference modules: Encoder.fine_tune(True, 7) + each decoder, pack_padded_sequence + CrossEntropyLoss (+ alpha
    regulariser), zero_grad, backward, utils.clip_gradient (exec'd from the source text: utils/utils.py imports
    h5py), torch.optim.Adam.  Two steps in eval mode (dropout / stochastic depth are random in train mode); stored:
    the losses and a digest of every trainable tensor after the second step."""
    import torch.nn as nn
    from torch.nn.utils.rnn import pack_padded_sequence
    from models.decoder import DecoderWithAttention
    from models.encoder import Encoder
    from models.transformerDecoder import TransformerDecoder
    from oracle import decoder_oracle as do
    from oracle.encoder_oracle import random_encoder_state

    src = open(os.path.join(REF, "utils", "utils.py")).read()
    a = src.index("def clip_gradient")
    b = src.index("def save_checkpoint")
    ns = {"torch": torch}
    exec(src[a:b], ns)
    clip_gradient = ns["clip_gradient"]
    criterion = nn.CrossEntropyLoss()
    out = {"B": 3, "image_seed": 1, "cap_seed": 2, "encoder_seed": 0, "decoder_seed": 7, "lr": 1e-3, "grad_clip": 5.0,
           "image_hw": 64}
    B = out["B"]
    imgs = torch.randn(B, 3, 64, 64, generator=torch.Generator().manual_seed(out["image_seed"]))
    caps, lens = do.synthetic_captions(B, out["cap_seed"], V)
    for kind in ("lstm", "transformer"):
        enc = Encoder().eval()
        enc.load_state_dict(random_encoder_state(seed=0, layer_scale=1.0))
        enc.fine_tune(True, 7)
        if kind == "lstm":
            dec = DecoderWithAttention(512, 512, 512, V, torch.device("cpu")).eval()
            dec.load_state_dict(do.random_lstm_decoder_state(7, V))
        else:
            dec = TransformerDecoder(512, 512, V, 52, torch.device("cpu"), None, None, True).eval()
            dec.load_state_dict(do.random_transformer_decoder_state(7, V))
        dec_opt = torch.optim.Adam(params=filter(lambda p: p.requires_grad, dec.parameters()), lr=out["lr"])
        enc_opt = torch.optim.Adam(params=filter(lambda p: p.requires_grad, enc.parameters()), lr=out["lr"])
        losses = []
        for _ in range(2):
            feats = enc(imgs)
            if kind == "lstm":
                scores, caps_sorted, dl, alphas, _ = dec(teacherForcing=True, encoder_out=feats, encoded_captions=caps,
                                                         caption_lengths=lens)
                targets = caps_sorted[:, 1:]
                scores = pack_padded_sequence(scores, dl, batch_first=True).data
                targets = pack_padded_sequence(targets, dl, batch_first=True).data
                loss = criterion(scores, targets)
                loss += 1.0 * ((1. - alphas.sum(dim=1)) ** 2).mean()
            else:
                kpm = caps == WORDMAP["<pad>"]
                scores, caps_sorted, dl = dec(teacherForcing=True, encoder_out=feats, encoded_captions=caps,
                                              caption_lengths=lens, tgt_key_padding_mask=kpm)
                targets = caps_sorted[:, 1:]
                scores = pack_padded_sequence(scores, dl, batch_first=True, enforce_sorted=False).data
                targets = pack_padded_sequence(targets, dl, batch_first=True, enforce_sorted=False).data
                loss = criterion(scores, targets)
            enc_opt.zero_grad()
            dec_opt.zero_grad()
            loss.backward()
            clip_gradient(dec_opt, out["grad_clip"])
            clip_gradient(enc_opt, out["grad_clip"])
            enc_opt.step()
            dec_opt.step()
            losses.append(float(loss))
        tr = [("decoder." + k, p.detach()) for k, p in dec.named_parameters() if p.requires_grad]
        tr += [("encoder." + k, p.detach()) for k, p in enc.named_parameters() if p.requires_grad]
        out[kind] = {"losses": losses, "weights": _grad_digest(tr, stride=1999)}
        print(kind, "train-step losses", losses, "trainable tensors", len(tr))
    torch.save(out, os.path.join(HERE, "train_step.pt"))


def golden_fine_tune():
    """Which parameters the reference's Encoder.fine_tune (models/encoder.py:29-34) leaves trainable, for every
    startingLayer, and the constructor default."""
    from models.encoder import Encoder
    enc = Encoder()
    out = {"default": sorted(n for n, p in enc.named_parameters() if p.requires_grad),
           "n_params": sum(p.numel() for p in enc.parameters()), "keys": sorted(enc.state_dict().keys())}
    for L in range(0, 9):
        enc.fine_tune(True, L)
        out[L] = sorted(n for n, p in enc.named_parameters() if p.requires_grad)
    enc.fine_tune(False)
    out["off"] = sorted(n for n, p in enc.named_parameters() if p.requires_grad)
    print("fine_tune: default trainable", len(out["default"]), "params", out["n_params"],
          {L: len(out[L]) for L in range(9)})
    torch.save(out, os.path.join(HERE, "fine_tune.pt"))


def golden_beam():
    """Reference caption.py beam search (k=5), full pipeline image file -> Encoder -> decoder, both decoders."""
    import numpy as np
    from PIL import Image
    import caption as cap                      # /root/reference/caption.py (matplotlib / skimage stubbed)
    from models.decoder import DecoderWithAttention
    from models.encoder import Encoder
    from models.transformerDecoder import TransformerDecoder
    from oracle import decoder_oracle as do
    from oracle.encoder_oracle import random_encoder_state

    cap.device = torch.device("cpu")
    enc = Encoder().eval()
    enc.load_state_dict(random_encoder_state(seed=0, layer_scale=1.0))
    lstm = DecoderWithAttention(512, 512, 512, V, torch.device("cpu")).eval()
    lstm.load_state_dict(do.random_lstm_decoder_state(0, V, end_bias=0.5))
    tr = TransformerDecoder(512, 512, V, 52, torch.device("cpu"), None, None, True).eval()
    tr.load_state_dict(do.random_transformer_decoder_state(0, V, end_bias=3.6))
    out = {"image_seeds": [11, 12], "k": 5, "lstm_end_bias": 0.5, "transformer_end_bias": 3.6, "lstm": [], "transformer": [], "features": [],
           "lstm_alphas": []}
    for seed in out["image_seeds"]:
        img = np.random.RandomState(seed).randint(0, 256, size=(256, 256, 3), dtype=np.uint8)
        path = f"/tmp/golden_beam_{seed}.png"
        Image.fromarray(img).save(path)
        with torch.no_grad():
            seq, alphas = cap.caption_image_beam_search(enc, lstm, path, WORDMAP, 5)
            seq_t, _ = cap.caption_image_beam_search_transformer(enc, tr, path, WORDMAP, 5)
            x = torch.from_numpy(img.transpose(2, 0, 1) / 255.).float()
            mean = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
            std = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)
            feats = enc(((x - mean) / std).unsqueeze(0)).contiguous()
        out["lstm"].append(seq)
        out["lstm_alphas"].append(torch.tensor(alphas))        # (len(seq), 7, 7), caption.py:153
        out["transformer"].append(seq_t)
        out["features"].append(feats[0, ::3, ::3, ::64].clone())   # spot-check values of the encoder output
        print("beam", seed, seq, seq_t)
    # longer LSTM captions so the attention-map chain has several steps: random-init decoders either emit <end> at
    # once or never, so <end> is re-mapped (wordMap is an input of caption.py) to word 351, which these weights emit
    # at step 3 — captions of 4 tokens with beam re-ordering at step 2
    end_word = 351
    wm = dict(WORDMAP)
    wm["<end>"], wm[f"w{end_word}"] = end_word, V - 1
    out["lstm_long"] = {"end_bias": 0.0, "end_word": end_word, "image_seeds": [11, 12], "seqs": [], "alphas": []}
    lstm.load_state_dict(do.random_lstm_decoder_state(0, V, end_bias=0.0))
    for seed in out["lstm_long"]["image_seeds"]:
        path = f"/tmp/golden_beam_{seed}.png"
        with torch.no_grad():
            seq, alphas = cap.caption_image_beam_search(enc, lstm, path, wm, 5)
        out["lstm_long"]["seqs"].append(seq)
        out["lstm_long"]["alphas"].append(torch.tensor(alphas))
        print("beam long", seed, seq)
    torch.save(out, os.path.join(HERE, "beam.pt"))


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    torch.set_flush_denormal(True)
    install_shims()
    which = sys.argv[1:] or ["encoder"]
    for w in which:
        globals()["golden_" + w]()
