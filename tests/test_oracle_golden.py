"""CPU-only: the oracle restatement reproduces the golden outputs generated from the reference itself."""
import os

import torch

from conftest import rel_err


def test_encoder_oracle_matches_reference_golden(golden_dir):
    from oracle.encoder_oracle import encoder_forward, random_encoder_state
    gold = torch.load(os.path.join(golden_dir, "encoder.pt"))
    for name in ("img64_s7", "img256_s7", "img256_s14"):
        g = gold[name]
        sd = random_encoder_state(seed=g["weight_seed"], layer_scale=1.0)
        x = torch.randn(*g["shape"], generator=torch.Generator().manual_seed(g["input_seed"]))
        y = encoder_forward(sd, x, g["enc_size"])
        assert y.shape == g["out"].shape
        assert rel_err(y, g["out"]) < 1e-5, name


def test_adaptive_pool_8_to_7_is_avgpool_2_1():
    # SURVEY.md §8c invariant (i)
    import torch.nn.functional as F
    x = torch.randn(2, 16, 8, 8)
    assert torch.equal(F.adaptive_avg_pool2d(x, (7, 7)), F.avg_pool2d(x, 2, 1))


V = 9490
WORDMAP = {"<pad>": 0, "<unk>": V - 3, "<start>": V - 2, "<end>": V - 1}


def _check_sub(preds, g, tol):
    assert rel_err(preds[..., ::31], g["sub"]) < tol
    assert rel_err(preds.logsumexp(-1), g["lse"]) < tol


def test_lstm_oracle_matches_reference_golden(golden_dir):
    from oracle import decoder_oracle as do
    gold = torch.load(os.path.join(golden_dir, "lstm_decoder.pt"))
    g = gold["tf"]
    sd = do.random_lstm_decoder_state(g["weight_seed"], V, end_bias=g["end_bias"])
    enc = do.synthetic_features(g["B"], g["feat_seed"])
    caps, lens = do.synthetic_captions(g["B"], g["cap_seed"], V)
    preds, caps_s, dl, alphas, sort_ind = do.lstm_teacher_forcing(sd, enc, caps, lens)
    assert dl == g["decode_lengths"] and torch.equal(sort_ind, g["sort_ind"]) and torch.equal(caps_s, g["caps_sorted"])
    _check_sub(preds, g["preds"], 1e-5)
    assert rel_err(alphas, g["alphas"]) < 1e-5
    assert abs(float(do.train_loss_lstm(preds, caps_s, dl, alphas)) - float(g["loss"])) < 1e-5
    # invariants (SURVEY.md §8c ii, iii): alpha rows sum to 1 inside, everything exactly 0 past each length
    for b, l in enumerate(dl):
        assert torch.allclose(alphas[b, :l].sum(-1), torch.ones(l), atol=1e-5)
        assert float(preds[b, l:].abs().max() if l < preds.shape[1] else 0) == 0.0
    gp, ga, gs = do.lstm_greedy(sd, enc, V - 2, V - 1, 51)
    assert torch.equal(gs, gold["greedy"]["sequences"])
    _check_sub(gp, gold["greedy"]["preds"], 1e-5)
    assert rel_err(ga, gold["greedy"]["alphas"]) < 1e-5


def test_transformer_oracle_matches_reference_golden(golden_dir):
    from oracle import decoder_oracle as do
    gold = torch.load(os.path.join(golden_dir, "transformer_decoder.pt"))
    g = gold["tf"]
    sd = do.random_transformer_decoder_state(g["weight_seed"], V, end_bias=g["end_bias"])
    enc = do.synthetic_features(g["B"], g["feat_seed"])
    caps, lens = do.synthetic_captions(g["B"], g["cap_seed"], V)
    preds, _, dl = do.transformer_teacher_forcing(sd, enc, caps, lens, caps == 0)
    assert dl == g["decode_lengths"]
    _check_sub(preds, g["preds"], 2e-5)
    assert abs(float(do.train_loss_transformer(preds, caps, dl)) - float(g["loss"])) < 1e-5
    gp, gs = do.transformer_greedy(sd, enc, V - 2, V - 1, 0, 51)
    assert torch.equal(gs, gold["greedy"]["sequences"])
    _check_sub(gp, gold["greedy"]["preds"], 2e-5)


def test_beam_k1_equals_greedy():
    # SURVEY.md §8c invariant (iv) / H4: caption.py's "greedy" is beam search with k=1
    from oracle import decoder_oracle as do
    sd = do.random_lstm_decoder_state(1, V, end_bias=0.21)
    enc = do.synthetic_features(1, 5)
    _, _, seqs = do.lstm_greedy(sd, enc, V - 2, V - 1, 51)
    best, done, _ = do.beam_search(sd, enc, "lstm", 1, V - 2, V - 1, V)
    row = seqs[0].tolist()
    if V - 1 in row:
        n = row.index(V - 1) + 1
        assert best == [V - 2] + row[:n]
    else:
        assert best is None and done == []


def _beam_inputs(gold):
    import numpy as np
    from oracle.encoder_oracle import encoder_forward, random_encoder_state
    esd = random_encoder_state(seed=0, layer_scale=1.0)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(3, 1, 1)
    feats, imgs = [], []
    for seed in gold["image_seeds"]:
        img = np.random.RandomState(seed).randint(0, 256, size=(256, 256, 3), dtype=np.uint8)
        x = ((torch.from_numpy(img.transpose(2, 0, 1) / 255.).float() - mean) / std).unsqueeze(0)
        imgs.append(x)
        feats.append(encoder_forward(esd, x, 7))
    return imgs, feats


def test_beam_oracle_matches_reference_caption_py_golden(golden_dir):
    """Full pipeline image -> Encoder -> beam search (k=5) against caption.py's own output for both decoders."""
    from oracle import decoder_oracle as do
    gold = torch.load(os.path.join(golden_dir, "beam.pt"))
    _, feats = _beam_inputs(gold)
    lsd = do.random_lstm_decoder_state(0, V, end_bias=gold["lstm_end_bias"])
    tsd = do.random_transformer_decoder_state(0, V, end_bias=gold["transformer_end_bias"])
    for i, f in enumerate(feats):
        assert rel_err(f[0, ::3, ::3, ::64], gold["features"][i]) < 1e-5
        assert do.beam_search(lsd, f, "lstm", gold["k"], V - 2, V - 1, V)[0] == gold["lstm"][i]
        assert do.beam_search(tsd, f, "transformer", gold["k"], V - 2, V - 1, V)[0] == gold["transformer"][i]
    # attention maps of the winning caption (caption.py:85,122,129,153) on 4-token captions: <end> re-mapped to a
    # word these random weights emit at step 3 (see make_golden.py)
    long = gold["lstm_long"]
    lsd = do.random_lstm_decoder_state(0, V, end_bias=long["end_bias"])
    for i, f in enumerate(feats):
        al = []
        best = do.beam_search(lsd, f, "lstm", gold["k"], V - 2, long["end_word"], V, alphas_out=al)[0]
        assert best == long["seqs"][i] and len(best) == 4
        ref = long["alphas"][i]
        assert al[0].shape == (4, 49) and torch.equal(al[0][0], torch.ones(49))
        assert rel_err(al[0].view(4, 7, 7), ref) < 1e-4


def _oracle_free_running(kind, g, V=9490):
    """Oracle greedy forward WITH autograd + the restated trainWithoutTeacherForcing loss -> (loss, seqs, grads)."""
    from oracle import decoder_oracle as do
    start, end, pad = V - 2, V - 1, 0
    if kind == "lstm":
        sd = do.random_lstm_decoder_state(g["weight_seed"], V, end_bias=g["end_bias"])
    else:
        sd = do.random_transformer_decoder_state(g["weight_seed"], V, end_bias=g["end_bias"])
    leaf = {k: v.clone().requires_grad_(v.is_floating_point() and k != "pos_encoding.pe") for k, v in sd.items()}
    enc = do.synthetic_features(g["B"], g["feat_seed"]).requires_grad_(True)
    caps, _ = do.synthetic_captions(g["B"], g["cap_seed"], V)
    if kind == "lstm":
        preds, alphas, seqs = do.lstm_greedy(leaf, enc, start, end, 51)
    else:
        preds, seqs = do.transformer_greedy(leaf, enc, start, end, pad, 51)
        alphas = None
    loss = do.free_running_loss(preds, seqs, caps, end, pad, 51, alphas=alphas)
    loss.backward()
    return loss.detach(), seqs, enc.grad, {k: v.grad for k, v in leaf.items() if v.requires_grad and v.grad is not None}


def test_free_running_training_oracle_matches_reference_golden(golden_dir):
    """The oracle's differentiable greedy loops + loss reproduce the reference's free-running train-step body
    (trainMultiGPU.py:444-460): loss, generated sequences and every parameter gradient (digest: norm + strided
    sample) of both decoders."""
    gold = torch.load(os.path.join(golden_dir, "free_running.pt"))
    for kind in ("lstm", "transformer"):
        g = gold[kind]
        loss, seqs, enc_grad, grads = _oracle_free_running(kind, g)
        assert torch.equal(seqs, g["sequences"]), kind
        assert abs(float(loss) - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
        assert rel_err(enc_grad[..., ::16], g["enc_grad_sub"]) < 1e-4
        assert abs(float(enc_grad.norm()) - float(g["enc_grad_norm"])) < 1e-4 * float(g["enc_grad_norm"])
        assert set(grads) == set(g["grads"]), (kind, set(grads) ^ set(g["grads"]))
        for k, d in g["grads"].items():
            ref_n = float(d["norm"])
            assert abs(float(grads[k].norm()) - ref_n) <= 1e-4 * ref_n + 1e-9, (kind, k)
            if ref_n > 0:
                assert rel_err(grads[k].reshape(-1)[::499], d["sub"]) < 1e-3, (kind, k)


def test_attention_viz_oracle_matches_reference_golden(golden_dir):
    """transformerDecoderAttVis.py outputs: same logits as the plain TransformerDecoder plus the attention maps —
    teacher forcing (H, B, P) = mean over layers and target positions (what the reference's reduction really does),
    greedy (B, 51, P) = mean over layers and heads of the newest token, zero once a row has finished."""
    from oracle import decoder_oracle as do
    V = 9490
    g = torch.load(os.path.join(golden_dir, "attvis.pt"))
    sd = do.random_transformer_decoder_state(g["weight_seed"], V, end_bias=g["end_bias"])
    enc = do.synthetic_features(g["B"], g["feat_seed"])
    caps, lens = do.synthetic_captions(g["B"], g["cap_seed"], V)
    with torch.no_grad():
        preds, _, dl, alphas = do.transformer_teacher_forcing(sd, enc, caps, lens, caps == 0, return_alphas=True)
        gp, gs, ga = do.transformer_greedy(sd, enc, V - 2, V - 1, 0, 51, return_alphas=True)
    assert dl == g["tf"]["decode_lengths"]
    _check_sub(preds, g["tf"]["preds"], 1e-4)
    assert alphas.shape == g["tf"]["alphas"].shape == (8, g["B"], 49)
    assert rel_err(alphas, g["tf"]["alphas"]) < 1e-4
    assert torch.equal(gs, g["greedy"]["sequences"])
    assert rel_err(ga, g["greedy"]["alphas"]) < 1e-4
    assert torch.equal(ga == 0, g["greedy"]["alphas"] == 0)


def _oracle_train_steps(kind, g, n_steps=2):
    """The oracle's restatement of the train-step body (the CPU baseline bench.py times): functional encoder /
    decoder on leaf tensors, packed CE (+ alpha term), clamp, torch.optim.Adam."""
    from oracle import decoder_oracle as do
    from oracle import encoder_oracle as eo
    V = 9490
    esd = eo.random_encoder_state(seed=g["encoder_seed"], layer_scale=1.0)
    dsd = (do.random_lstm_decoder_state(g["decoder_seed"], V) if kind == "lstm"
           else do.random_transformer_decoder_state(g["decoder_seed"], V))
    e_leaf = {k: v.clone().requires_grad_(k.startswith("convnext.7.")) for k, v in esd.items()}
    d_leaf = {k: v.clone().requires_grad_(v.is_floating_point() and k != "pos_encoding.pe") for k, v in dsd.items()}
    tr_e = [v for v in e_leaf.values() if v.requires_grad]
    tr_d = [v for v in d_leaf.values() if v.requires_grad]
    opt_e, opt_d = torch.optim.Adam(tr_e, lr=g["lr"]), torch.optim.Adam(tr_d, lr=g["lr"])
    imgs = torch.randn(g["B"], 3, g["image_hw"], g["image_hw"], generator=torch.Generator().manual_seed(g["image_seed"]))
    caps, lens = do.synthetic_captions(g["B"], g["cap_seed"], V)
    losses = []
    for _ in range(n_steps):
        feats = eo.encoder_forward(e_leaf, imgs, 7)
        if kind == "lstm":
            p, cs, dl, al, _ = do.lstm_teacher_forcing(d_leaf, feats, caps, lens)
            loss = do.train_loss_lstm(p, cs, dl, al)
        else:
            p, _, dl = do.transformer_teacher_forcing(d_leaf, feats, caps, lens, caps == 0)
            loss = do.train_loss_transformer(p, caps, dl)
        opt_e.zero_grad()
        opt_d.zero_grad()
        loss.backward()
        for prm in tr_e + tr_d:
            prm.grad.clamp_(-g["grad_clip"], g["grad_clip"])
        opt_e.step()
        opt_d.step()
        losses.append(float(loss))
    weights = {"decoder." + k: v.detach() for k, v in d_leaf.items() if v.requires_grad}
    weights.update({"encoder." + k: v.detach() for k, v in e_leaf.items() if v.requires_grad})
    return losses, weights


def test_train_step_oracle_matches_reference_step_body_golden(golden_dir):
    """Two optimizer steps of the reference's own loop body (pack_padded_sequence + CrossEntropyLoss, utils.clip_gradient,
    torch.optim.Adam on the reference modules) against the oracle's restatement: losses and every updated tensor."""
    g = torch.load(os.path.join(golden_dir, "train_step.pt"))
    for kind in ("lstm", "transformer"):
        losses, weights = _oracle_train_steps(kind, g)
        ref = g[kind]
        assert max(abs(a - b) / abs(b) for a, b in zip(losses, ref["losses"])) < 2e-5, (losses, ref["losses"])
        assert set(weights) == set(ref["weights"]), set(weights) ^ set(ref["weights"])
        for k, d in ref["weights"].items():
            w = weights[k]
            assert abs(float(w.norm()) - float(d["norm"])) <= 1e-4 * float(d["norm"]) + 1e-9, (kind, k)
            # an Adam step moves an element by ~lr whatever its gradient: elements with a noise-level gradient
            # may move the other way, so compare the fraction that differs by more than a fifth of a step
            off = ((w.reshape(-1)[::1999] - d["sub"]).abs() > 0.2 * g["lr"]).float().mean()
            assert float(off) < 0.02, (kind, k, float(off))
