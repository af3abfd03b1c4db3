"""GPU parity of the two caption decoders (libccx through the C ABI) vs the CPU oracle and the reference goldens."""
import os

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

V = 9490
WORDMAP = {"<pad>": 0, "<unk>": V - 3, "<start>": V - 2, "<end>": V - 1}
TOL = {torch.float32: 1e-3, torch.bfloat16: 2e-2}   # BASELINE.json north_star tolerances for logits
TAU = 1e-4                                          # argmax asserted where the oracle's top-2 gap > TAU


def _lstm(sd, dtype):
    from imagecaptioningconvnext_b200 import DecoderWithAttention
    m = DecoderWithAttention(512, 512, 512, V, torch.device("cuda"), compute_dtype=dtype)
    m.load_state_dict(sd)
    return m.cuda().eval()


def _transformer(sd, dtype):
    from imagecaptioningconvnext_b200 import TransformerDecoder
    m = TransformerDecoder(512, 512, V, 52, torch.device("cuda"), None, None, True, compute_dtype=dtype)
    m.load_state_dict(sd)
    return m.cuda().eval()


def _greedy_tokens_agree(seq_gpu, seq_ref, preds_ref):
    """Token-exact wherever the oracle's decision was not a near-tie; after the first near-tie mismatch in a row the
    sequences may legitimately diverge (free-running), so the row is only checked up to there."""
    top2 = preds_ref.topk(2, dim=-1).values
    gap = top2[..., 0] - top2[..., 1]
    checked, exits = 0, 0
    for b in range(seq_ref.shape[0]):
        for t in range(seq_ref.shape[1]):
            if seq_gpu[b, t] != seq_ref[b, t]:
                assert gap[b, t] <= TAU, f"row {b} step {t}: token differs with oracle gap {float(gap[b, t]):.3e}"
                exits += 1
                break
            checked += 1
    return checked, exits


def _assert_greedy(seq_gpu, seq_ref, preds_ref, max_near_tie_rows=1):
    """-> number of rows that left the comparison at a labelled near-tie (oracle top-2 gap <= TAU).  Every other row
    must be token-exact over all steps; at most `max_near_tie_rows` rows may leave."""
    n, exits = _greedy_tokens_agree(seq_gpu, seq_ref, preds_ref)
    B, T = seq_ref.shape
    assert exits <= max_near_tie_rows, f"{exits} rows diverged at near-ties"
    assert n >= (B - exits) * T, (n, exits)
    return exits


def test_state_dict_keys_match_reference_layout():
    from oracle import decoder_oracle as do
    m = _lstm(do.random_lstm_decoder_state(0, V), torch.float32)
    assert set(m.state_dict()) == set(do.random_lstm_decoder_state(0, V))
    t = _transformer(do.random_transformer_decoder_state(0, V), torch.float32)
    assert set(t.state_dict()) == set(do.random_transformer_decoder_state(0, V))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_lstm_teacher_forcing_vs_oracle_and_golden(golden_dir, dtype):
    from oracle import decoder_oracle as do
    gold = torch.load(os.path.join(golden_dir, "lstm_decoder.pt"))["tf"]
    sd = do.random_lstm_decoder_state(gold["weight_seed"], V, end_bias=gold["end_bias"])
    enc = do.synthetic_features(gold["B"], gold["feat_seed"])
    caps, lens = do.synthetic_captions(gold["B"], gold["cap_seed"], V)
    ref = do.lstm_teacher_forcing(sd, enc, caps, lens)
    m = _lstm(sd, dtype)
    with torch.no_grad():
        preds, caps_s, dl, alphas, sort_ind = m(teacherForcing=True, encoder_out=enc.cuda(),
                                                encoded_captions=caps.cuda(), caption_lengths=lens.cuda())
    assert dl == ref[2] == gold["decode_lengths"] and isinstance(dl, list)
    assert torch.equal(sort_ind.cpu(), ref[4]) and torch.equal(caps_s.cpu(), ref[1])
    tol = TOL[dtype]
    assert rel_err(preds, ref[0]) < tol
    assert rel_err(alphas, ref[3]) < tol
    assert rel_err(preds.cpu()[..., ::31], gold["preds"]["sub"]) < tol
    assert rel_err(alphas, gold["alphas"]) < tol
    for b, l in enumerate(dl):   # exact zeros past each caption's decode length (models/decoder.py:94-95,110-111)
        if l < preds.shape[1]:
            assert float(preds[b, l:].abs().max()) == 0.0 and float(alphas[b, l:].abs().max()) == 0.0
    loss = do.train_loss_lstm(preds.cpu(), caps_s.cpu(), dl, alphas.cpu())
    assert abs(float(loss) - float(gold["loss"])) < tol * 10


def test_lstm_teacher_forcing_train_mode_injected_dropout():
    from oracle import decoder_oracle as do
    sd = do.random_lstm_decoder_state(2, V)
    B = 6
    enc = do.synthetic_features(B, 7)
    caps, lens = do.synthetic_captions(B, 8, V)
    T = int(lens.max()) - 1
    mask = (torch.rand(B, T, 512, generator=torch.Generator().manual_seed(3)) > 0.5).float() / 0.5
    ref = do.lstm_teacher_forcing(sd, enc, caps, lens, dropmask=mask)
    m = _lstm(sd, torch.float32).train()
    m.inject_dropmask = mask
    with torch.no_grad():
        preds = m(teacherForcing=True, encoder_out=enc.cuda(), encoded_captions=caps.cuda(),
                  caption_lengths=lens.cuda())[0]
    assert rel_err(preds, ref[0]) < 1e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_lstm_greedy_vs_oracle_and_golden(golden_dir, dtype):
    from oracle import decoder_oracle as do
    g = torch.load(os.path.join(golden_dir, "lstm_decoder.pt"))
    sd = do.random_lstm_decoder_state(g["tf"]["weight_seed"], V, end_bias=g["tf"]["end_bias"])
    enc = do.synthetic_features(g["tf"]["B"], g["tf"]["feat_seed"])
    rp, ra, rs = do.lstm_greedy(sd, enc, V - 2, V - 1, 51)
    m = _lstm(sd, dtype)
    preds, alphas, seqs = m(teacherForcing=False, encoder_out=enc.cuda(), wordMap=WORDMAP, maxDecodeLen=51)
    assert preds.shape == (5, 51, V) and alphas.shape == (5, 51, 49) and seqs.dtype == torch.long
    if dtype == torch.float32:
        if _assert_greedy(seqs.cpu(), rs, rp) == 0:
            assert torch.equal(seqs.cpu(), rs) and torch.equal(seqs.cpu(), g["greedy"]["sequences"])
            assert rel_err(preds, rp) < 1e-3 and rel_err(alphas, ra) < 1e-3
    # first step is prefix-independent: logits comparable in both dtypes
    assert rel_err(preds[:, 0], rp[:, 0]) < TOL[dtype]
    # finished rows stay exactly zero (models/decoder.py:141-154)
    for b in range(5):
        row = seqs[b].tolist()
        if V - 1 in row:
            e = row.index(V - 1)
            assert float(preds[b, e + 1:].abs().max()) == 0.0 and float(alphas[b, e + 1:].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_transformer_teacher_forcing_vs_oracle_and_golden(golden_dir, dtype):
    from oracle import decoder_oracle as do
    gold = torch.load(os.path.join(golden_dir, "transformer_decoder.pt"))["tf"]
    sd = do.random_transformer_decoder_state(gold["weight_seed"], V, end_bias=gold["end_bias"])
    enc = do.synthetic_features(gold["B"], gold["feat_seed"])
    caps, lens = do.synthetic_captions(gold["B"], gold["cap_seed"], V)
    ref, _, rdl = do.transformer_teacher_forcing(sd, enc, caps, lens, caps == 0)
    m = _transformer(sd, dtype)
    with torch.no_grad():
        preds, caps_o, dl = m(teacherForcing=True, encoder_out=enc.cuda(), encoded_captions=caps.cuda(),
                              caption_lengths=lens.cuda(), tgt_key_padding_mask=(caps == 0).cuda())
    assert dl == rdl == gold["decode_lengths"] and torch.equal(caps_o.cpu(), caps)
    assert preds.shape == (gold["B"], 52, V)
    assert rel_err(preds, ref) < TOL[dtype]
    assert rel_err(preds.cpu()[..., ::31], gold["preds"]["sub"]) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_transformer_greedy_kv_cache_vs_oracle_prefix_recompute(golden_dir, dtype):
    from oracle import decoder_oracle as do
    g = torch.load(os.path.join(golden_dir, "transformer_decoder.pt"))
    sd = do.random_transformer_decoder_state(g["tf"]["weight_seed"], V, end_bias=g["tf"]["end_bias"])
    enc = do.synthetic_features(g["tf"]["B"], g["tf"]["feat_seed"])
    rp, rs = do.transformer_greedy(sd, enc, V - 2, V - 1, 0, 51)
    m = _transformer(sd, dtype)
    preds, seqs = m(teacherForcing=False, encoder_out=enc.cuda(), wordMap=WORDMAP, maxDecodeLen=51)
    assert preds.shape == (4, 51, V) and seqs.shape == (4, 51)
    assert rel_err(preds[:, 0], rp[:, 0]) < TOL[dtype]
    if dtype == torch.float32:
        if _assert_greedy(seqs.cpu(), rs, rp) == 0:
            assert torch.equal(seqs.cpu(), rs) and torch.equal(seqs.cpu(), g["greedy"]["sequences"])
            assert rel_err(preds, rp) < 1e-3


def test_poked_submodules_run_on_libccx():
    """caption.py pokes decoder.attention / f_beta / decode_step / fc / embedding directly."""
    from oracle import decoder_oracle as do
    sd = do.random_lstm_decoder_state(4, V)
    m = _lstm(sd, torch.float32)
    enc = do.synthetic_features(3, 9).view(3, 49, 1024)
    h0, c0 = do.init_hidden_state(sd, enc)
    h, c = m.init_hidden_state(enc.cuda())
    assert rel_err(h, h0) < 1e-4 and rel_err(c, c0) < 1e-4
    awe_r, alpha_r = do.attention(sd, enc, h0)
    awe, alpha = m.attention(enc.cuda(), h)
    assert rel_err(awe, awe_r) < 1e-4 and rel_err(alpha, alpha_r) < 1e-4
    tok = torch.tensor([[5], [17], [V - 2]])
    emb = m.embedding(tok.cuda()).squeeze(1)
    assert torch.equal(emb.cpu(), sd["embedding.weight"][tok.squeeze(1)])
    gate = torch.sigmoid(m.f_beta(h))
    x = torch.cat([emb, gate * awe], dim=1)
    h2, c2 = m.decode_step(x, (h, c))
    hr, cr, _ = do.lstm_step(sd, enc, sd["embedding.weight"][tok.squeeze(1)], h0, c0)
    assert rel_err(h2, hr) < 1e-4 and rel_err(c2, cr) < 1e-4
    assert rel_err(m.fc(h2), torch.nn.functional.linear(hr, sd["fc.weight"], sd["fc.bias"])) < 1e-4


def test_poked_transformer_submodules_run_on_libccx():
    """caption.py:181,204-216 pokes encoder_proj / embedding / pos_encoding / transformer_decoder / fc_out directly."""
    import torch.nn as nn
    from oracle import decoder_oracle as do
    sd = do.random_transformer_decoder_state(6, V)
    m = _transformer(sd, torch.float32)
    enc = do.synthetic_features(2, 11).view(2, 49, 1024)
    toks = torch.randint(1, V - 4, (2, 9), generator=torch.Generator().manual_seed(2))
    mem = m.encoder_proj(enc.cuda()).permute(1, 0, 2)                       # (P, B, D) as caption.py:181
    emb = m.pos_encoding(m.dropout(m.embedding(toks.cuda())))
    tgt = emb.permute(1, 0, 2)
    mask = nn.Transformer.generate_square_subsequent_mask(9).cuda().bool()
    out = m.transformer_decoder(tgt, mem, tgt_mask=mask)
    logits = m.fc_out(out[-1])
    ref = do.transformer_last_logits(sd, do.transformer_memory(sd, enc), toks)
    assert rel_err(logits, ref) < 1e-3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_attention_viz_decoder_vs_oracle_and_golden(golden_dir, dtype):
    """TransformerDecoderForAttentionViz: reference key names, 4-/3-tuple outputs, attention
    maps from the fused attention kernel against the oracle and the reference-run golden."""
    from oracle import decoder_oracle as do
    from imagecaptioningconvnext_b200 import TransformerDecoderForAttentionViz
    g = torch.load(os.path.join(golden_dir, "attvis.pt"))
    sd = do.random_transformer_decoder_state(g["weight_seed"], V, end_bias=g["end_bias"])
    m = TransformerDecoderForAttentionViz(512, 512, V, 52, torch.device("cuda"), compute_dtype=dtype)
    assert sorted(m.state_dict().keys()) == g["state_dict_keys"]
    m.load_state_dict({k.replace("transformer_decoder.layers.", "decoder_layers."): v for k, v in sd.items()})
    m = m.cuda().eval()
    enc = do.synthetic_features(g["B"], g["feat_seed"])
    caps, lens = do.synthetic_captions(g["B"], g["cap_seed"], V)
    with torch.no_grad():
        preds, caps_o, dl, alphas = m(teacherForcing=True, encoder_out=enc.cuda(), encoded_captions=caps.cuda(),
                                      caption_lengths=lens.cuda(), tgt_key_padding_mask=(caps == 0).cuda())
        rp, _, _, ra = do.transformer_teacher_forcing(sd, enc, caps, lens, caps == 0, return_alphas=True)
    assert dl == g["tf"]["decode_lengths"]
    assert rel_err(preds, rp) < TOL[dtype]
    assert alphas.shape == (8, g["B"], 49)
    assert rel_err(alphas, ra) < TOL[dtype] and rel_err(alphas, g["tf"]["alphas"]) < TOL[dtype]
    gp, gs, ga = m(teacherForcing=False, encoder_out=enc.cuda(), wordMap=WORDMAP, maxDecodeLen=51)
    with torch.no_grad():
        op, os_, oa = do.transformer_greedy(sd, enc, V - 2, V - 1, 0, 51, return_alphas=True)
    _greedy_tokens_agree(gs.cpu(), os_, op)
    if torch.equal(gs.cpu(), os_):
        assert rel_err(ga, oa) < TOL[dtype] and rel_err(ga, g["greedy"]["alphas"]) < TOL[dtype]
        assert torch.equal(ga.cpu() == 0, oa == 0)           # maps stay zero after a row's <end>
    else:
        assert dtype == torch.bfloat16
    # rows of an active step are probability distributions (mean of softmaxes)
    live = ga.sum(-1)
    assert bool(((live - 1).abs() < 1e-3)[live > 0].all())
    # training forward (autograd) returns the maps too, detached
    m.train()
    m.dropout_p = 0.0
    preds_t, _, _, al_t = m(teacherForcing=True, encoder_out=enc.cuda().requires_grad_(True),
                            encoded_captions=caps.cuda(), caption_lengths=lens.cuda(),
                            tgt_key_padding_mask=(caps == 0).cuda())
    assert preds_t.requires_grad and not al_t.requires_grad
    assert rel_err(al_t, ra) < TOL[dtype]
