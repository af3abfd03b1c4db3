"""GPU parity on the edges of the input domain: single sample, shortest / longest captions, encoded_image_size 14
(196 pixels: SURVEY.md H1 — north_star quotes (B,14,14,C) while every reference call site uses 7), batch of one beam."""
import pytest
import torch

from conftest import rel_err
from test_decoders_gpu import V, WORDMAP, _lstm, _transformer

pytestmark = pytest.mark.gpu


def _caps(lengths):
    B, T = len(lengths), 52
    g = torch.Generator().manual_seed(sum(lengths))
    caps = torch.zeros(B, T, dtype=torch.long)
    for b, L in enumerate(lengths):
        caps[b, 0] = V - 2
        if L > 2:
            caps[b, 1:L - 1] = torch.randint(1, V - 3, (L - 2,), generator=g)
        caps[b, L - 1] = V - 1
    return caps, torch.tensor(lengths).view(-1, 1)


@pytest.mark.parametrize("lengths", [[52], [2], [52, 2, 3, 52, 17], [5, 5, 5]])
def test_lstm_teacher_forcing_length_extremes(lengths):
    from oracle import decoder_oracle as do
    sd = do.random_lstm_decoder_state(11, V)
    enc = do.synthetic_features(len(lengths), 70)
    caps, lens = _caps(lengths)
    ref = do.lstm_teacher_forcing(sd, enc, caps, lens)
    m = _lstm(sd, torch.float32)
    with torch.no_grad():
        out = m(teacherForcing=True, encoder_out=enc.cuda(), encoded_captions=caps.cuda(), caption_lengths=lens.cuda())
    assert out[2] == ref[2] and out[0].shape == ref[0].shape
    # equal lengths: torch.sort(descending) may order ties differently on CUDA and on the CPU (the reference has the
    # same freedom), so rows are compared in ORIGINAL sample order through each side's own sort_ind
    inv_g, inv_r = torch.argsort(out[4].cpu()), torch.argsort(ref[4])
    assert rel_err(out[0].cpu()[inv_g], ref[0][inv_r]) < 1e-3 and rel_err(out[3].cpu()[inv_g], ref[3][inv_r]) < 1e-3
    assert torch.equal(out[1].cpu()[inv_g], ref[1][inv_r])


@pytest.mark.parametrize("lengths", [[52], [2], [52, 2, 9]])
def test_transformer_teacher_forcing_length_extremes(lengths):
    from oracle import decoder_oracle as do
    sd = do.random_transformer_decoder_state(12, V)
    enc = do.synthetic_features(len(lengths), 71)
    caps, lens = _caps(lengths)
    ref, _, dl = do.transformer_teacher_forcing(sd, enc, caps, lens, caps == 0)
    m = _transformer(sd, torch.float32)
    with torch.no_grad():
        preds, _, dl2 = m(teacherForcing=True, encoder_out=enc.cuda(), encoded_captions=caps.cuda(),
                          caption_lengths=lens.cuda(), tgt_key_padding_mask=(caps == 0).cuda())
    assert dl2 == dl
    assert rel_err(preds, ref) < 1e-3


def test_decoders_with_encoded_image_size_14():
    """196 pixels instead of 49: attention-over-pixels, cross-attention and the LSTM backward at P=196."""
    from oracle import decoder_oracle as do
    B = 3
    enc = do.synthetic_features(B, 72, P=196)
    assert enc.shape == (B, 14, 14, 1024)
    caps, lens = do.synthetic_captions(B, 73, V)
    lsd = do.random_lstm_decoder_state(13, V, end_bias=0.21)
    ref = do.lstm_teacher_forcing(lsd, enc, caps, lens)
    m = _lstm(lsd, torch.float32)
    with torch.no_grad():
        out = m(teacherForcing=True, encoder_out=enc.cuda(), encoded_captions=caps.cuda(), caption_lengths=lens.cuda())
        gp, ga, gs = m(teacherForcing=False, encoder_out=enc.cuda(), wordMap=WORDMAP, maxDecodeLen=10)
    assert out[3].shape == (B, max(ref[2]), 196)
    assert rel_err(out[0], ref[0]) < 1e-3 and rel_err(out[3], ref[3]) < 1e-3
    rp, ra, rs = do.lstm_greedy(lsd, enc, V - 2, V - 1, 10)
    assert rel_err(gp[:, 0], rp[:, 0]) < 1e-3 and ga.shape == (B, 10, 196)
    # gradients at P=196 (register-spilling variant of the attention backward)
    enc_leaf = enc.clone().requires_grad_(True)
    leaf = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in lsd.items()}
    p, cs, dl, al, _ = do.lstm_teacher_forcing(leaf, enc_leaf, caps, lens)
    do.train_loss_lstm(p, cs, dl, al).backward()
    m.train()
    m.dropout_p = 0.0
    enc_g = enc.cuda().requires_grad_(True)
    p2, cs2, dl2, al2, _ = m(teacherForcing=True, encoder_out=enc_g, encoded_captions=caps.cuda(),
                             caption_lengths=lens.cuda())
    do.train_loss_lstm(p2, cs2, dl2, al2).backward()
    assert rel_err(enc_g.grad, enc_leaf.grad) < 5e-3
    assert rel_err(m.attention.encoder_att.weight.grad, leaf["attention.encoder_att.weight"].grad) < 5e-3
    tsd = do.random_transformer_decoder_state(14, V)
    tref, _, _ = do.transformer_teacher_forcing(tsd, enc, caps, lens, caps == 0)
    t = _transformer(tsd, torch.float32)
    with torch.no_grad():
        tp, _, _ = t(teacherForcing=True, encoder_out=enc.cuda(), encoded_captions=caps.cuda(),
                     caption_lengths=lens.cuda(), tgt_key_padding_mask=(caps == 0).cuda())
    assert rel_err(tp, tref) < 1e-3


def test_single_image_single_beam_and_wide_beam():
    from imagecaptioningconvnext_b200.beam import beam_search_lstm, beam_search_transformer
    from oracle import decoder_oracle as do
    lsd = do.random_lstm_decoder_state(0, V, end_bias=0.21)
    tsd = do.random_transformer_decoder_state(0, V, end_bias=3.2)
    feats = do.synthetic_features(1, 300)
    for k in (1, 8):                                  # beam widths 1 (caption.py's default) and 8 (kernel maximum)
        got = beam_search_lstm(_lstm(lsd, torch.float32), feats.cuda(), WORDMAP, beamSize=k)[0]
        assert got == do.beam_search(lsd, feats, "lstm", k, V - 2, V - 1, V)[0]
        got = beam_search_transformer(_transformer(tsd, torch.float32), feats.cuda(), WORDMAP, beamSize=k)[0]
        assert got == do.beam_search(tsd, feats, "transformer", k, V - 2, V - 1, V)[0]
    with pytest.raises(RuntimeError):                # beam width above the kernel maximum fails loudly
        beam_search_lstm(_lstm(lsd, torch.float32), feats.cuda(), WORDMAP, beamSize=9)


def test_unsupported_inputs_fail_loudly():
    from imagecaptioningconvnext_b200 import Encoder
    e = Encoder().cuda().eval()
    with pytest.raises(ValueError):
        e(torch.zeros(1, 3, 100, 100, device="cuda"))          # not a multiple of 32
    with pytest.raises(ValueError):
        e(torch.zeros(1, 4, 64, 64, device="cuda"))            # not 3 channels
    with pytest.raises(ValueError):
        e(torch.zeros(1, 3, 64, 64, device="cuda", dtype=torch.float16))
