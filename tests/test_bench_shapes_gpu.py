"""Module-level parity AT THE BENCHMARK SHAPES (BASELINE.json configs[1] and configs[3]) against the CPU oracle:
the kernels / tilings that dominate bench.py (gemm_tn_kernel<256>, the 64x64 stage-1 dwconv tiles, the persistent LSTM
recurrence at B = 32) run here on exactly the sizes that are timed there."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

V = 9490


def _cos(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))


def test_encoder_forward_batch64_256x256_vs_oracle():
    """configs[1]: Encoder.forward, batch 64, 256x256, fp32 (3xTF32) within 1e-3 and bf16 within 2e-2 of the oracle."""
    from imagecaptioningconvnext_b200 import Encoder
    from oracle.encoder_oracle import encoder_forward
    from synthetic import random_encoder_state, synthetic_images
    torch.set_num_threads(max(1, torch.get_num_threads()))
    sd = random_encoder_state(seed=0, layer_scale=1.0)
    x = synthetic_images(64, 1234)
    with torch.no_grad():
        ref = encoder_forward(sd, x, 7)                       # CPU fp32, tens of seconds
    for dtype, tol in ((torch.float32, 1e-3), (torch.bfloat16, 2e-2)):
        e = Encoder(compute_dtype=dtype)
        e.load_state_dict(sd)
        e = e.cuda().eval()
        with torch.no_grad():
            y = e(x.cuda())
        assert y.shape == (64, 7, 7, 1024)
        err = rel_err(y, ref)
        print("Encoder.forward B=64", dtype, "rel err", err)
        assert err < tol, (dtype, err)
        del e
        torch.cuda.empty_cache()


def test_train_step_batch32_loss_and_gradients_vs_oracle():
    """configs[3] at the benchmark shape: encoder fine-tuned from child 7 + LSTM-attention decoder, bf16, batch 32,
    256x256 images, captions 7..52 tokens — loss within 2e-2 of the oracle's, every parameter gradient (decoder and
    fine-tuned encoder stage) with cosine >= 0.99 against torch autograd through the oracle (dropout / stochastic
    depth off so that both sides see the same function)."""
    from imagecaptioningconvnext_b200 import DecoderWithAttention, Encoder
    from imagecaptioningconvnext_b200.losses import packed_cross_entropy
    from oracle import decoder_oracle as do
    from oracle import encoder_oracle as eo
    from synthetic import random_encoder_state, random_lstm_decoder_state, synthetic_captions, synthetic_images
    B = 32
    esd = random_encoder_state(seed=0, layer_scale=1.0)
    dsd = random_lstm_decoder_state(0, V)
    imgs = synthetic_images(B, 1234)
    caps, lens = synthetic_captions(B, 7, V)
    # oracle: autograd through child 7 and the decoder
    e_leaf = {k: v.clone().requires_grad_(k.startswith("convnext.7.")) for k, v in esd.items()}
    d_leaf = {k: v.clone().requires_grad_(True) for k, v in dsd.items()}
    feats = eo.encoder_forward(e_leaf, imgs, 7)
    p, cs, dl, al, _ = do.lstm_teacher_forcing(d_leaf, feats, caps, lens)
    ref_loss = do.train_loss_lstm(p, cs, dl, al)
    ref_loss.backward()
    # ours
    enc = Encoder(compute_dtype=torch.bfloat16)
    enc.load_state_dict(esd)
    enc = enc.cuda().eval()
    enc.fine_tune(True, 7)
    dec = DecoderWithAttention(512, 512, 512, V, torch.device("cuda"), compute_dtype=torch.bfloat16)
    dec.load_state_dict(dsd)
    dec = dec.cuda().train()
    dec.dropout_p = 0.0
    assert dec._persist_ok(B, 49)                              # the persistent recurrence kernels serve this shape
    f = enc(imgs.cuda())
    s, cs2, dl2, al2, _ = dec(teacherForcing=True, encoder_out=f, encoded_captions=caps.cuda(),
                              caption_lengths=lens.cuda())
    loss = packed_cross_entropy(s, cs2, dl2) + ((1.0 - al2.sum(dim=1)) ** 2).mean()
    loss.backward()
    assert sorted(dl2) == sorted(dl)
    assert abs(float(loss) - float(ref_loss)) < 2e-2 * abs(float(ref_loss)), (float(loss), float(ref_loss))
    worst = []
    for n, prm in dec.named_parameters():
        if n == "attention.full_att.bias":
            continue
        worst.append((_cos(prm.grad, d_leaf[n].grad), "dec." + n))
    for n, prm in enc.named_parameters():
        if prm.requires_grad:
            worst.append((_cos(prm.grad, e_leaf[n].grad), "enc." + n))
    worst.sort()
    print("lowest gradient cosines:", worst[:5])
    assert worst[0][0] >= 0.99, worst[:5]
