"""CapturedTrainStep (the whole train step replayed as one CUDA graph) against the eager ``caption_train_step`` on twin
models: same batches, dropout / stochastic depth off, six optimizer steps each (three eager warm-up steps, the
capture, two replays) -> same losses and same weights."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

V = 9490


def _models(kind, start):
    from imagecaptioningconvnext_b200 import DecoderWithAttention, Encoder, TransformerDecoder
    from imagecaptioningconvnext_b200.train_step import make_optimizers
    from synthetic import random_encoder_state, random_lstm_decoder_state, random_transformer_decoder_state
    enc = Encoder(compute_dtype=torch.bfloat16)
    enc.load_state_dict(random_encoder_state(seed=0, layer_scale=1.0))
    enc = enc.cuda().eval()                                   # eval: no stochastic depth
    enc.fine_tune(start is not None, start if start is not None else 7)
    if kind == "lstm":
        dec = DecoderWithAttention(512, 512, 512, V, torch.device("cuda"), compute_dtype=torch.bfloat16)
        dec.load_state_dict(random_lstm_decoder_state(1, V))
    else:
        dec = TransformerDecoder(512, 512, V, 52, torch.device("cuda"), None, None, True,
                                 compute_dtype=torch.bfloat16)
        dec.load_state_dict(random_transformer_decoder_state(1, V))
    dec = dec.cuda().train()
    dec.dropout_p = 0.0
    if kind != "lstm":
        for m in dec.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
    d_opt, e_opt = make_optimizers(enc, dec, decoder_lr=1e-3, encoder_lr=1e-4)
    return enc, dec, d_opt, e_opt


@pytest.mark.parametrize("kind,start,B", [("lstm", 7, 32), ("lstm", None, 8), ("transformer", None, 8)])
def test_captured_step_equals_eager_steps(kind, start, B):
    from imagecaptioningconvnext_b200.train_step import CapturedTrainStep, caption_train_step
    from synthetic import synthetic_captions, synthetic_images
    batches = []
    for i in range(6):
        caps, lens = synthetic_captions(B, 40 + i, V)
        batches.append((synthetic_images(B, 50 + i)[:, :, :64, :64].contiguous().cuda(), caps.cuda(), lens.cuda()))
    enc_a, dec_a, d_a, e_a = _models(kind, start)
    enc_b, dec_b, d_b, e_b = _models(kind, start)
    step = CapturedTrainStep(enc_b, dec_b, d_b, e_b, warmup_steps=3)
    la, lb = [], []
    for imgs, caps, lens in batches:
        la.append(float(caption_train_step(enc_a, dec_a, imgs, caps, lens, d_a, e_a)))
        lb.append(float(step(imgs, caps, lens)))
    assert step.graph is not None
    for x, y in zip(la, lb):
        assert abs(x - y) < 2e-3 * abs(x), (la, lb)
    # Adam normalises every element's gradient to O(1): where the gradient is noise-level, two runs may legitimately
    # step in opposite directions, so weights are compared against the size of an update (lr per step), not relative
    # to the weights: no element may be further apart than a few updates, and on average they must be much closer
    def close(pa, pb, lr, name):
        d = (pa.detach() - pb.detach()).abs()
        assert float(d.max()) <= 3.0 * lr * len(batches), (name, float(d.max()))
        assert float(d.mean()) <= 0.15 * lr * len(batches), (name, float(d.mean()))
    for (n, pa), (_, pb) in zip(dec_a.named_parameters(), dec_b.named_parameters()):
        close(pa, pb, 1e-3, n)
    if start is not None:
        for (n, pa), (_, pb) in zip(enc_a.convnext[7].named_parameters(), enc_b.convnext[7].named_parameters()):
            close(pa, pb, 1e-4, n)
    # optimizer state follows torch.optim.Adam's layout, step counts included
    sd = d_b.state_dict()
    assert float(sd["state"][0]["step"]) == 6.0
    # an eager forward after the replays sees the current weights (kernel-side copies are refreshed)
    with torch.no_grad():
        imgs, caps, lens = batches[0]
        fa, fb = enc_a(imgs), enc_b(imgs)
        pa = dec_a.eval()(teacherForcing=True, encoder_out=fa, encoded_captions=caps, caption_lengths=lens)[0]
        dec_b.fixed_T = False
        pb = dec_b.eval()(teacherForcing=True, encoder_out=fb, encoded_captions=caps, caption_lengths=lens)[0]
    assert rel_err(pb, pa) < 5e-2          # two independently trained twins (see above), same inputs
