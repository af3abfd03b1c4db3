"""GPU parity of the train-step gradients (libccx backward kernels) against torch autograd on the CPU oracle."""
import pytest
import torch

from conftest import rel_err
from test_decoders_gpu import V, _lstm, _transformer

pytestmark = pytest.mark.gpu
GRAD_TOL = {torch.float32: 2e-3, torch.bfloat16: 6e-2}


def _oracle_grads(sd, loss_fn):
    leaf = {k: v.clone().requires_grad_(v.is_floating_point() and k != "pos_encoding.pe") for k, v in sd.items()}
    loss = loss_fn(leaf)
    loss.backward()
    return float(loss), {k: v.grad for k, v in leaf.items() if v.requires_grad and v.grad is not None}


def _compare_grads(model, ref_grads, tol, skip=()):
    worst = ("", 0.0)
    for n, p in model.named_parameters():
        if n in skip or n not in ref_grads:
            continue
        assert p.grad is not None, f"no grad for {n}"
        e = rel_err(p.grad, ref_grads[n])
        if e > worst[1]:
            worst = (n, e)
    print("worst grad rel err", worst)
    assert worst[1] < tol, worst


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("train_mode", [False, True])
def test_transformer_teacher_forcing_gradients(dtype, train_mode):
    from oracle import decoder_oracle as do
    torch.manual_seed(0)
    sd = do.random_transformer_decoder_state(3, V)
    B, T, Pn, D, H = 3, 52, 49, 512, 8
    enc = do.synthetic_features(B, 21)
    caps, lens = do.synthetic_captions(B, 22, V)
    kpm = caps == 0
    drop_ref, drop_inj = None, None
    if train_mode:
        g = torch.Generator().manual_seed(5)
        mk = lambda *s: (torch.rand(*s, generator=g) > 0.5).float() * 2.0
        drop_ref = {"emb": mk(B, T, D)}
        for l in range(6):
            drop_ref.update({(l, "sa_p"): mk(B, H, T, T), (l, "d1"): mk(B, T, D), (l, "ca_p"): mk(B, H, T, Pn),
                             (l, "d2"): mk(B, T, D), (l, "ff"): mk(B, T, D), (l, "d3"): mk(B, T, D)})
        drop_inj = {k: (v.reshape(B * T, -1) if v.dim() == 3 else v) for k, v in drop_ref.items()}
    enc_leaf = enc.clone().requires_grad_(True)

    def loss_fn(leaf):
        preds, _, dl = do.transformer_teacher_forcing(leaf, enc_leaf, caps, lens, kpm, drop=drop_ref)
        return do.train_loss_transformer(preds, caps, dl)

    ref_loss, ref_grads = _oracle_grads(sd, loss_fn)
    m = _transformer(sd, dtype)
    m.train(train_mode)
    m.inject_dropout = drop_inj
    if not train_mode:
        m.dropout_p = 0.0     # eval-mode autograd: no dropout anywhere
    enc_g = enc.cuda().requires_grad_(True)
    preds, _, dl = m(teacherForcing=True, encoder_out=enc_g, encoded_captions=caps.cuda(),
                     caption_lengths=lens.cuda(), tgt_key_padding_mask=kpm.cuda())
    loss = do.train_loss_transformer(preds, caps.cuda(), dl)
    loss.backward()
    tol = GRAD_TOL[dtype]
    assert abs(float(loss) - ref_loss) < tol * 5
    _compare_grads(m, ref_grads, tol)
    assert rel_err(enc_g.grad, enc_leaf.grad) < tol
