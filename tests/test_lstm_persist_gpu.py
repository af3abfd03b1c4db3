"""The persistent LSTM-attention recurrence kernels (csrc/lstm_persist.cu: one cooperative tcgen05 kernel per
direction) against (i) the per-step launch loop of the same library on identical inputs — every buffer both write —
and (ii) the CPU oracle (models/decoder.py:69-113 restated in oracle/decoder_oracle.py)."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu

V = 9490
BF16_TOL = 2e-2          # BASELINE.json north_star: logits within 2e-2 relative in bf16


def _dec(sd, persist, train=False):
    from imagecaptioningconvnext_b200 import DecoderWithAttention
    m = DecoderWithAttention(512, 512, 512, V, torch.device("cuda"), compute_dtype=torch.bfloat16)
    m.load_state_dict(sd)
    m = m.cuda()
    m.use_persist = persist
    return m.train() if train else m.eval()


def _run(m, enc, caps, lens, mask=None):
    """mask: dropout multipliers (B, T, D) in the CALLER's row order."""
    with torch.no_grad():
        out = m._tf_forward(enc.cuda(), caps.cuda(), lens.cuda(),
                            dropmask_unsorted=None if mask is None else mask.cuda())
    torch.cuda.synchronize()
    return out


def _unsort(x, sort_ind):
    """Rows back in the caller's order: a CUDA sort and a CPU sort may order equal caption lengths differently."""
    out = torch.empty_like(x)
    out[sort_ind.to(x.device)] = x
    return out


@pytest.mark.parametrize("B,cap_seed,full", [(32, 11, False), (5, 12, False), (32, 13, True), (1, 14, False)])
def test_forward_persist_matches_loop_and_oracle(B, cap_seed, full):
    from oracle import decoder_oracle as do
    sd = do.random_lstm_decoder_state(3, V)
    enc = do.synthetic_features(B, 21)
    caps, lens = do.synthetic_captions(B, cap_seed, V)
    if full:                                           # every caption at the maximum length: T = 51, bt = B throughout
        lens = torch.full_like(lens, 52)
        caps = caps.clone()
        caps[:, 1:51] = torch.randint(1, V - 4, (B, 50), generator=torch.Generator().manual_seed(5))
        caps[:, 51] = V - 1
    T = int(lens.max()) - 1
    mask = (torch.rand(B, T, 512, generator=torch.Generator().manual_seed(3)) > 0.5).float() / 0.5
    outs = {}
    for persist in (False, True):
        m = _dec(sd, persist, train=True)
        assert m._persist_ok(B, 49) == persist
        outs[persist] = _run(m, enc, caps, lens, mask)
    pl, pp = outs[False], outs[True]
    dl = pl[2]
    assert pp[2] == dl and torch.equal(pp[4], pl[4])
    # predictions and alphas: persistent kernel vs launch loop (both bf16 operands; the persistent kernel also keeps
    # att1 / enc in bf16 inside the attention), and both vs the fp32 oracle
    assert rel_err(pp[0], pl[0]) < BF16_TOL
    assert rel_err(pp[3], pl[3]) < BF16_TOL
    ref_sort = lens.squeeze(1).sort(dim=0, descending=True)[1]
    ref = do.lstm_teacher_forcing(sd, enc, caps, lens, dropmask=mask[ref_sort])   # the oracle takes SORTED-order masks
    assert ref[2] == dl
    assert rel_err(_unsort(pp[0], pp[4]), _unsort(ref[0], ref[4])) < BF16_TOL
    assert rel_err(_unsort(pp[3], pp[4]), _unsort(ref[3], ref[4])) < BF16_TOL
    for b, l in enumerate(dl):                         # exact zeros past each caption's decode length
        if l < pp[0].shape[1]:
            assert float(pp[0][b, l:].abs().max()) == 0.0 and float(pp[3][b, l:].abs().max()) == 0.0
    # the step buffers BPTT reads: only rows that were active at a step are defined
    sl, sp = pl[5], pp[5]
    act = torch.tensor([[dl[b] > t for b in range(B)] for t in range(T)], device="cuda")[:, :, None]
    live = lambda x: torch.where(act, x.float(), torch.zeros((), device="cuda"))   # inactive rows are undefined
    for name in ("HG", "G"):
        assert rel_err(live(sp[name]), live(sl[name])) < BF16_TOL, name
    assert rel_err(live(sp["C_all"][1:]), live(sl["C_all"][1:])) < BF16_TOL
    assert torch.equal(sp["C_all"][0], sl["C_all"][0])
    assert rel_err(sp["H_all"].hi.float(), sl["H_all"].hi.float()) < BF16_TOL
    assert rel_err(live(sp["XH"].hi[:T]), live(sl["XH"].hi[:T])) < BF16_TOL
    assert rel_err(live(sp["XH"].hi[1:T + 1, :, 1536:]), live(sl["XH"].hi[1:T + 1, :, 1536:])) < BF16_TOL
    # alphas rows sum to one on active steps (models/decoder.py:29)
    s = pp[3].sum(dim=2)
    for b, l in enumerate(dl):
        assert float((s[b, :l] - 1).abs().max()) < 1e-5


def test_forward_persist_is_deterministic():
    from oracle import decoder_oracle as do
    B = 32
    sd = do.random_lstm_decoder_state(4, V)
    enc = do.synthetic_features(B, 22)
    caps, lens = do.synthetic_captions(B, 15, V)
    m = _dec(sd, True)
    a = _run(m, enc, caps, lens)
    for _ in range(3):
        b = _run(m, enc, caps, lens)
        assert torch.equal(a[0], b[0]) and torch.equal(a[3], b[3])


def _grads(m, enc, caps, lens, mask, enc_grad):
    """One teacher-forced forward + backward of a loss with the train step's shape (trainMultiGPU.py:364-369:
    cross-entropy-like weighting of the logits + the alpha regulariser) -> {name: grad}, d encoder_out."""
    for p in m.parameters():
        p.grad = None
    e = enc.cuda().requires_grad_(enc_grad)
    from imagecaptioningconvnext_b200.decoder_train import lstm_teacher_forcing_with_grad
    preds, _, dl, alphas, sort_ind = lstm_teacher_forcing_with_grad(m, e, caps.cuda(), lens.cuda(),
                                                                    dropmask_unsorted=mask.cuda())
    w = torch.randn(preds.shape, generator=torch.Generator().manual_seed(9)).cuda()
    valid = torch.tensor([[1.0 if t < l else 0.0 for t in range(preds.shape[1])] for l in dl], device="cuda")
    # weights follow the CALLER's row order so that both runs see the same loss whatever the tie order of the sort
    w_sorted = w[sort_ind] * valid[:, :, None]
    loss = (preds * w_sorted).sum() / valid.sum() + ((1.0 - alphas.sum(dim=1)) ** 2).mean()
    loss.backward()
    torch.cuda.synchronize()
    g = {n: p.grad.detach().clone() for n, p in m.named_parameters() if p.grad is not None}
    return float(loss), g, (None if not enc_grad else e.grad.detach().clone())


def _cos(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))


@pytest.mark.parametrize("B,cap_seed,enc_grad", [(32, 31, True), (7, 32, False)])
def test_backward_persist_matches_loop(B, cap_seed, enc_grad):
    """Gradients of every parameter (and of encoder_out): persistent BPTT kernel vs the per-step launch loop."""
    from oracle import decoder_oracle as do
    sd = do.random_lstm_decoder_state(5, V)
    enc = do.synthetic_features(B, 23)
    caps, lens = do.synthetic_captions(B, cap_seed, V)
    T = int(lens.max()) - 1
    mask = (torch.rand(B, T, 512, generator=torch.Generator().manual_seed(4)) > 0.5).float() / 0.5
    res = {}
    for persist in (False, True):
        m = _dec(sd, persist, train=True)
        res[persist] = _grads(m, enc, caps, lens, mask, enc_grad)
    (l0, g0, e0), (l1, g1, e1) = res[False], res[True]
    assert abs(l0 - l1) < 2e-2 * abs(l0)
    assert set(g0) == set(g1)
    for n in g0:
        if n == "attention.full_att.bias":
            continue                                   # identically zero (softmax is shift invariant)
        c = _cos(g1[n], g0[n])
        r = float(g1[n].norm() / g0[n].norm().clamp_min(1e-30))
        assert c > 0.995 and 0.97 < r < 1.03, (n, c, r)
    if enc_grad:
        assert _cos(e1, e0) > 0.995


def test_backward_persist_vs_oracle_autograd():
    """Persistent forward + BPTT against torch autograd through the CPU oracle's restatement of
    models/decoder.py:69-113 (fp32): per-tensor gradient cosine."""
    from oracle import decoder_oracle as do
    B = 16
    sd = do.random_lstm_decoder_state(6, V)
    enc = do.synthetic_features(B, 24)
    caps, lens = do.synthetic_captions(B, 33, V)
    T = int(lens.max()) - 1
    mask = (torch.rand(B, T, 512, generator=torch.Generator().manual_seed(4)) > 0.5).float() / 0.5
    m = _dec(sd, True, train=True)
    l1, g1, e1 = _grads(m, enc, caps, lens, mask, True)
    # oracle: same loss under autograd
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    eg = enc.clone().requires_grad_(True)
    ref_sort = lens.squeeze(1).sort(dim=0, descending=True)[1]
    preds, _, dl, alphas, sort_ind = do.lstm_teacher_forcing(sdg, eg, caps, lens, dropmask=mask[ref_sort])
    w = torch.randn(preds.shape, generator=torch.Generator().manual_seed(9))
    valid = torch.tensor([[1.0 if t < l else 0.0 for t in range(preds.shape[1])] for l in dl])
    loss = (preds * (w[sort_ind] * valid[:, :, None])).sum() / valid.sum() + ((1.0 - alphas.sum(dim=1)) ** 2).mean()
    loss.backward()
    assert abs(l1 - float(loss)) < 2e-2 * abs(float(loss))
    for n, gr in g1.items():
        if n == "attention.full_att.bias":
            continue
        c = _cos(gr.cpu(), sdg[n].grad)
        assert c > 0.99, (n, c)
    assert _cos(e1.cpu(), eg.grad) > 0.99
