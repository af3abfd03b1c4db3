"""Kernel-side weight copies after an optimizer step (trainMultiGPU.py:387-394 moves the fp32 masters; the kernels read
bf16 / re-laid-out copies): the one-launch refresh (``ccx_cast_segments`` through ``_host.RefreshPlan``) and the
in-place operand refresh of ``PreparedCache`` must leave exactly what a fresh ``_prepare()`` builds — bit for bit, in
the same buffers (captured CUDA graphs and weight tables hold their addresses)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

V = 1000


def _flat(x, out, path=""):
    from imagecaptioningconvnext_b200._lib import Operand
    if isinstance(x, dict):
        for k in x:
            _flat(x[k], out, f"{path}.{k}")
    elif isinstance(x, (list, tuple)):
        for i, v in enumerate(x):
            _flat(v, out, f"{path}[{i}]")
    elif isinstance(x, Operand):
        out.append((path + ".hi", x.hi))
        if x.lo is not None:
            out.append((path + ".lo", x.lo))
    elif torch.is_tensor(x):
        out.append((path, x))
    return out


def _decoder(kind, cd):
    from imagecaptioningconvnext_b200 import DecoderWithAttention, TransformerDecoder
    from synthetic import random_lstm_decoder_state, random_transformer_decoder_state
    dev = torch.device("cuda")
    if kind == "lstm":
        dec = DecoderWithAttention(512, 512, 512, V, dev, compute_dtype=cd)
        dec.load_state_dict(random_lstm_decoder_state(3, V))
    else:
        dec = TransformerDecoder(512, 512, V, 52, dev, None, None, True, compute_dtype=cd)
        dec.load_state_dict(random_transformer_decoder_state(3, V))
    return dec.cuda()


@pytest.mark.parametrize("kind", ["lstm", "transformer"])
@pytest.mark.parametrize("cd,plan", [(torch.bfloat16, True), (torch.bfloat16, False), (torch.float32, True)])
def test_refresh_equals_fresh_preparation(kind, cd, plan):
    from imagecaptioningconvnext_b200 import _host
    dec = _decoder(kind, cd)
    old_switch = _host.PLAN_REFRESH[0]
    _host.PLAN_REFRESH[0] = plan
    try:
        first = _flat(dec._cache.get(), [])
        ptrs = [t.data_ptr() for _, t in first]
        g = torch.Generator(device="cuda").manual_seed(5)
        for step in range(2):                       # the second refresh runs the cached plan / recorded sequence
            with torch.no_grad():
                for p in dec.parameters():
                    p.add_(torch.randn(p.shape, device=p.device, generator=g) * 0.05)
            got = _flat(dec._cache.get(), [])
            assert [t.data_ptr() for _, t in got] == ptrs, "a refresh must not move the kernel-side buffers"
            want = _flat(dec._prepare(), [])
            assert len(want) == len(got)
            for (name, a), (_, b) in zip(got, want):
                assert a.dtype == b.dtype and a.shape == b.shape, name
                assert torch.equal(a, b), f"{kind} {name} differs after refresh {step}"
        used_plan = isinstance(dec._cache._plan, _host.RefreshPlan)
        assert used_plan == (plan and cd == torch.bfloat16)
    finally:
        _host.PLAN_REFRESH[0] = old_switch


def test_cast_segments_odd_shapes_transpose_rowmap_and_sum():
    """The segment kernel directly: ragged sizes (scalar paths), a row permutation, a transposed piece written into a
    column window of a wider buffer, fp32 destination with two summed sources."""
    from imagecaptioningconvnext_b200._host import RefreshPlan
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(11)
    a = torch.randn(70, 37, device=dev, generator=g)
    b = torch.randn(130, 64, device=dev, generator=g)
    c = torch.randn(9, device=dev, generator=g)
    c2 = torch.randn(9, device=dev, generator=g)
    perm = torch.randperm(130, device=dev, generator=g)
    d_a = torch.zeros(70, 40, dtype=torch.bfloat16, device=dev)          # plain, odd width -> scalar stores
    d_at = torch.zeros(37, 200, dtype=torch.bfloat16, device=dev)        # transposed into columns [100, 170)
    d_b = torch.zeros(130, 64, dtype=torch.bfloat16, device=dev)         # row permutation
    d_bt = torch.zeros(64, 130, dtype=torch.bfloat16, device=dev)        # permuted and transposed
    d_c = torch.zeros(9, dtype=torch.float32, device=dev)
    plan = RefreshPlan()
    plan.add(d_a[:, :37], a).add(d_at[:, 100:170], a, transpose=True)
    plan.add(d_b, b, row_map=perm).add(d_bt, b, row_map=perm, transpose=True)
    plan.add(d_c, c, src2=c2)
    plan.run()
    torch.cuda.synchronize()
    assert torch.equal(d_a[:, :37], a.bfloat16()) and float(d_a[:, 37:].abs().max()) == 0.0
    assert torch.equal(d_at[:, 100:170], a.t().bfloat16())
    assert float(d_at[:, :100].abs().max()) == 0.0 and float(d_at[:, 170:].abs().max()) == 0.0
    assert torch.equal(d_b, b[perm].bfloat16())
    assert torch.equal(d_bt, b[perm].t().bfloat16())
    assert torch.equal(d_c, c + c2)
    assert ctypes.sizeof(__import__("imagecaptioningconvnext_b200")._lib.CastSeg) == 64
