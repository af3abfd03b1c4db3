"""Fused loss of the train step: ``CrossEntropyLoss`` over the rows ``pack_padded_sequence`` would select
(trainMultiGPU.py:365-367 / :375-377), forward + backward in one kernel (``ccx_softmax_ce``), no packed copy."""
import torch

from . import _lib
from ._lib import ptr


class _PackedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, targets, n_valid):
        B, T, V = scores.shape
        s2 = scores.contiguous().view(B * T, V)
        loss = torch.zeros(1, dtype=torch.float32, device=scores.device)
        need_grad = scores.requires_grad
        dlogits = torch.empty_like(s2) if need_grad else None
        _lib.check(_lib.lib().ccx_softmax_ce(ptr(s2), V, ptr(targets), B * T, V, 1.0 / n_valid, ptr(loss),
                                             ptr(dlogits), V, None, _lib.stream_ptr()), "softmax_ce")
        ctx.dlogits, ctx.shape = dlogits, scores.shape
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        return ctx.dlogits.view(ctx.shape) * g, None, None


def packed_cross_entropy(scores, captions, decode_lengths):
    """scores (B, T, V) logits; captions (B, Tc) token ids aligned with `scores` rows; targets are captions[:, 1:]
    restricted to the first decode_lengths[b] positions of row b (the reference's pack_padded_sequence + CE)."""
    B, T, V = scores.shape
    dev = scores.device
    dl = torch.as_tensor(decode_lengths, device=dev)
    tgt = captions[:, 1:T + 1].to(torch.long)
    if tgt.shape[1] < T:                                   # Transformer: T = 52 positions, targets exist for 51
        tgt = torch.nn.functional.pad(tgt, (0, T - tgt.shape[1]), value=0)
    valid = torch.arange(T, device=dev).unsqueeze(0) < dl.unsqueeze(1)
    targets = torch.where(valid, tgt, torch.full_like(tgt, -1)).contiguous().view(-1)
    return _PackedCE.apply(scores, targets, float(sum(decode_lengths)))
