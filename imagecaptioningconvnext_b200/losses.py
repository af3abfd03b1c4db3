"""Fused loss of the train step: ``CrossEntropyLoss`` over the rows ``pack_padded_sequence`` would select
(trainMultiGPU.py:365-367 / :375-377), forward + backward in one kernel (``ccx_softmax_ce``), no packed copy."""
import torch

from . import _lib
from ._host import device_twin
from ._lib import ptr


class _PackedCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, scores, targets, n_valid, unit_grad):
        """n_valid: number of scored rows — a python number, or a float32 DEVICE tensor (then nothing about the
        caption lengths is needed on the host: CUDA-graph replay).  unit_grad: the caller guarantees that the loss
        enters ``backward()`` with gradient 1 (``(ce + regulariser).backward()`` as in trainMultiGPU.py:367-384), so
        the stored d logits are returned as they are instead of being multiplied by the incoming gradient (a full
        extra pass over the B x T x V tensor)."""
        B, T, V = scores.shape
        s2 = scores.contiguous().view(B * T, V)
        loss = torch.zeros(1, dtype=torch.float32, device=scores.device)
        need_grad = scores.requires_grad
        dlogits = torch.empty_like(s2) if need_grad else None
        if torch.is_tensor(n_valid):
            _lib.check(_lib.lib().ccx_softmax_ce_dev(ptr(s2), V, ptr(targets), B * T, V, ptr(n_valid), ptr(loss),
                                                     ptr(dlogits), V, None, 0, _lib.stream_ptr()), "softmax_ce_dev")
        else:
            _lib.check(_lib.lib().ccx_softmax_ce(ptr(s2), V, ptr(targets), B * T, V, 1.0 / n_valid, ptr(loss),
                                                 ptr(dlogits), V, None, 0, _lib.stream_ptr()), "softmax_ce")
        ctx.dlogits, ctx.shape, ctx.unit_grad = dlogits, scores.shape, unit_grad
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        d = ctx.dlogits.view(ctx.shape)
        return (d if ctx.unit_grad else d * g), None, None, None


def packed_cross_entropy(scores, captions, decode_lengths, n_valid_dev=None, unit_grad=False):
    """scores (B, T, V) logits; captions (B, Tc) token ids aligned with `scores` rows; targets are captions[:, 1:]
    restricted to the first decode_lengths[b] positions of row b (the reference's pack_padded_sequence + CE).
    n_valid_dev / unit_grad: see _PackedCE.forward."""
    B, T, V = scores.shape
    dev = scores.device
    dl = device_twin(decode_lengths, dev)
    tgt = captions[:, 1:T + 1].to(torch.long)
    if tgt.shape[1] < T:                                   # Transformer: T = 52 positions, targets exist for 51
        tgt = torch.nn.functional.pad(tgt, (0, T - tgt.shape[1]), value=0)
    valid = torch.arange(T, device=dev).unsqueeze(0) < dl.unsqueeze(1)
    targets = torch.where(valid, tgt, torch.full_like(tgt, -1)).contiguous().view(-1)
    return _PackedCE.apply(scores, targets, float(sum(decode_lengths)) if n_valid_dev is None else n_valid_dev,
                           unit_grad)


def packed_targets(captions, decode_lengths, T):
    """int64 (B*T,) targets for ``ccx_softmax_ce``: captions[b, 1+t] for t < decode_lengths[b], else -1."""
    dev = captions.device
    dl = device_twin(decode_lengths, dev)
    tgt = captions[:, 1:T + 1].to(torch.long)
    if tgt.shape[1] < T:
        tgt = torch.nn.functional.pad(tgt, (0, T - tgt.shape[1]), value=0)
    valid = torch.arange(T, device=dev).unsqueeze(0) < dl.unsqueeze(1)
    return torch.where(valid, tgt, torch.full_like(tgt, -1)).contiguous().view(-1)


@torch.no_grad()
def step_metrics(scores, targets, topk=5, group=None):
    """Loss / token count / top-k hits of one step in ONE kernel pass and ONE 12-byte all-reduce, with no host sync
    until the caller reads the result (the reference: 4 scalar all-reduces + 5 ``.item()`` syncs per step,
    trainMultiGPU.py:96-108,396-403).  scores (B,T,V); targets (B*T,) with -1 = ignored.
    Returns a device tensor [global mean token loss, total tokens, top-k accuracy in percent]."""
    B, T, V = scores.shape
    s2 = scores.detach().contiguous().view(B * T, V)
    stats = torch.zeros(3, dtype=torch.float32, device=scores.device)
    _lib.check(_lib.lib().ccx_softmax_ce(ptr(s2), V, ptr(targets), B * T, V, 1.0, None, None, 0, ptr(stats), topk,
                                         _lib.stream_ptr()), "softmax_ce")
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(stats, group=group)
    return torch.stack([stats[0] / stats[1], stats[1], 100.0 * stats[2] / stats[1]])


@torch.no_grad()
def free_running_targets(sequences, captions, end_token, pad_token):
    """Device-side ``preprocessDecoderOutputForMetrics`` (utils/utils.py:261-295): -> (targets (B*T,), decode_len (B,))."""
    B, T = sequences.shape
    caps = captions.contiguous()
    targets = torch.empty(B * T, dtype=torch.long, device=sequences.device)
    dlen = torch.empty(B, dtype=torch.int32, device=sequences.device)
    _lib.check(_lib.lib().ccx_free_running_targets(ptr(sequences.contiguous()), ptr(caps), caps.stride(0), ptr(targets),
                                                   ptr(dlen), B, T, caps.shape[1], end_token, pad_token,
                                                   _lib.stream_ptr()), "free_running_targets")
    return targets, dlen


def free_running_cross_entropy(scores, sequences, captions, end_token, pad_token):
    """Loss of trainWithoutTeacherForcing (trainMultiGPU.py:448-450): CrossEntropyLoss over the rows
    preprocessDecoderOutputForMetrics keeps — positions before each row's first generated <end> (inclusive) whose
    ground-truth token is not <pad>.  Returns (loss with autograd, targets (B*T,) with -1 = ignored, token count)."""
    targets, _ = free_running_targets(sequences, captions, end_token, pad_token)
    n_valid = int((targets >= 0).sum())                  # the reference syncs here too (totalValidTokenCount)
    return _PackedCE.apply(scores, targets, float(max(n_valid, 1)), False), targets, n_valid
