"""One teacher-forced train step — the body of the reference's hot loop (trainMultiGPU.py:357-394 = train.py:257-291):
encoder -> decoder(teacherForcing=True) -> packed cross-entropy (+ doubly-stochastic attention regulariser for the
LSTM decoder) -> zero_grad -> backward -> clip_gradient(+-5) -> Adam.  Everything heavy runs on libccx; wrap the
modules in ``torch.nn.parallel.DistributedDataParallel`` (as the reference does, trainMultiGPU.py:233-235) and the
gradient all-reduce over NCCL overlaps with the explicit backward.
"""

from ._host import stash_host_copy
from .decoder import DecoderWithAttention
from .losses import free_running_cross_entropy, packed_cross_entropy
from .optim import ClampAdam


def _unwrap(m):
    return m.module if hasattr(m, "module") else m


def make_optimizers(encoder, decoder, decoder_lr=1e-4, encoder_lr=1e-4, grad_clip=5.0):
    """trainMultiGPU.py:191-198,254: Adam over the trainable parameters; clip_gradient fused into the update."""
    dec_opt = ClampAdam([p for p in decoder.parameters() if p.requires_grad], lr=decoder_lr, grad_clip=grad_clip)
    enc_params = [p for p in encoder.parameters() if p.requires_grad]
    enc_opt = ClampAdam(enc_params, lr=encoder_lr, grad_clip=grad_clip) if enc_params else None
    return dec_opt, enc_opt


def caption_train_step(encoder, decoder, imgs, caps, caplens, decoder_optimizer, encoder_optimizer=None, pad_token=0,
                       alpha_c=1.0, teacher_forcing=True, wordMap=None, max_decode_len=51, caplens_host=None):
    """Returns the loss tensor (no host sync).  imgs (B,3,256,256) fp32, caps (B,52) int64, caplens (B,1) int64.
    teacher_forcing=False: the free-running step of trainWithoutTeacherForcing (trainMultiGPU.py:423-460; needs
    wordMap for <start>/<end>).  caplens_host: optional host copy of caplens (what the DataLoader yielded before
    `.to(device)`, trainMultiGPU.py:355-359) — saves the step's only device-to-host read."""
    if teacher_forcing:
        # the decoders need the lengths on the host: read them before the encoder is queued — or not at all when the
        # caller still has the data loader's host tensor (caplens_host), which leaves the step free of host syncs
        stash_host_copy(caplens, caplens_host)
    feats = encoder(imgs)                                                              # trainMultiGPU.py:361
    if not teacher_forcing:
        out = decoder(teacherForcing=False, encoder_out=feats, wordMap=wordMap, maxDecodeLen=max_decode_len)  # :445,:451
        scores, sequences = out[0], out[-1]
        loss, _, _ = free_running_cross_entropy(scores, sequences, caps, wordMap['<end>'], pad_token)   # :446-450
        if isinstance(_unwrap(decoder), DecoderWithAttention):
            loss = loss + alpha_c * ((1.0 - out[1].sum(dim=1)) ** 2).mean()           # :448
    elif isinstance(_unwrap(decoder), DecoderWithAttention):
        scores, caps_sorted, decode_lengths, alphas, _ = decoder(teacherForcing=True, encoder_out=feats,
                                                                 encoded_captions=caps, caption_lengths=caplens)
        loss = packed_cross_entropy(scores, caps_sorted, decode_lengths)              # :364-367
        loss = loss + alpha_c * ((1.0 - alphas.sum(dim=1)) ** 2).mean()               # :369
    else:
        kpm = caps == pad_token                                                        # :371
        scores, caps_out, decode_lengths = decoder(teacherForcing=True, encoder_out=feats, encoded_captions=caps,
                                                   caption_lengths=caplens, tgt_key_padding_mask=kpm)
        loss = packed_cross_entropy(scores, caps_out, decode_lengths)                 # :373-377
    if encoder_optimizer is not None:
        encoder_optimizer.zero_grad(set_to_none=False)
    decoder_optimizer.zero_grad(set_to_none=False)                                     # :381-383
    loss.backward()                                                                    # :384 (DDP all-reduce inside)
    if encoder_optimizer is not None:
        encoder_optimizer.step()                                                       # :387-394 (clamp fused)
    decoder_optimizer.step()
    return loss.detach()
