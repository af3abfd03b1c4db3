"""torchrun worker (one rank per GPU).  Multi-GPU gradient semantics of the train step, checked against the ORACLE:

  every rank runs the reference's step body restated in stock torch ops (oracle/, CPU, fp32) on its OWN batch under
  autograd; the mean of those gradients over the ranks is what DistributedDataParallel's averaged all-reduce must
  deliver on every rank (trainMultiGPU.py:233-235,384) — "N-rank gradients == gradients of the big batch" with the
  per-rank loss normalisation the reference has (SURVEY.md §4 iv).

modes:  lstm | transformer   the libccx modules wrapped in torch DDP (fp32 / 3xTF32 compute), one backward
        captured             CapturedTrainStep (CUDA-graph replay, own flat NCCL buckets, bf16 compute) with lr = 0, so
                             that after warm-up, capture and replays the buckets still hold gradients at the initial
                             weights; plus: every rank ends with bit-identical weights when lr > 0
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
V = 9490


def oracle_mean_grads(kind, esd, dsd, imgs, caps, lens, dev, world):
    """Gradients of this rank's loss from the oracle (CPU autograd), averaged over the ranks."""
    from oracle import decoder_oracle as do
    from oracle import encoder_oracle as eo
    e_leaf = {k: v.clone().requires_grad_(k.startswith("convnext.7.")) for k, v in esd.items()}
    d_leaf = {k: v.clone().requires_grad_(v.is_floating_point() and k != "pos_encoding.pe") for k, v in dsd.items()}
    feats = eo.encoder_forward(e_leaf, imgs.cpu(), 7)
    if kind == "lstm":
        p, cs, dl, al, _ = do.lstm_teacher_forcing(d_leaf, feats, caps.cpu(), lens.cpu())
        loss = do.train_loss_lstm(p, cs, dl, al)
    else:
        p, cs, dl = do.transformer_teacher_forcing(d_leaf, feats, caps.cpu(), lens.cpu(), caps.cpu() == 0)
        loss = do.train_loss_transformer(p, cs, dl)
    loss.backward()
    out = {}
    for prefix, leaf in (("enc.", e_leaf), ("dec.", d_leaf)):
        for k, v in leaf.items():
            if v.requires_grad:
                g = (v.grad if v.grad is not None else torch.zeros_like(v)).to(dev)
                dist.all_reduce(g)
                out[prefix + k] = g / world
    return out, float(loss)


def compare(named, ref, rel_tol=None, cos_tol=None):
    """Per tensor: Frobenius-relative error and cosine against the oracle mean gradient.  (Not the max-norm: one ReLU
    unit whose pre-activation is within rounding of zero legitimately changes single entries by a visible amount.)"""
    worst_rel, worst_cos = 0.0, 1.0
    for name, g in named:
        r = ref[name]
        assert g is not None, f"{name} received no gradient"
        if name.endswith("attention.full_att.bias"):
            assert float(g.abs().max()) == 0.0          # softmax is shift invariant: identically zero
            continue
        rel = float((g - r).double().norm() / r.double().norm().clamp_min(1e-300))
        cos = float((g.double().flatten() @ r.double().flatten()) / (g.double().norm() * r.double().norm()).clamp_min(1e-300))
        worst_rel, worst_cos = max(worst_rel, rel), min(worst_cos, cos)
        if rel_tol is not None:
            assert rel < rel_tol, (name, rel)
        if cos_tol is not None:
            assert cos > cos_tol, (name, cos)
    return worst_rel, worst_cos


def main():
    from torch.nn.parallel import DistributedDataParallel as DDP
    from imagecaptioningconvnext_b200 import DecoderWithAttention, Encoder, TransformerDecoder
    from imagecaptioningconvnext_b200.losses import packed_cross_entropy
    from imagecaptioningconvnext_b200.train_step import CapturedTrainStep, make_optimizers
    from synthetic import (random_encoder_state, random_lstm_decoder_state, random_transformer_decoder_state,
                           synthetic_captions)

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    mode = sys.argv[1] if len(sys.argv) > 1 else "lstm"
    kind = "lstm" if mode == "captured" else mode
    cd = torch.bfloat16 if mode == "captured" else torch.float32
    B = 4
    imgs = torch.randn(B, 3, 64, 64, generator=torch.Generator().manual_seed(10 + rank)).to(dev)
    caps, lens = synthetic_captions(B, 20 + rank, V)
    caps, lens = caps.to(dev), lens.to(dev)
    esd = random_encoder_state(seed=0, layer_scale=1.0)
    dsd = random_lstm_decoder_state(1, V) if kind == "lstm" else random_transformer_decoder_state(1, V)
    ref, ref_loss = oracle_mean_grads(kind, esd, dsd, imgs, caps, lens, dev, world)

    def build():
        enc = Encoder(compute_dtype=cd)
        enc.load_state_dict(esd)
        enc = enc.to(dev).eval()                          # eval: no stochastic depth
        enc.fine_tune(True, 7)
        if kind == "lstm":
            dec = DecoderWithAttention(512, 512, 512, V, dev, compute_dtype=cd)
        else:
            dec = TransformerDecoder(512, 512, V, 52, dev, None, None, True, compute_dtype=cd)
        dec.load_state_dict(dsd)
        dec = dec.to(dev).train()
        dec.dropout_p = 0.0
        for m in dec.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        return enc, dec

    def named_grads(enc, dec):
        return [("enc." + n, p.grad) for n, p in enc.named_parameters() if p.requires_grad] + \
               [("dec." + n, p.grad) for n, p in dec.named_parameters() if p.requires_grad]

    if mode != "captured":
        enc, dec = build()
        enc_w, dec_w = DDP(enc, device_ids=[local]), DDP(dec, device_ids=[local])
        feats = enc_w(imgs)
        if kind == "lstm":
            s, cs, dl, al, _ = dec_w(teacherForcing=True, encoder_out=feats, encoded_captions=caps,
                                     caption_lengths=lens)
            loss = packed_cross_entropy(s, cs, dl) + ((1.0 - al.sum(dim=1)) ** 2).mean()
        else:
            s, co, dl = dec_w(teacherForcing=True, encoder_out=feats, encoded_captions=caps, caption_lengths=lens,
                              tgt_key_padding_mask=(caps == 0))
            loss = packed_cross_entropy(s, co, dl)
        loss.backward()
        assert abs(float(loss) - ref_loss) < 1e-3 * abs(ref_loss), (float(loss), ref_loss)
        worst, cos = compare(named_grads(enc, dec), ref, rel_tol=5e-3, cos_tol=0.9999)
    else:
        # (1) lr = 0: gradients in the buckets after warm-up + capture + replays vs the oracle mean gradients
        enc, dec = build()
        d_opt, e_opt = make_optimizers(enc, dec, decoder_lr=0.0, encoder_lr=0.0)
        step = CapturedTrainStep(enc, dec, d_opt, e_opt, warmup_steps=2)
        for _ in range(5):
            loss = step(imgs, caps, lens)
        assert step.graph is not None
        assert abs(float(loss) - ref_loss) < 2e-2 * abs(ref_loss), (float(loss), ref_loss)
        worst, cos = compare(named_grads(enc, dec), ref, cos_tol=0.99)
        # (2) lr > 0: the ranks see different batches, the averaged buckets keep their weights identical
        enc, dec = build()
        d_opt, e_opt = make_optimizers(enc, dec, decoder_lr=1e-3, encoder_lr=1e-4)
        step = CapturedTrainStep(enc, dec, d_opt, e_opt, warmup_steps=2)
        for _ in range(6):
            step(imgs, caps, lens)
        moved = 0.0
        for n, p in list(dec.named_parameters()) + list(enc.convnext[7].named_parameters()):
            mine = p.detach().clone()
            lo, hi = mine.clone(), mine.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert torch.equal(lo, hi), f"{n}: weights differ between ranks"
        moved = float((dec.fc.weight.detach().cpu() - dsd["fc.weight"]).abs().max())
        assert moved > 0, "the optimizer did not move the weights"
    t = torch.tensor([worst, 1.0 - cos], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"DDP_OK mode={mode} world={world} worst_rel_err_vs_oracle={float(t[0]):.3e} "
              f"worst_1_minus_cos={float(t[1]):.3e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
