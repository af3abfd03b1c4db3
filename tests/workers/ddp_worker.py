"""torchrun worker: one DDP train step on per-rank batches must produce, on every rank, the MEAN over ranks of the
gradients each rank computes alone (DistributedDataParallel semantics the reference relies on,
trainMultiGPU.py:233-235,384), with the explicit libccx backward underneath."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    from torch.nn.parallel import DistributedDataParallel as DDP
    from imagecaptioningconvnext_b200 import DecoderWithAttention, Encoder, TransformerDecoder
    from imagecaptioningconvnext_b200.losses import packed_cross_entropy
    from oracle import decoder_oracle as do
    from oracle.encoder_oracle import random_encoder_state

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    V = 9490
    kind = sys.argv[1] if len(sys.argv) > 1 else "lstm"
    B = 4
    imgs = torch.randn(B, 3, 64, 64, generator=torch.Generator().manual_seed(10 + rank)).to(dev)
    caps, lens = do.synthetic_captions(B, 20 + rank, V)
    caps, lens = caps.to(dev), lens.to(dev)

    def build():
        enc = Encoder()
        enc.load_state_dict(random_encoder_state(seed=0, layer_scale=1.0))
        enc = enc.to(dev).eval()
        enc.fine_tune(True, 7)
        if kind == "lstm":
            dec = DecoderWithAttention(512, 512, 512, V, dev)
            dec.load_state_dict(do.random_lstm_decoder_state(1, V))
        else:
            dec = TransformerDecoder(512, 512, V, 52, dev, None, None, True)
            dec.load_state_dict(do.random_transformer_decoder_state(1, V))
        dec = dec.to(dev).train()
        dec.dropout_p = 0.0
        return enc, dec

    def step(enc, dec):
        feats = enc(imgs)
        if kind == "lstm":
            s, cs, dl, al, _ = dec(teacherForcing=True, encoder_out=feats, encoded_captions=caps, caption_lengths=lens)
            loss = packed_cross_entropy(s, cs, dl) + ((1.0 - al.sum(dim=1)) ** 2).mean()
        else:
            s, co, dl = dec(teacherForcing=True, encoder_out=feats, encoded_captions=caps, caption_lengths=lens,
                            tgt_key_padding_mask=(caps == 0))
            loss = packed_cross_entropy(s, co, dl)
        loss.backward()
        return loss

    # local gradients without DDP, averaged by hand
    enc, dec = build()
    step(enc, dec)
    manual = {}
    for name, m in (("enc", enc), ("dec", dec)):
        for n, p in m.named_parameters():
            if p.grad is not None:
                g = p.grad.clone()
                dist.all_reduce(g)
                manual[f"{name}.{n}"] = g / world
    # the same step through DDP
    enc2, dec2 = build()
    enc_w, dec_w = DDP(enc2, device_ids=[local]), DDP(dec2, device_ids=[local])
    step(enc_w, dec_w)
    worst = 0.0
    for name, m in (("enc", enc2), ("dec", dec2)):
        for n, p in m.named_parameters():
            if p.requires_grad:
                assert p.grad is not None, f"{name}.{n} received no gradient under DDP"
                ref = manual[f"{name}.{n}"]
                err = float((p.grad - ref).abs().max() / ref.abs().max().clamp_min(1e-20))
                worst = max(worst, err)
    t = torch.tensor([worst], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"DDP_OK kind={kind} world={world} worst_rel_err={float(t):.3e}")
    assert float(t) < 1e-4, float(t)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
