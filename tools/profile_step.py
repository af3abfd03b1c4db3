"""One unit of work between cudaProfilerStart/Stop for `ncu --profile-from-start off` (profiles/README.md):
   python tools/profile_step.py lstm_step | transformer_step | beam | encoder
Everything runs through the eager launch path (a CUDA-graph replay hides the kernels from per-kernel filters)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagecaptioningconvnext_b200 import DecoderWithAttention, Encoder, TransformerDecoder  # noqa: E402
from imagecaptioningconvnext_b200.train_step import caption_train_step, make_optimizers  # noqa: E402
from synthetic import (random_encoder_state, random_lstm_decoder_state, random_transformer_decoder_state,  # noqa: E402
                       synthetic_captions, synthetic_images)

V = 9490
WORDMAP = {"<pad>": 0, "<unk>": V - 3, "<start>": V - 2, "<end>": V - 1}


def main():
    mode = sys.argv[1]
    dev = torch.device("cuda")
    bf16 = torch.bfloat16
    esd = random_encoder_state(seed=0, layer_scale=1.0)
    B = 64 if mode == "encoder" else 32
    imgs = synthetic_images(B, 1).to(dev)
    caps, lens = synthetic_captions(B, 7, V)
    caps, lens, lens_host = caps.to(dev), lens.to(dev), lens
    enc = Encoder(compute_dtype=bf16)
    enc.load_state_dict(esd)
    if mode == "lstm_step":
        enc = enc.to(dev).train()
        enc.fine_tune(True, 7)
        dec = DecoderWithAttention(512, 512, 512, V, dev, compute_dtype=bf16)
        dec.load_state_dict(random_lstm_decoder_state(0, V))
        dec = dec.to(dev).train()
        d_opt, e_opt = make_optimizers(enc, dec)
        unit = lambda: caption_train_step(enc, dec, imgs, caps, lens, d_opt, e_opt, caplens_host=lens_host)
    elif mode == "transformer_step":
        enc = enc.to(dev).train()
        enc.fine_tune(False)
        dec = TransformerDecoder(512, 512, V, 52, dev, None, None, True, compute_dtype=bf16)
        dec.load_state_dict(random_transformer_decoder_state(0, V))
        dec = dec.to(dev).train()
        d_opt, _ = make_optimizers(enc, dec)
        unit = lambda: caption_train_step(enc, dec, imgs, caps, lens, d_opt, None, caplens_host=lens_host)
    elif mode == "beam":
        from imagecaptioningconvnext_b200.beam import beam_search_transformer
        enc = enc.to(dev).eval()
        dec = TransformerDecoder(512, 512, V, 52, dev, None, None, True, compute_dtype=bf16)
        dec.load_state_dict(random_transformer_decoder_state(0, V, end_bias=3.2))
        dec = dec.to(dev).eval()
        with torch.no_grad():
            feats = enc(imgs)
        unit = lambda: beam_search_transformer(dec, feats, WORDMAP, beamSize=5, max_decode_len=6)
    else:
        enc = enc.to(dev).eval()

        def unit():
            with torch.no_grad():
                return enc(imgs)
    for _ in range(4):
        unit()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    unit()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled unit done:", mode)


if __name__ == "__main__":
    main()
