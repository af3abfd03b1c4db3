"""Per-launch spans of one LSTM train step grouped by (kind, position in step): python tools/profile_lstm_spans.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecaptioningconvnext_b200 import DecoderWithAttention, Encoder, _lib
from imagecaptioningconvnext_b200.train_step import caption_train_step, make_optimizers
from oracle.decoder_oracle import random_lstm_decoder_state, synthetic_captions
from oracle.encoder_oracle import random_encoder_state
import bench
dev = torch.device("cuda")
V, B = 9490, 32
enc = Encoder(compute_dtype=torch.bfloat16); enc.load_state_dict(random_encoder_state(0, 1.0)); enc = enc.to(dev).train()
enc.fine_tune(True, 7)
dec = DecoderWithAttention(512, 512, 512, V, dev, compute_dtype=torch.bfloat16); dec.load_state_dict(random_lstm_decoder_state(0, V))
dec = dec.to(dev).train()
d_opt, e_opt = make_optimizers(enc, dec)
imgs = bench.synthetic_images(B, 1).to(dev)
caps, lens = synthetic_captions(B, 2, V); caps, lens = caps.to(dev), lens.to(dev)
step = lambda: caption_train_step(enc, dec, imgs, caps, lens, d_opt, e_opt)
for _ in range(4): step()
_lib.prof_begin()
step()
spans = _lib.prof_spans()
_lib.prof_end()
t = 0.0
for i, (k, ms, w) in enumerate(spans):
    t += ms
    print(f"{i:4d} {k:12s} {ms*1e3:8.1f} us  work={w:.3g}  cum={t:.3f} ms")
