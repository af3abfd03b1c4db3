import os, sys, time
sys.path.insert(0, "/root/repo")
import torch
from imagecaptioningconvnext_b200 import DecoderWithAttention, Encoder
from imagecaptioningconvnext_b200.train_step import caption_train_step, make_optimizers
from oracle.decoder_oracle import random_lstm_decoder_state, synthetic_captions
from oracle.encoder_oracle import random_encoder_state
import bench
dev = torch.device("cuda"); V, B = 9490, 32
enc = Encoder(compute_dtype=torch.bfloat16); enc.load_state_dict(random_encoder_state(0, 1.0)); enc = enc.to(dev).train(); enc.fine_tune(True, 7)
dec = DecoderWithAttention(512, 512, 512, V, dev, compute_dtype=torch.bfloat16); dec.load_state_dict(random_lstm_decoder_state(0, V)); dec = dec.to(dev).train()
d_opt, e_opt = make_optimizers(enc, dec)
imgs = bench.synthetic_images(B, 1).to(dev); caps, lens = synthetic_captions(B, 2, V); caps, lens = caps.to(dev), lens.to(dev)
ts = []
for i in range(16):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    caption_train_step(enc, dec, imgs, caps, lens, d_opt, e_opt)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    ts.append((round((t1 - t0) * 1e3, 1), round((t2 - t0) * 1e3, 1)))
print("per step (host ms, total ms):", ts)
print("mem MB", torch.cuda.memory_allocated() // 2**20, torch.cuda.memory_reserved() // 2**20)
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(8): caption_train_step(enc, dec, imgs, caps, lens, d_opt, e_opt)
torch.cuda.synchronize(); print("8 steps unsynced: ms/step", round((time.perf_counter() - t0) / 8 * 1e3, 2))

import gc
gc.collect(); gc.disable()
ts = []
for i in range(16):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    caption_train_step(enc, dec, imgs, caps, lens, d_opt, e_opt)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    ts.append((round((t1 - t0) * 1e3, 1), round((t2 - t0) * 1e3, 1)))
print("gc disabled:", ts)
gc.enable()
print("gc counts", gc.get_count(), "objects", len(gc.get_objects()))
