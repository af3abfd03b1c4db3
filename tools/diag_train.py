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

# ---- GPU-time per phase (CUDA events at phase boundaries, no host sync) vs host enqueue time per phase ----
from imagecaptioningconvnext_b200.losses import packed_cross_entropy
def one(events, host):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    t = [time.perf_counter()]
    ev[0].record()
    feats = enc(imgs); ev[1].record(); t.append(time.perf_counter())
    s, cs, dl, al, _ = dec(teacherForcing=True, encoder_out=feats, encoded_captions=caps, caption_lengths=lens)
    loss = packed_cross_entropy(s, cs, dl) + ((1.0 - al.sum(dim=1)) ** 2).mean(); ev[2].record(); t.append(time.perf_counter())
    e_opt.zero_grad(set_to_none=False); d_opt.zero_grad(set_to_none=False); loss.backward(); ev[3].record(); t.append(time.perf_counter())
    e_opt.step(); d_opt.step(); ev[4].record(); t.append(time.perf_counter())
    events.append(ev); host.append([t[i + 1] - t[i] for i in range(4)])
events, host = [], []
for _ in range(10): one(events, host)
torch.cuda.synchronize()
names = ["encoder fwd", "decoder fwd+loss", "backward", "optimizer"]
for i, n in enumerate(names):
    g = sum(ev[i].elapsed_time(ev[i + 1]) for ev in events[3:]) / len(events[3:])
    h = sum(hh[i] for hh in host[3:]) / len(host[3:]) * 1e3
    print(f"{n:18s} gpu-span {g:6.2f} ms   host-enqueue {h:6.2f} ms")
