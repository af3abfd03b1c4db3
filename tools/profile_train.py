"""Per-kernel-kind time of one train step (library event hooks) vs wall-clock: python tools/profile_train.py [lstm|transformer]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecaptioningconvnext_b200 import DecoderWithAttention, Encoder, TransformerDecoder, _lib
from imagecaptioningconvnext_b200.train_step import caption_train_step, make_optimizers
from oracle.decoder_oracle import random_lstm_decoder_state, random_transformer_decoder_state, synthetic_captions
from oracle.encoder_oracle import random_encoder_state
import bench

kind = sys.argv[1] if len(sys.argv) > 1 else "lstm"
dev = torch.device("cuda")
V, B = 9490, 32
enc = Encoder(compute_dtype=torch.bfloat16); enc.load_state_dict(random_encoder_state(0, 1.0)); enc = enc.to(dev).train()
if kind == "lstm":
    enc.fine_tune(True, 7)
    dec = DecoderWithAttention(512, 512, 512, V, dev, compute_dtype=torch.bfloat16); dec.load_state_dict(random_lstm_decoder_state(0, V))
else:
    enc.fine_tune(False)
    dec = TransformerDecoder(512, 512, V, 52, dev, None, None, True, compute_dtype=torch.bfloat16); dec.load_state_dict(random_transformer_decoder_state(0, V))
dec = dec.to(dev).train()
d_opt, e_opt = make_optimizers(enc, dec)
imgs = bench.synthetic_images(B, 1).to(dev)
caps, lens = synthetic_captions(B, 2, V); caps, lens = caps.to(dev), lens.to(dev)
step = lambda: caption_train_step(enc, dec, imgs, caps, lens, d_opt, e_opt)
for _ in range(3): step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): step()
torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 5
_lib.prof_begin()
for _ in range(5): step()
spans = _lib.prof_spans()
prof = _lib.prof_end()
print(f"{kind}: wall {wall*1e3:.2f} ms/step; libccx kernel time {sum(v['ms'] for v in prof.values())/5:.2f} ms/step in {sum(v['launches'] for v in prof.values())//5} launches")
for k, v in prof.items():
    if v["launches"]: print(f"  {k:12s} {v['ms']/5:8.3f} ms  {v['launches']//5:5d} launches")
big = sorted(spans[:len(spans)//5], key=lambda s: -s[1])[:12]
for k, ms, w in big: print(f"   top: {k:10s} {ms*1e3:8.1f} us work={w:.3g}")
# host-only cost: time the python side with a CPU profiler-free trick: run under cuda graphs impossible; report wall - gpu

# ---- where does the wall time go: host-side phase timing (each phase followed by a sync, so GPU time is included) ----
import contextlib
from imagecaptioningconvnext_b200.losses import packed_cross_entropy
def phase_times():
    t = {}
    def mark(name, t0):
        torch.cuda.synchronize(); t[name] = t.get(name, 0.0) + time.perf_counter() - t0
    t0 = time.perf_counter(); feats = enc(imgs); mark("encoder fwd", t0)
    t0 = time.perf_counter()
    if kind == "lstm":
        s, cs, dl, al, _ = dec(teacherForcing=True, encoder_out=feats, encoded_captions=caps, caption_lengths=lens)
        loss = packed_cross_entropy(s, cs, dl) + ((1.0 - al.sum(dim=1)) ** 2).mean()
    else:
        s, co, dl = dec(teacherForcing=True, encoder_out=feats, encoded_captions=caps, caption_lengths=lens, tgt_key_padding_mask=(caps == 0))
        loss = packed_cross_entropy(s, co, dl)
    mark("decoder fwd + loss", t0)
    t0 = time.perf_counter()
    if e_opt is not None: e_opt.zero_grad(set_to_none=False)
    d_opt.zero_grad(set_to_none=False)
    loss.backward(); mark("backward", t0)
    t0 = time.perf_counter()
    if e_opt is not None: e_opt.step()
    d_opt.step(); mark("optimizer", t0)
    return t
phase_times()
acc = {}
for _ in range(5):
    for k, v in phase_times().items(): acc[k] = acc.get(k, 0.0) + v / 5
print("phase wall times (ms, each synced):", {k: round(v * 1e3, 2) for k, v in acc.items()}, "sum", round(sum(acc.values()) * 1e3, 2))
