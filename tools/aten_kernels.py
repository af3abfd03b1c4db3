"""Which ATen kernels does one CapturedTrainStep body launch, and from which source lines?  (torch.profiler with
stacks over one eager execution of the step body — the same launches the CUDA graph replays.)"""
import collections, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagecaptioningconvnext_b200 import DecoderWithAttention, Encoder  # noqa: E402
from imagecaptioningconvnext_b200.train_step import CapturedTrainStep, make_optimizers  # noqa: E402
from synthetic import random_encoder_state, random_lstm_decoder_state, synthetic_captions, synthetic_images  # noqa: E402
V = 9490
dev = torch.device("cuda")
enc = Encoder(compute_dtype=torch.bfloat16); enc.load_state_dict(random_encoder_state(seed=0, layer_scale=1.0))
enc = enc.to(dev).train(); enc.fine_tune(True, 7)
dec = DecoderWithAttention(512, 512, 512, V, dev, compute_dtype=torch.bfloat16)
dec.load_state_dict(random_lstm_decoder_state(0, V)); dec = dec.to(dev).train()
d_opt, e_opt = make_optimizers(enc, dec)
step = CapturedTrainStep(enc, dec, d_opt, e_opt)
imgs = synthetic_images(32, 1).to(dev); caps, lens = synthetic_captions(32, 7, V); caps, lens = caps.to(dev), lens.to(dev)
for _ in range(2): step.eager_step(imgs, caps, lens)
torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA],
                            with_stack=True) as prof:
    step.eager_step(imgs, caps, lens)
    torch.cuda.synchronize()
agg = collections.Counter(); dur = collections.Counter()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for evt in prof.events():
    if evt.device_type != torch.autograd.DeviceType.CPU or not evt.kernels:
        continue
    for k in evt.kernels:
        if "ccx::" in k.name or "lp::" in k.name:
            continue
        where = next((f.strip() for f in (evt.stack or []) if root in f and "tools/" not in f), "(autograd / no frame)")
        key = (evt.name, k.name[:70], where.replace(root + "/", "")[:90])
        agg[key] += 1; dur[key] += k.duration
tot = sum(agg.values())
print(f"{tot} non-ccx kernel launches, {sum(dur.values()):.0f} us")
for key, n in sorted(agg.items(), key=lambda kv: -dur[kv[0]]):
    print(f"{n:3d} x {dur[key]:7.1f} us  {key[0]:28s} {key[1]:70s} {key[2]}")
