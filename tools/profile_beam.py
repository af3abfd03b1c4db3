"""Per-kernel-kind time of one batched beam search (library event hooks): python tools/profile_beam.py"""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecaptioningconvnext_b200 import TransformerDecoder, _lib
from imagecaptioningconvnext_b200.beam import beam_search_transformer
from oracle.decoder_oracle import random_transformer_decoder_state, synthetic_features
V = 9490
WORDMAP = {"<pad>": 0, "<unk>": V - 3, "<start>": V - 2, "<end>": V - 1}
dev = torch.device("cuda")
tr = TransformerDecoder(512, 512, V, 52, dev, None, None, True, compute_dtype=torch.bfloat16)
tr.load_state_dict(random_transformer_decoder_state(0, V, end_bias=3.2)); tr = tr.to(dev).eval()
feats = synthetic_features(128, 3).to(dev)
run = lambda: beam_search_transformer(tr, feats, WORDMAP, beamSize=5, _state_only=True)
run(); torch.cuda.synchronize()
t0 = time.perf_counter(); run(); torch.cuda.synchronize(); wall = time.perf_counter() - t0
_lib.prof_begin(); run(); spans = _lib.prof_spans(); prof = _lib.prof_end()
print(f"wall {wall*1e3:.1f} ms; kernel time {sum(v['ms'] for v in prof.values()):.1f} ms in {sum(v['launches'] for v in prof.values())} launches")
for k, v in prof.items():
    if v["launches"]: print(f"  {k:12s} {v['ms']:8.2f} ms {v['launches']:6d} launches  avg {v['ms']*1e3/v['launches']:.1f} us")
# break 'elementwise' and 'gemm' down by work size
agg = collections.defaultdict(lambda: [0, 0.0])
for k, ms, w in spans:
    key = (k, f"{w:.3g}")
    agg[key][0] += 1; agg[key][1] += ms
for (k, w), (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"   {k:12s} work={w:>10s} n={n:5d} total {ms:7.2f} ms avg {ms*1e3/n:6.1f} us")
