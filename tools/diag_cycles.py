"""Which objects of one train step are only freed by the cyclic GC (reference cycles keep activations alive)."""
import collections, gc, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecaptioningconvnext_b200 import DecoderWithAttention, Encoder, TransformerDecoder
from imagecaptioningconvnext_b200.train_step import caption_train_step, make_optimizers
from oracle.decoder_oracle import random_lstm_decoder_state, random_transformer_decoder_state, synthetic_captions
from oracle.encoder_oracle import random_encoder_state
import bench
kind = sys.argv[1] if len(sys.argv) > 1 else "lstm"
dev = torch.device("cuda"); V, B = 9490, 8
enc = Encoder(compute_dtype=torch.bfloat16); enc.load_state_dict(random_encoder_state(0, 1.0)); enc = enc.to(dev).train()
if kind == "lstm":
    enc.fine_tune(True, 7)
    dec = DecoderWithAttention(512, 512, 512, V, dev, compute_dtype=torch.bfloat16); dec.load_state_dict(random_lstm_decoder_state(0, V))
else:
    enc.fine_tune(False)
    dec = TransformerDecoder(512, 512, V, 52, dev, None, None, True, compute_dtype=torch.bfloat16); dec.load_state_dict(random_transformer_decoder_state(0, V))
dec = dec.to(dev).train()
d_opt, e_opt = make_optimizers(enc, dec)
caps, lens = synthetic_captions(B, 7, V)
b = (bench.synthetic_images(B, 1).to(dev), caps.to(dev), lens.to(dev))
for _ in range(3):
    caption_train_step(enc, dec, *b, d_opt, e_opt)
torch.cuda.synchronize(); gc.collect()
gc.disable()
m0 = torch.cuda.memory_allocated()
caption_train_step(enc, dec, *b, d_opt, e_opt)
torch.cuda.synchronize()
m1 = torch.cuda.memory_allocated()
gc.set_debug(gc.DEBUG_SAVEALL)
n = gc.collect()
m2 = torch.cuda.memory_allocated()
print(f"{kind}: allocated before {m0/2**20:.1f} MiB, after step {m1/2**20:.1f} MiB; gc found {n} unreachable objects")
cnt = collections.Counter(type(o).__name__ for o in gc.garbage)
print(cnt.most_common(15))
tb = sum(o.numel() * o.element_size() for o in gc.garbage if isinstance(o, torch.Tensor))
print(f"tensors in cycles: {tb/2**20:.1f} MiB")
for o in gc.garbage:
    if not isinstance(o, (torch.Tensor, dict, list, tuple, int, float, str)):
        print("  ", type(o), str(o)[:100])
