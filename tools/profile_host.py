"""Where the HOST time of one LSTM train step goes (cProfile over 30 steps, no syncs inside)."""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
caps, lens = synthetic_captions(B, 7, V)
b = (bench.synthetic_images(B, 1).to(dev), caps.to(dev), lens.to(dev))
for _ in range(10):
    caption_train_step(enc, dec, *b, d_opt, e_opt, caplens_host=lens)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(30):
    caption_train_step(enc, dec, *b, d_opt, e_opt, caplens_host=lens)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/30:.2f} ms/step, wall {1e3*(t2-t0)/30:.2f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    caption_train_step(enc, dec, *b, d_opt, e_opt, caplens_host=lens)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
