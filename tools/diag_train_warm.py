"""Host enqueue time vs device time of the LSTM train step over 300 steps from a cold start."""
import os, sys, time
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
bufs = []
for i in range(4):
    caps, lens = synthetic_captions(B, 7 + i, V)
    bufs.append((bench.synthetic_images(B, 1234 + i).to(dev), caps.to(dev), lens.to(dev)))
nsame = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.cuda.synchronize()
for blk in range(12):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host = 0.0
    torch.cuda.synchronize(); w0 = time.perf_counter(); e0.record()
    for i in range(25):
        t0 = time.perf_counter()
        b = bufs[i % nsame]
        caption_train_step(enc, dec, b[0], b[1], b[2], d_opt, e_opt)
        host += time.perf_counter() - t0
    e1.record(); torch.cuda.synchronize(); w1 = time.perf_counter()
    print(f"block {blk}: device {e0.elapsed_time(e1)/25:.2f} ms/step  host-enqueue {host/25*1e3:.2f} ms/step  wall {(w1-w0)/25*1e3:.2f}  "
          f"alloc {torch.cuda.memory_allocated()/2**20:.0f} MiB reserved {torch.cuda.memory_reserved()/2**20:.0f} MiB "
          f"retries {torch.cuda.memory_stats()['num_alloc_retries']} segs {torch.cuda.memory_stats()['segment.all.allocated']}", flush=True)
