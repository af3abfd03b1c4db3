"""Times DecoderWithAttention teacher forcing (forward, and forward+backward) with the persistent recurrence kernels
against the per-step launch loop, B=32, bf16.  Run on the GPU box: python tools/bench_lstm_persist.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imagecaptioningconvnext_b200 import DecoderWithAttention  # noqa: E402
from imagecaptioningconvnext_b200 import _lib  # noqa: E402

V = 9490


def synth(B, seed):
    g = torch.Generator().manual_seed(seed)
    enc = torch.randn(B, 7, 7, 1024, generator=g) * 0.7
    lens = torch.randint(7, 53, (B, 1), generator=g)
    caps = torch.zeros(B, 52, dtype=torch.long)
    for b in range(B):
        L = int(lens[b])
        caps[b, 0] = V - 2
        caps[b, 1:L - 1] = torch.randint(1, V - 3, (L - 2,), generator=g)
        caps[b, L - 1] = V - 1
    return enc.cuda(), caps.cuda(), lens.cuda()


def timed(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, host


def show(d, names, tag):
    T = d.shape[1]
    for r_, (nm, pts) in names.items():
        rows = d[r_, 5:min(T, 30)]
        rows = rows[rows[:, 0] > 0]
        if len(rows) == 0:
            continue
        rel = (rows[:, 1:len(pts)] - rows[:, :len(pts) - 1]).mean(dim=0)
        period = (rows[1:, 0] - rows[:-1, 0]).abs().mean() if len(rows) > 1 else float("nan")
        print(f"   {tag} {nm}: step period {float(period):.2f} us; " +
              ", ".join(f"{pts[i]}->{pts[i + 1]} {float(rel[i]):.2f}" for i in range(len(pts) - 1)))


def main():
    B = 32
    torch.manual_seed(0)
    m = DecoderWithAttention(512, 512, 512, V, torch.device("cuda"), compute_dtype=torch.bfloat16).cuda().train()
    enc, caps, lens = synth(B, 1)
    from imagecaptioningconvnext_b200._host import stash_host_copy
    lens_host = lens.cpu()
    for persist in (False, True):
        m.use_persist = persist

        def fwd():
            stash_host_copy(lens, lens_host)
            with torch.no_grad():
                m._tf_forward(enc, caps, lens)

        def fwdbwd():
            stash_host_copy(lens, lens_host)
            for p in m.parameters():
                p.grad = None
            out = m(teacherForcing=True, encoder_out=enc, encoded_captions=caps, caption_lengths=lens)
            (out[0].float().mean() + out[3].mean()).backward()

        g, h = timed(fwd)
        print(f"persist={persist}: TF forward   {g:.3f} ms gpu, {h:.3f} ms host enqueue")
        g, h = timed(fwdbwd)
        print(f"persist={persist}: TF fwd+bwd   {g:.3f} ms gpu, {h:.3f} ms host enqueue")
        _lib.prof_begin()
        fwdbwd()
        r = _lib.prof_end()
        print("   per-kind:", {k: (round(v["ms"], 3), v["launches"]) for k, v in r.items() if v["launches"]})
        if persist:
            m._persist_dbg = True
            fwd()
            torch.cuda.synchronize()
            d = m._persist_dbg.cpu().double() / 1.9e3          # cycles -> us at ~1.9 GHz
            m._persist_dbg = None
            T = d.shape[1]
            names = {0: ("G2", ["top", "h part done", "awe seen", "own mma done", "all mma done", "signalled"]),
                     1: ("G1", ["top", "h seen", "mma committed", "mma done", "signalled"]),
                     2: ("ATT", ["top", "hg seen", "scores", "softmax", "awe+store", "signalled"])}
            show(d, names, "fwd")
            m._persist_dbg = True
            stash_host_copy(lens, lens_host)
            out = m(teacherForcing=True, encoder_out=enc, encoded_captions=caps, caption_lengths=lens)
            m._persist_dbg = True
            (out[0].float().mean() + out[3].mean()).backward()
            torch.cuda.synchronize()
            d = m._persist_dbg.cpu().double() / 1.9e3
            m._persist_dbg = None
            names = {0: ("HP", ["top", "dhg seen", "own mma done", "all mma done", "signalled"]),
                     1: ("X", ["top", "dg seen", "mma committed", "signalled"]),
                     2: ("ATT", ["top", "x seen", "dgp", "dalpha", "datt2", "signalled"])}
            show(d, names, "bwd")


if __name__ == "__main__":
    main()
