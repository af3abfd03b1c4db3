"""GPU parity of the fused step metrics against the oracle restatement of utils/utils.py."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
V = 9490


def test_teacher_forced_step_metrics():
    from imagecaptioningconvnext_b200.losses import packed_targets, step_metrics
    from oracle import decoder_oracle as do
    from oracle import metrics_oracle as mo
    B, T = 6, 52
    caps, lens = do.synthetic_captions(B, 3, V)
    dl = (lens.squeeze(1) - 1).tolist()
    g = torch.Generator().manual_seed(0)
    scores = torch.randn(B, T, V, generator=g)
    for b in range(B):                                   # make some targets land in / out of the top-5
        for t in range(dl[b]):
            if (b + t) % 3 == 0:
                scores[b, t, caps[b, t + 1]] += 3.0
    packed_s = torch.cat([scores[b, :dl[b]] for b in range(B)])
    packed_t = torch.cat([caps[b, 1:1 + dl[b]] for b in range(B)])
    ref = mo.step_metrics(packed_s, packed_t, 5)
    got = step_metrics(scores.cuda(), packed_targets(caps.cuda(), dl, T), 5).cpu()
    assert abs(float(got[0]) - ref[0]) < 1e-4 and int(got[1]) == ref[1] and abs(float(got[2]) - ref[2]) < 1e-4


def test_free_running_metrics_match_preprocessDecoderOutputForMetrics():
    from imagecaptioningconvnext_b200.losses import free_running_targets, step_metrics
    from oracle import decoder_oracle as do
    from oracle import metrics_oracle as mo
    B, T = 7, 51
    g = torch.Generator().manual_seed(1)
    caps, _ = do.synthetic_captions(B, 5, V)
    seqs = torch.randint(1, V - 4, (B, T), generator=g)
    for b, e in enumerate([0, 3, 50, None, 17, None, 9]):        # first <end> positions (None = never ends)
        if e is not None:
            seqs[b, e] = V - 1
            seqs[b, min(e + 5, T - 1)] = V - 1                    # a later <end> must be ignored
    preds = torch.randn(B, T, V, generator=g)
    rp, rt, total, lens = mo.preprocess_decoder_output_for_metrics(preds, seqs, caps, V - 1, 0, T)
    ref = mo.step_metrics(rp, rt, 5)
    targets, dlen = free_running_targets(seqs.cuda(), caps.cuda(), V - 1, 0)
    assert dlen.cpu().tolist() == lens
    assert int((targets >= 0).sum()) == total
    got = step_metrics(preds.cuda(), targets, 5).cpu()
    assert abs(float(got[0]) - ref[0]) < 1e-4 and int(got[1]) == total and abs(float(got[2]) - ref[2]) < 1e-4
