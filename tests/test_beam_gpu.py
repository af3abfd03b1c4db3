"""GPU parity of the batched beam search against the oracle's restatement of caption.py (per image)."""
import os

import pytest
import torch

from conftest import rel_err
from test_decoders_gpu import V, WORDMAP, _lstm, _transformer
from test_oracle_golden import _beam_inputs

pytestmark = pytest.mark.gpu
TAU = 1e-4


def _compare(kind, sd, model, feats, k, fn):
    from oracle import decoder_oracle as do
    trace = []
    best, done = fn(model, feats.cuda(), WORDMAP, beamSize=k, trace=trace, return_all=True)
    agree = 0
    for i in range(feats.shape[0]):
        otrace = []
        obest, odone, oscores = do.beam_search(sd, feats[i:i + 1], kind, k, V - 2, V - 1, V, trace=otrace)
        # per-step contract: top-k scores within tolerance, (prev, word) exact unless a near-tie
        diverged = False
        for s, (ts, tp, tw) in enumerate(otrace):
            kr = int(trace[s][0][i])
            assert kr == len(ts), (i, s)
            gs, gp, gw = trace[s][1][i, :kr].cpu(), trace[s][2][i, :kr].cpu(), trace[s][3][i, :kr].cpu()
            same = torch.equal(gp.long(), tp) and torch.equal(gw.long(), tw)
            if not same:
                gaps = (ts[:-1] - ts[1:]).abs()
                assert float(gaps.min()) <= TAU or s > 0, f"image {i} step {s}: beams differ without a near-tie"
                diverged = True
                break
            assert torch.allclose(gs, ts, rtol=1e-4, atol=1e-4), (i, s)
        if not diverged:
            agree += 1
            assert best[i] == obest
            assert done[i][0] == odone
            assert torch.allclose(torch.tensor(done[i][1]), torch.tensor(oscores), rtol=1e-4, atol=1e-4)
    return agree


def test_lstm_beam_vs_oracle():
    from imagecaptioningconvnext_b200.beam import beam_search_lstm
    from oracle import decoder_oracle as do
    sd = do.random_lstm_decoder_state(0, V, end_bias=0.21)
    feats = do.synthetic_features(4, 300)
    assert _compare("lstm", sd, _lstm(sd, torch.float32), feats, 5, beam_search_lstm) >= 3


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_lstm_beam_attention_maps_vs_oracle(dtype):
    """beam_search_lstm(return_alphas=True) (caption.py:85,122,129,153): the maps of every
    completed caption, followed back through the beam re-orderings, equal the oracle's seqsAlpha bookkeeping."""
    from imagecaptioningconvnext_b200.beam import beam_search_lstm
    from oracle import decoder_oracle as do
    sd = do.random_lstm_decoder_state(0, V, end_bias=0.21)
    feats = do.synthetic_features(4, 300)
    m = _lstm(sd, dtype)
    best, every = beam_search_lstm(m, feats.cuda(), WORDMAP, beamSize=5, return_all=True, return_alphas=True)
    plain = beam_search_lstm(m, feats.cuda(), WORDMAP, beamSize=5)
    checked = longest = 0
    for i in range(feats.shape[0]):
        al = []
        obest, odone, _ = do.beam_search(sd, feats[i:i + 1], "lstm", 5, V - 2, V - 1, V, alphas_out=al)
        assert (best[i] is None) == (obest is None) and (best[i] is None or best[i][0] == plain[i])
        if [s for s, _ in every[i]] != odone:
            assert dtype == torch.bfloat16          # a near-tie resolved differently in bf16
            continue
        for (seq, maps), omaps in zip(every[i], al[1]):
            assert maps.shape == (len(seq), 7, 7) and torch.equal(maps[0], torch.ones(7, 7))
            assert rel_err(maps.view(len(seq), 49), omaps) < (1e-3 if dtype == torch.float32 else 2e-2)
            checked += 1
            longest = max(longest, len(seq))
        if obest is not None:
            assert best[i][0] == obest and rel_err(best[i][1].view(len(obest), 49), al[0]) < 2e-2
    assert checked >= 5 and longest >= 4, (checked, longest)


def test_transformer_beam_vs_oracle():
    from imagecaptioningconvnext_b200.beam import beam_search_transformer
    from oracle import decoder_oracle as do
    sd = do.random_transformer_decoder_state(0, V, end_bias=3.2)
    feats = do.synthetic_features(3, 300)
    assert _compare("transformer", sd, _transformer(sd, torch.float32), feats, 5, beam_search_transformer) >= 2


def test_beam_k1_equals_greedy_tokens():
    """caption.py hard-codes beamSize=1 (caption.py:484): k=1 beam == greedy argmax sequence."""
    from imagecaptioningconvnext_b200.beam import beam_search_lstm
    from oracle import decoder_oracle as do
    sd = do.random_lstm_decoder_state(1, V, end_bias=0.21)
    feats = do.synthetic_features(6, 5)
    m = _lstm(sd, torch.float32)
    _, _, seqs = m(teacherForcing=False, encoder_out=feats.cuda(), wordMap=WORDMAP, maxDecodeLen=51)
    best = beam_search_lstm(m, feats.cuda(), WORDMAP, beamSize=1)
    for i in range(6):
        row = seqs[i].tolist()
        if V - 1 in row:
            assert best[i] == [V - 2] + row[:row.index(V - 1) + 1]
        else:
            assert best[i] is None


def test_full_pipeline_matches_reference_caption_py_golden(golden_dir):
    """image -> Encoder -> beam search k=5 on libccx equals caption.py's output (tests/golden/beam.pt)."""
    from imagecaptioningconvnext_b200 import Encoder
    from imagecaptioningconvnext_b200.beam import beam_search_lstm, beam_search_transformer
    from oracle import decoder_oracle as do
    from oracle.encoder_oracle import random_encoder_state
    gold = torch.load(os.path.join(golden_dir, "beam.pt"))
    imgs, feats = _beam_inputs(gold)
    enc = Encoder()
    enc.load_state_dict(random_encoder_state(seed=0, layer_scale=1.0))
    enc = enc.cuda().eval()
    with torch.no_grad():
        f = enc(torch.cat(imgs).cuda())
    assert rel_err(f, torch.cat(feats)) < 1e-3
    lsd = do.random_lstm_decoder_state(0, V, end_bias=gold["lstm_end_bias"])
    tsd = do.random_transformer_decoder_state(0, V, end_bias=gold["transformer_end_bias"])
    assert beam_search_lstm(_lstm(lsd, torch.float32), f, WORDMAP, beamSize=gold["k"]) == gold["lstm"]
    assert beam_search_transformer(_transformer(tsd, torch.float32), f, WORDMAP, beamSize=gold["k"]) == gold["transformer"]


@pytest.mark.parametrize("kind", ["lstm", "transformer"])
def test_cuda_graph_captured_beam_search_equals_eager(kind):
    from imagecaptioningconvnext_b200.beam import CapturedBeamSearch, beam_search_lstm, beam_search_transformer
    from oracle import decoder_oracle as do
    if kind == "lstm":
        sd = do.random_lstm_decoder_state(0, V, end_bias=0.21)
        m, fn = _lstm(sd, torch.float32), beam_search_lstm
    else:
        sd = do.random_transformer_decoder_state(0, V, end_bias=3.2)
        m, fn = _transformer(sd, torch.float32), beam_search_transformer
    cap = CapturedBeamSearch(m, WORDMAP, kind, beamSize=5)
    for seed in (300, 301):                      # second batch re-uses the captured graph with new features
        feats = do.synthetic_features(4, seed).cuda()
        eager = fn(m, feats, WORDMAP, beamSize=5, return_all=True)
        graphed = cap(feats, return_all=True)
        assert graphed[0] == eager[0]
        for (gs, gsc), (es, esc) in zip(graphed[1], eager[1]):
            assert gs == es and gsc == esc
