"""CPU oracle for the step metrics.  Test infrastructure only.

Restates utils/utils.py:239-254 (``accuracy``), utils/utils.py:261-295 (``preprocessDecoderOutputForMetrics``) and the
loss/token bookkeeping of trainMultiGPU.py:96-108,396-403 for a single rank."""
import torch
import torch.nn.functional as F


def accuracy_counts(scores, targets, k):
    """utils/utils.py:239-254 with gpu='multi': (number of rows whose target is in the top-k, number of rows)."""
    _, ind = scores.topk(k, 1, True, True)
    correct = ind.eq(targets.view(-1, 1).expand_as(ind))
    return float(correct.view(-1).float().sum()), targets.size(0)


def preprocess_decoder_output_for_metrics(predictions, sequences, caps, end_tok, pad_tok, max_len):
    """utils/utils.py:261-295, statement by statement."""
    logits, tgts, total, lens = [], [], 0, []
    for i in range(predictions.size(0)):
        if (sequences[i] == end_tok).any():
            L = int((sequences[i] == end_tok).nonzero(as_tuple=True)[0][0]) + 1
        else:
            L = max_len
        lens.append(L)
        p = predictions[i, :L, :]
        g = caps[i, 1:1 + L]
        m = g != pad_tok
        p, g = p[m], g[m]
        if g.numel() == 0:
            continue
        logits.append(p)
        tgts.append(g)
        total += g.numel()
    return torch.cat(logits, 0), torch.cat(tgts, 0), total, lens


def step_metrics(scores_packed, targets_packed, k=5):
    loss = F.cross_entropy(scores_packed, targets_packed)
    c, n = accuracy_counts(scores_packed, targets_packed, k)
    return float(loss), n, 100.0 * c / n
