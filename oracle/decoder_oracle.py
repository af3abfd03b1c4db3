"""CPU oracle for the two caption decoders and the search loops.  Test infrastructure only (see oracle/__init__.py).

Restates, in plain torch tensor ops on a reference-format ``state_dict``:
  * models/decoder.py:25-31        Attention.forward                          -> ``attention``
  * models/decoder.py:63-67        init_hidden_state                          -> ``init_hidden_state``
  * models/decoder.py:69-113       forwardWithTeacherForcing                  -> ``lstm_teacher_forcing``
  * models/decoder.py:119-163      forwardWithoutTeacherForcing (greedy)      -> ``lstm_greedy``
  * caption.py:39-155              caption_image_beam_search (LSTM)           -> ``beam_search`` (step_fn = lstm)
  * models/transformerDecoder.py:14-27,88-108   PositionalEncoding, TF forward -> ``transformer_teacher_forcing``
  * models/transformerDecoder.py:110-160        greedy without KV cache        -> ``transformer_greedy``
  * caption.py:160-255             caption_image_beam_search_transformer      -> ``beam_search`` (step_fn = transformer)
  * trainMultiGPU.py:357-394       loss of the train step (packed CE + alpha regulariser) -> ``train_loss_*``
The third-party arithmetic restated here is torch 2.11's nn.LSTMCell (torch/nn/modules/rnn.py:1755-1778),
nn.TransformerDecoderLayer post-norm forward (torch/nn/modules/transformer.py:1089-1199) and
F.multi_head_attention_forward (packed in-proj, 1/sqrt(hd) scaling, additive -inf masks).
Every function works in the dtype of the tensors it is given (fp32 for parity, fp64 to label near-ties, H12).
"""
import math

import torch
import torch.nn.functional as F


# deterministic random-init weights / synthetic inputs live in the neutral ``synthetic`` module (bench.py's own arm
# uses them too and must not import the oracle); re-exported here for the tests
from synthetic import (random_lstm_decoder_state, random_transformer_decoder_state, synthetic_captions,  # noqa: F401,E402
                       synthetic_features)


# ------------------------------------------------------------------------------------------------
# LSTM + Bahdanau attention decoder
# ------------------------------------------------------------------------------------------------
def attention(sd, enc, h, att1=None):
    """models/decoder.py:25-31.  enc (b,P,E), h (b,D) -> awe (b,E), alpha (b,P)."""
    if att1 is None:
        att1 = F.linear(enc, sd["attention.encoder_att.weight"], sd["attention.encoder_att.bias"])
    att2 = F.linear(h, sd["attention.decoder_att.weight"], sd["attention.decoder_att.bias"])
    att = F.linear(F.relu(att1 + att2.unsqueeze(1)), sd["attention.full_att.weight"],
                   sd["attention.full_att.bias"]).squeeze(2)
    alpha = F.softmax(att, dim=1)
    awe = (enc * alpha.unsqueeze(2)).sum(dim=1)
    return awe, alpha


def init_hidden_state(sd, enc):
    m = enc.mean(dim=1)
    return F.linear(m, sd["init_h.weight"], sd["init_h.bias"]), F.linear(m, sd["init_c.weight"], sd["init_c.bias"])


def lstm_cell(sd, x, h, c):
    """torch nn.LSTMCell: gates = W_ih x + b_ih + W_hh h + b_hh, chunks (i, f, g, o)."""
    g = F.linear(x, sd["decode_step.weight_ih"], sd["decode_step.bias_ih"]) + \
        F.linear(h, sd["decode_step.weight_hh"], sd["decode_step.bias_hh"])
    i, f, gg, o = g.chunk(4, dim=1)
    c2 = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
    h2 = torch.sigmoid(o) * torch.tanh(c2)
    return h2, c2


def lstm_step(sd, enc, emb, h, c):
    """One decode step shared by TF / greedy / beam: models/decoder.py:102-108."""
    awe, alpha = attention(sd, enc, h)
    gate = torch.sigmoid(F.linear(h, sd["f_beta.weight"], sd["f_beta.bias"]))
    h, c = lstm_cell(sd, torch.cat([emb, gate * awe], dim=1), h, c)
    return h, c, alpha


def lstm_teacher_forcing(sd, encoder_out, caps, caplens, dropmask=None):
    """models/decoder.py:69-113.  dropmask: optional (B, max_dl, D) multiplier applied to h before fc (the
    nn.Dropout of :109 with the mask injected, rows in SORTED order); None = eval mode."""
    B, E = encoder_out.size(0), encoder_out.size(-1)
    enc = encoder_out.reshape(B, -1, E)
    lens, sort_ind = caplens.squeeze(1).sort(dim=0, descending=True)
    enc, caps = enc[sort_ind], caps[sort_ind]
    emb = sd["embedding.weight"][caps]
    h, c = init_hidden_state(sd, enc)
    dl = (lens - 1).tolist()
    V = sd["fc.weight"].shape[0]
    preds = torch.zeros(B, max(dl), V, dtype=enc.dtype, device=enc.device)
    alphas = torch.zeros(B, max(dl), enc.size(1), dtype=enc.dtype, device=enc.device)
    for t in range(max(dl)):
        bt = sum(l > t for l in dl)
        h, c, alpha = lstm_step(sd, enc[:bt], emb[:bt, t], h[:bt], c[:bt])
        hd = h if dropmask is None else h * dropmask[:bt, t]
        preds[:bt, t] = F.linear(hd, sd["fc.weight"], sd["fc.bias"])
        alphas[:bt, t] = alpha
    return preds, caps, dl, alphas, sort_ind


def lstm_greedy(sd, encoder_out, start_tok, end_tok, max_len, dropmask=None):
    """models/decoder.py:119-163.  dropmask: optional (B, max_len, D) multiplier applied to h before fc (the
    nn.Dropout of :150 with the mask injected, rows in ORIGINAL order); None = eval mode.  Differentiable: the
    free-running training mode (trainMultiGPU.py:423-498) back-propagates through this loop."""
    B, E = encoder_out.size(0), encoder_out.size(-1)
    enc = encoder_out.reshape(B, -1, E)
    h, c = init_hidden_state(sd, enc)
    V = sd["fc.weight"].shape[0]
    dev = enc.device
    inputs = sd["embedding.weight"][torch.full((B,), start_tok, dtype=torch.long, device=dev)].clone()
    preds = torch.zeros(B, max_len, V, dtype=enc.dtype, device=dev)
    alphas = torch.zeros(B, max_len, enc.size(1), dtype=enc.dtype, device=dev)
    seqs = torch.zeros(B, max_len, dtype=torch.long, device=dev)
    finished = torch.zeros(B, dtype=torch.bool, device=dev)
    for t in range(max_len):
        act = (~finished).nonzero(as_tuple=False).squeeze(1)
        if len(act) == 0:
            break
        hn, cn, alpha = lstm_step(sd, enc[act], inputs[act], h[act], c[act])
        hd = hn if dropmask is None else hn * dropmask[act, t]
        p = F.linear(hd, sd["fc.weight"], sd["fc.bias"])
        preds[act, t] = p
        alphas[act, t] = alpha
        ids = p.argmax(dim=1)
        seqs[act, t] = ids
        finished[act] |= ids == end_tok
        inputs[act] = sd["embedding.weight"][ids]
        h[act], c[act] = hn, cn
    return preds, alphas, seqs


# ------------------------------------------------------------------------------------------------
# Transformer decoder
# ------------------------------------------------------------------------------------------------
def positional_encoding(embed_dim, max_len, dtype=torch.float32):
    """models/transformerDecoder.py:14-27."""
    pe = torch.zeros(max_len, embed_dim)
    pos = torch.arange(0, max_len, dtype=torch.float).unsqueeze(1)
    div = torch.exp(torch.arange(0, embed_dim, 2).float() * (-math.log(10000.0) / embed_dim))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe.to(dtype)


def _mha(x_q, x_kv, w_in, b_in, w_out, b_out, nheads, add_mask=None, prob_mask=None, probs_sink=None):
    """F.multi_head_attention_forward restated, batch-first: x_q (B,Tq,D), x_kv (B,Tk,D);
    add_mask broadcastable to (B,H,Tq,Tk) with 0 / -inf entries; prob_mask = attention-dropout multiplier."""
    B, Tq, D = x_q.shape
    Tk = x_kv.shape[1]
    hd = D // nheads
    q = F.linear(x_q, w_in[:D], b_in[:D])
    k = F.linear(x_kv, w_in[D:2 * D], b_in[D:2 * D])
    v = F.linear(x_kv, w_in[2 * D:], b_in[2 * D:])
    q = q.view(B, Tq, nheads, hd).transpose(1, 2)
    k = k.view(B, Tk, nheads, hd).transpose(1, 2)
    v = v.view(B, Tk, nheads, hd).transpose(1, 2)
    s = (q / math.sqrt(hd)) @ k.transpose(-1, -2)
    if add_mask is not None:
        s = s + add_mask
    p = F.softmax(s, dim=-1)
    if prob_mask is not None:
        p = p * prob_mask
    if probs_sink is not None:
        probs_sink.append(p)          # what need_weights=True, average_attn_weights=False returns: (B, H, Tq, Tk)
    ctx = (p @ v).transpose(1, 2).reshape(B, Tq, D)
    return F.linear(ctx, w_out, b_out)


def transformer_layers(sd, x, mem, nheads, nlayers, self_mask, drop=None, cross_probs=None):
    """nn.TransformerDecoder of post-norm ReLU layers (torch/nn/modules/transformer.py:1089-1199), batch-first.
    drop: optional dict of injected dropout multipliers keyed (layer, name), names: 'sa_p','ca_p' (attention
    probabilities), 'd1','d2','d3' (residual dropouts), 'ff' (FFN hidden)."""
    drop = drop or {}
    for l in range(nlayers):
        p = f"transformer_decoder.layers.{l}."
        sa = _mha(x, x, sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"],
                  sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"], nheads, self_mask,
                  drop.get((l, "sa_p")))
        if (l, "d1") in drop:
            sa = sa * drop[(l, "d1")]
        x = F.layer_norm(x + sa, (x.shape[-1],), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
        ca = _mha(x, mem, sd[p + "multihead_attn.in_proj_weight"], sd[p + "multihead_attn.in_proj_bias"],
                  sd[p + "multihead_attn.out_proj.weight"], sd[p + "multihead_attn.out_proj.bias"], nheads, None,
                  drop.get((l, "ca_p")), probs_sink=cross_probs)
        if (l, "d2") in drop:
            ca = ca * drop[(l, "d2")]
        x = F.layer_norm(x + ca, (x.shape[-1],), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
        hdn = F.relu(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"]))
        if (l, "ff") in drop:
            hdn = hdn * drop[(l, "ff")]
        ff = F.linear(hdn, sd[p + "linear2.weight"], sd[p + "linear2.bias"])
        if (l, "d3") in drop:
            ff = ff * drop[(l, "d3")]
        x = F.layer_norm(x + ff, (x.shape[-1],), sd[p + "norm3.weight"], sd[p + "norm3.bias"], 1e-5)
    return x


def _num_layers(sd):
    return 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer_decoder.layers."))


def transformer_memory(sd, encoder_out):
    B, E = encoder_out.size(0), encoder_out.size(-1)
    enc = encoder_out.reshape(B, -1, E)
    if "encoder_proj.weight" in sd:
        return F.linear(enc, sd["encoder_proj.weight"], sd["encoder_proj.bias"])
    return enc


def transformer_teacher_forcing(sd, encoder_out, caps, caplens, key_padding_mask, nheads=8, drop=None,
                                return_alphas=False):
    """models/transformerDecoder.py:88-108.  key_padding_mask (B,T) bool, True = pad.
    drop: injected dropout multipliers; additionally key 'emb' (B,T,D) for the embedding dropout of :98.
    return_alphas: also return the AttVis variant's fourth output (models/transformerDecoderAttVis.py:163-165)."""
    drop = drop or {}
    dl = (caplens.squeeze(1) - 1).tolist()
    mem = transformer_memory(sd, encoder_out)
    emb = sd["embedding.weight"][caps]
    if "emb" in drop:
        emb = emb * drop["emb"]
    T = caps.shape[1]
    x = emb + sd["pos_encoding.pe"][0, :T].to(emb.dtype)
    causal = torch.full((T, T), float("-inf"), dtype=x.dtype, device=x.device).triu(1)
    mask = causal.view(1, 1, T, T)
    if key_padding_mask is not None:
        kp = torch.zeros(key_padding_mask.shape, dtype=x.dtype, device=x.device).masked_fill(key_padding_mask, float("-inf"))
        mask = mask + kp.view(-1, 1, 1, T)
    cross = [] if return_alphas else None
    y = transformer_layers(sd, x, mem, nheads, _num_layers(sd), mask, drop, cross_probs=cross)
    preds = F.linear(y, sd["fc_out.weight"], sd["fc_out.bias"])
    if return_alphas:
        # (L,B,H,T,P).mean(dim=(0,3)).permute(1,0,2): the reference averages over layers and TARGET POSITIONS and
        # returns (H, B, P) — not the (B, T, P) its comment promises; reproduced as is
        return preds, caps, dl, torch.stack(cross, 0).mean(dim=(0, 3)).permute(1, 0, 2)
    return preds, caps, dl


def transformer_last_logits(sd, mem, tokens, nheads=8, alpha_out=None, drop=None):
    """Re-run the whole prefix (no KV cache, as the reference does) and return fc_out of the last position.
    alpha_out (list): receives the last position's cross-attention map averaged over layers and heads
    (models/transformerDecoderAttVis.py:223-226).  drop: injected dropout multipliers of THIS prefix pass (train mode:
    models/transformerDecoder.py:129-130 applies self.dropout to the re-embedded prefix and the layers run with their
    dropouts live, a fresh realisation at every step), keys as in transformer_teacher_forcing."""
    drop = drop or {}
    T = tokens.shape[1]
    emb = sd["embedding.weight"][tokens]
    if "emb" in drop:
        emb = emb * drop["emb"]
    x = emb + sd["pos_encoding.pe"][0, :T].to(mem.dtype)
    causal = torch.full((T, T), float("-inf"), dtype=x.dtype, device=x.device).triu(1).view(1, 1, T, T)
    cross = [] if alpha_out is not None else None
    y = transformer_layers(sd, x, mem, nheads, _num_layers(sd), causal, drop, cross_probs=cross)
    if alpha_out is not None:
        alpha_out.append(torch.stack(cross, 0)[:, :, :, -1, :].mean(dim=(0, 2)))
    return F.linear(y[:, -1], sd["fc_out.weight"], sd["fc_out.bias"])


def transformer_greedy(sd, encoder_out, start_tok, end_tok, pad_tok, max_len, nheads=8, return_alphas=False,
                       drops=None):
    """models/transformerDecoder.py:110-160; return_alphas: the AttVis variant's third output.
    drops: None = eval mode; else a list over steps t of injected dropout multipliers for step t's prefix pass
    (full-batch shapes with T = t + 1; the rows of the still-active samples are used) — the train-mode behaviour of
    trainWithoutTeacherForcing (trainMultiGPU.py:425-452)."""
    B = encoder_out.size(0)
    mem = transformer_memory(sd, encoder_out)
    V = sd["fc_out.weight"].shape[0]
    dev = mem.device
    inputs = torch.full((B, 1), start_tok, dtype=torch.long, device=dev)
    preds = torch.zeros(B, max_len, V, dtype=mem.dtype, device=dev)
    seqs = torch.zeros(B, max_len, dtype=torch.long, device=dev)
    alphas = torch.zeros(B, max_len, mem.size(1), dtype=mem.dtype, device=dev)
    finished = torch.zeros(B, dtype=torch.bool, device=dev)
    for t in range(max_len):
        act = (~finished).nonzero(as_tuple=False).squeeze(1)
        if len(act) == 0:
            break
        a_out = [] if return_alphas else None
        drop_t = None if drops is None else {k: v[act] for k, v in drops[t].items()}
        p = transformer_last_logits(sd, mem[act], inputs[act], nheads, alpha_out=a_out, drop=drop_t)
        if return_alphas:
            alphas[act, t] = a_out[0]
        preds[act, t] = p
        ids = p.argmax(dim=-1)
        seqs[act, t] = ids
        finished[act] |= ids == end_tok
        new = torch.full((B, t + 2), pad_tok, dtype=torch.long, device=dev)
        new[:, :t + 1] = inputs
        new[act, t + 1] = ids
        inputs = new
    if return_alphas:
        return preds, seqs, alphas
    return preds, seqs


# ------------------------------------------------------------------------------------------------
# beam search (caption.py:39-155 and :160-255) — one image, beams as batch
# ------------------------------------------------------------------------------------------------
def beam_search(sd, encoder_out, kind, k, start_tok, end_tok, vocab, max_steps=50, nheads=8, trace=None,
                alphas_out=None):
    """Returns (best_seq or None if nothing completed, complete_seqs, complete_scores).
    ``trace`` (a list) receives per step: (top-k scores, prev beam indices, next words) — the per-step contract.
    ``alphas_out`` (a list, LSTM only) receives [the best sequence's attention maps (len(seq), P) — the reference's
    second return value (caption.py:85,122,129,153), first entry all ones —, the maps of all completed sequences]."""
    E = encoder_out.size(-1)
    enc1 = encoder_out.reshape(1, -1, E)
    dev = enc1.device          # device-agnostic: bench.py also runs this loop with torch's eager CUDA kernels
    seqs = torch.full((k, 1), start_tok, dtype=torch.long, device=dev)
    top = torch.zeros(k, 1, dtype=enc1.dtype, device=dev)
    done_seqs, done_scores, done_alpha = [], [], []
    seqs_alpha = torch.ones(k, 1, enc1.size(1), dtype=enc1.dtype, device=dev)
    if kind == "lstm":
        enc = enc1.expand(k, -1, -1)
        h, c = init_hidden_state(sd, enc)
        prev_words = seqs[:, 0]
    else:
        mem = transformer_memory(sd, enc1).expand(k, -1, -1)
    step = 1
    while True:
        if kind == "lstm":
            h, c, alpha = lstm_step(sd, enc, sd["embedding.weight"][prev_words], h, c)
            logits = F.linear(h, sd["fc.weight"], sd["fc.bias"])
        else:
            logits = transformer_last_logits(sd, mem[:seqs.shape[0]], seqs, nheads)
        scores = top.expand_as(logits) + F.log_softmax(logits, dim=1)
        if step == 1:
            top_s, top_w = scores[0].topk(k, 0, True, True)
        else:
            top_s, top_w = scores.view(-1).topk(k, 0, True, True)
        prev = (top_w / vocab).long()       # caption.py:116,118 (true division then trunc)
        nxt = top_w % vocab
        if trace is not None:
            trace.append((top_s.clone(), prev.clone(), nxt.clone()))
        seqs = torch.cat([seqs[prev], nxt.unsqueeze(1)], dim=1)
        if kind == "lstm":
            seqs_alpha = torch.cat([seqs_alpha[prev], alpha[prev].unsqueeze(1)], dim=1)
        inc = [i for i, w in enumerate(nxt.tolist()) if w != end_tok]
        com = [i for i in range(len(nxt)) if i not in inc]
        if com:
            done_seqs.extend(seqs[com].tolist())
            done_scores.extend(top_s[com].tolist())
            if kind == "lstm":
                done_alpha.extend(seqs_alpha[com])
        k -= len(com)
        if k == 0:
            break
        seqs = seqs[inc]
        if kind == "lstm":
            seqs_alpha = seqs_alpha[inc]
            h, c, enc = h[prev[inc]], c[prev[inc]], enc[prev[inc]]
            prev_words = nxt[inc]
        top = top_s[inc].unsqueeze(1)
        # caption.py:147 (`if step > 50`, step counted from 1) and caption.py:249 (`if step + 1 >= 51`, step counted
        # from 0) are the same bound: at most max_steps + 1 = 51 decode steps.
        if step > max_steps:
            break
        step += 1
    if not done_scores:
        return None, done_seqs, done_scores
    best = done_scores.index(max(done_scores))
    if alphas_out is not None and kind == "lstm":
        alphas_out.append(done_alpha[best])
        alphas_out.append(done_alpha)           # every completed sequence's maps, in completion order
    return done_seqs[best], done_seqs, done_scores


# ------------------------------------------------------------------------------------------------
# train-step loss (trainMultiGPU.py:357-378)
# ------------------------------------------------------------------------------------------------
def packed_cross_entropy(scores, targets, decode_lengths):
    """pack_padded_sequence(...).data on both sides then CrossEntropyLoss (mean over packed rows).  The mean is
    order-independent, so rows are simply gathered by (b, t < len)."""
    rows, tg = [], []
    for b, l in enumerate(decode_lengths):
        rows.append(scores[b, :l])
        tg.append(targets[b, :l])
    return F.cross_entropy(torch.cat(rows), torch.cat(tg))


def train_loss_lstm(preds, caps_sorted, decode_lengths, alphas, alpha_c=1.0):
    """trainMultiGPU.py:364-369: CE over packed rows + alpha_c * ((1 - sum_t alpha)^2).mean()."""
    loss = packed_cross_entropy(preds, caps_sorted[:, 1:], decode_lengths)
    return loss + alpha_c * ((1.0 - alphas.sum(dim=1)) ** 2).mean()


def train_loss_transformer(preds, caps, decode_lengths):
    """trainMultiGPU.py:373-377."""
    return packed_cross_entropy(preds, caps[:, 1:], decode_lengths)


def free_running_loss(preds, seqs, caps, end_tok, pad_tok, max_len, alphas=None, alpha_c=1.0):
    """trainMultiGPU.py:448-458 (trainWithoutTeacherForcing): preprocessDecoderOutputForMetrics -> CrossEntropyLoss
    (+ the alpha regulariser for the LSTM decoder)."""
    from .metrics_oracle import preprocess_decoder_output_for_metrics
    scores, targets, _, _ = preprocess_decoder_output_for_metrics(preds, seqs, caps, end_tok, pad_tok, max_len)
    loss = F.cross_entropy(scores, targets)
    if alphas is not None:
        loss = loss + alpha_c * ((1.0 - alphas.sum(dim=1)) ** 2).mean()
    return loss
