"""``PositionalEncoding`` / ``TransformerDecoder`` — drop-in for ``models/transformerDecoder.py:14-27,53-168``.

Same constructor signature, attribute names and ``state_dict`` keys (``embedding``, ``pos_encoding.pe``,
``transformer_decoder.layers.{i}.{self_attn,multihead_attn}.{in_proj_weight,in_proj_bias,out_proj.*}``,
``linear1/2``, ``norm1/2/3``, ``fc_out``, ``encoder_proj``); same forward keyword names and return tuples.
The gensim pre-trained-embedding loader (models/transformerDecoder.py:29-42) is out of scope:
passing ``pretrained_embeddings_path`` raises.

Underneath (libccx kernels only):
  * activations are batch-first [B*T, 512] (the reference is seq-first; the math is per (b, t) row);
  * QKV / out-proj / FFN / fc_out / encoder_proj are tcgen05 GEMMs with bias / ReLU / residual epilogues,
    LayerNorm (post-norm, eps 1e-5) is one kernel that also emits the next GEMM's operand;
  * attention (52x52 causal+padding self-attention, 52x49 cross-attention) is a small-sequence kernel with the
    masks computed arithmetically — no (B*8,T,T) float mask is materialised;
  * cross-attention K/V of the image memory are computed once per layer per batch;
  * greedy decoding uses a KV cache (the reference re-runs the whole prefix every step, O(T^2)).
"""
import math

import torch
from torch import nn

from . import _lib
from ._host import (CcxEmbedding, CcxLinear, PreparedCache, RefreshPlan, any_requires_grad, host_copy,
                    stash_device_twin)
from ._lib import Operand, ptr


class PositionalEncoding(nn.Module):
    """models/transformerDecoder.py:14-27.  The table is built once on the host with the reference's formula;
    adding it is fused into the embedding kernel on the hot path."""

    def __init__(self, embed_dim, maxLen):
        super().__init__()
        pe = torch.zeros(maxLen, embed_dim)
        position = torch.arange(0, maxLen, dtype=torch.float).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, embed_dim, 2).float() * (-math.log(10000.0) / embed_dim))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer('pe', pe.unsqueeze(0))

    def forward(self, x):
        """x (B,T,D) + pe[:, :T] via the embed kernel's sibling path is not needed externally; this stays a view-add
        for API parity (caption.py:205) and is never called by this package's own hot path."""
        return x + self.pe[:, :x.size(1)]


class _MHAParams(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * d, d))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d))
        self.out_proj = CcxLinear(d, d)
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.constant_(self.out_proj.bias, 0.0)


class _Norm(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.bias = nn.Parameter(torch.zeros(d))


class _Layer(nn.Module):
    """Parameter tree of nn.TransformerDecoderLayer (same names / registration order)."""

    def __init__(self, d, dff):
        super().__init__()
        self.self_attn = _MHAParams(d)
        self.multihead_attn = _MHAParams(d)
        self.linear1 = CcxLinear(d, dff)
        self.linear2 = CcxLinear(dff, d)
        self.norm1, self.norm2, self.norm3 = _Norm(d), _Norm(d), _Norm(d)


class _Stack(nn.Module):
    """Parameter tree of nn.TransformerDecoder; ``forward`` keeps its seq-first signature for external callers
    (caption.py:204-213 pokes ``decoder.transformer_decoder(tgt, memory, tgt_mask=...)``) and runs on libccx."""

    def __init__(self, d, dff, n):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(d, dff) for _ in range(n)])
        self._owner = []     # [TransformerDecoder]; a list so the parent is not registered as a sub-module

    @torch.no_grad()
    def forward(self, tgt, memory, tgt_mask=None, tgt_key_padding_mask=None):
        """tgt (T,B,D), memory (P,B,D) seq-first -> (T,B,D).  Inference only (eval-mode math, no autograd).
        tgt_mask must be None or the causal mask (the only masks the reference ever passes)."""
        dec = self._owner[0]
        T, B, D = tgt.shape
        causal = 0
        if tgt_mask is not None:
            want = torch.ones(T, T, dtype=torch.bool, device=tgt.device).triu(1)
            got = tgt_mask if tgt_mask.dtype == torch.bool else (tgt_mask == float("-inf"))
            if got.shape != (T, T) or not torch.equal(got, want):
                raise ValueError("only the causal tgt_mask (generate_square_subsequent_mask) is supported")
            causal = 1
        Pw = dec._cache.get()
        cd = dec.compute_dtype
        x_plain = tgt.permute(1, 0, 2).contiguous().float().view(B * T, D)          # batch-first rows
        x_op = Operand.prepare(x_plain, cd)
        mem = Operand.prepare(memory.permute(1, 0, 2).contiguous().float().view(-1, D), cd)
        kpm = None if tgt_key_padding_mask is None else tgt_key_padding_mask.to(torch.uint8).contiguous()
        x_plain, _ = dec._run_layers(Pw, x_plain, x_op, mem, B, T, memory.shape[0], causal, kpm)
        return x_plain.view(B, T, D).permute(1, 0, 2)


class TransformerDecoder(nn.Module):
    def __init__(self, embed_dim, decoder_dim, vocab_size, maxLen, device, wordMap, pretrained_embeddings_path,
                 fine_tune_embeddings, dropout=0.5, encoder_dim=1024, num_heads=8, num_layers=6,
                 compute_dtype=torch.float32):
        super().__init__()
        _lib.dt_code(compute_dtype)
        if pretrained_embeddings_path:
            raise NotImplementedError("pre-trained (gensim) embeddings are out of scope; load them via state_dict")
        if embed_dim % num_heads or embed_dim % 128:
            raise ValueError("embed_dim must be a multiple of 128 and of num_heads (no fallback kernels)")
        self.encoder_dim, self.decoder_dim, self.embed_dim = encoder_dim, decoder_dim, embed_dim
        self.vocab_size, self.num_heads, self.num_layers = vocab_size, num_heads, num_layers
        self.maxLen = maxLen
        self.compute_dtype = compute_dtype
        self.dropout_p = dropout
        self.embedding = CcxEmbedding(vocab_size, embed_dim)
        self.pos_encoding = PositionalEncoding(embed_dim, maxLen)
        self.dropout = nn.Dropout(p=dropout)
        self.transformer_decoder = _Stack(embed_dim, decoder_dim, num_layers)
        self.fc_out = CcxLinear(embed_dim, vocab_size, compute_dtype=compute_dtype)
        self.encoder_proj = (CcxLinear(encoder_dim, embed_dim, compute_dtype=compute_dtype)
                             if encoder_dim != embed_dim else nn.Identity())
        self.device = device
        self._cache = PreparedCache(self)
        self.transformer_decoder._owner.append(self)
        self.inject_dropout = None   # tests: dict of multipliers, see _drop()
        self._want_alphas = False    # TransformerDecoderForAttentionViz: also produce the averaged cross-attention maps

    # ---- prepared weights ---------------------------------------------------------------------------------------
    def _prepare(self):
        cd, D = self.compute_dtype, self.embed_dim
        d = lambda p: p.detach()
        P = {"layers": []}
        for lyr in self.transformer_decoder.layers:
            sa, ca = lyr.self_attn, lyr.multihead_attn
            P["layers"].append({
                "sa_in": Operand.prepare(d(sa.in_proj_weight), cd), "sa_in_b": d(sa.in_proj_bias).contiguous(),
                "sa_out": Operand.prepare(d(sa.out_proj.weight), cd), "sa_out_b": d(sa.out_proj.bias).contiguous(),
                "ca_q": Operand.prepare(d(ca.in_proj_weight)[:D].contiguous(), cd),
                "ca_q_b": d(ca.in_proj_bias)[:D].contiguous(),
                "ca_kv": Operand.prepare(d(ca.in_proj_weight)[D:].contiguous(), cd),
                "ca_kv_b": d(ca.in_proj_bias)[D:].contiguous(),
                "ca_out": Operand.prepare(d(ca.out_proj.weight), cd), "ca_out_b": d(ca.out_proj.bias).contiguous(),
                "l1": Operand.prepare(d(lyr.linear1.weight), cd), "l1_b": d(lyr.linear1.bias).contiguous(),
                "l2": Operand.prepare(d(lyr.linear2.weight), cd), "l2_b": d(lyr.linear2.bias).contiguous(),
                "n": [(d(n.weight).contiguous(), d(n.bias).contiguous()) for n in (lyr.norm1, lyr.norm2, lyr.norm3)],
            })
        P["fc"] = Operand.prepare(d(self.fc_out.weight), cd)
        P["fc_b"] = d(self.fc_out.bias).contiguous()
        if isinstance(self.encoder_proj, nn.Identity):
            P["proj"] = None
        else:
            P["proj"] = Operand.prepare(d(self.encoder_proj.weight), cd)
            P["proj_b"] = d(self.encoder_proj.bias).contiguous()
        P["pe"] = d(self.pos_encoding.pe)[0].contiguous().float()
        return P

    def _refresh_plan(self, P):
        """One-launch refresh of the bf16 copies ``_prepare()`` made (see _host.RefreshPlan).  Everything else in P is
        a view of its parameter, except the two halves of the cross-attention in-proj bias (copied as fp32)."""
        if self.compute_dtype != torch.bfloat16:
            return None
        D = self.embed_dim
        plan = RefreshPlan()
        for lyr, L in zip(self.transformer_decoder.layers, P["layers"]):
            sa, ca = lyr.self_attn, lyr.multihead_attn
            plan.add(L["sa_in"].hi, sa.in_proj_weight).add(L["sa_out"].hi, sa.out_proj.weight)
            plan.add(L["ca_q"].hi, ca.in_proj_weight[:D]).add(L["ca_kv"].hi, ca.in_proj_weight[D:])
            for key, src in (("ca_q_b", ca.in_proj_bias[:D]), ("ca_kv_b", ca.in_proj_bias[D:])):
                if L[key].data_ptr() != src.data_ptr():
                    plan.add(L[key], src)
            plan.add(L["ca_out"].hi, ca.out_proj.weight)
            plan.add(L["l1"].hi, lyr.linear1.weight).add(L["l2"].hi, lyr.linear2.weight)
        plan.add(P["fc"].hi, self.fc_out.weight)
        if P["proj"] is not None:
            plan.add(P["proj"].hi, self.encoder_proj.weight)
        return plan

    # ---- building blocks ----------------------------------------------------------------------------------------
    def _memory(self, Pw, encoder_out):
        """encoder_proj (models/transformerDecoder.py:94-95) -> GEMM operand [B*P, D]."""
        B, E = encoder_out.size(0), encoder_out.size(-1)
        enc = encoder_out.reshape(B, -1, E).float().contiguous()
        Pn = enc.size(1)
        enc_op = Operand.prepare(enc.view(B * Pn, E), self.compute_dtype)
        if Pw["proj"] is None:
            return enc_op, Pn
        if self.compute_dtype == torch.bfloat16:
            mem = Operand(_lib.linear(enc_op, Pw["proj"], bias=Pw["proj_b"], out_dtype=torch.bfloat16), None,
                          torch.bfloat16)
        else:
            mem = _lib.linear(enc_op, Pw["proj"], bias=Pw["proj_b"], split=True)
        return mem, Pn

    def _ln(self, y, gb, rows):
        """post-norm LayerNorm(eps 1e-5): fp32 rows -> (plain fp32, GEMM operand)."""
        D, cd = self.embed_dim, self.compute_dtype
        plain = torch.empty((rows, D), dtype=torch.float32, device=y.device)
        op = Operand.empty((rows, D), cd, y.device)
        _lib.check(_lib.lib().ccx_ln_rows(ptr(y), ptr(gb[0]), ptr(gb[1]), ptr(op.hi), op.lo_ptr, ptr(plain), rows, D,
                                          1e-5, _lib.dt_code(cd), 0, 1, 1, _lib.stream_ptr()), "ln_rows")
        return plain, op

    def _mha(self, q, q_sb, q_st, k, k_sb, k_st, v, B, Tq, Tk, causal, q_pos0, key_pad, prob_mask, kv_group, dev,
             probs_out=None):
        D, H, cd = self.embed_dim, self.num_heads, self.compute_dtype
        ctx = Operand.empty((B * Tq, D), cd, dev)
        _lib.check(_lib.lib().ccx_mha_small(q, q_sb, q_st, k, k_sb, k_st, v, k_sb, k_st, ptr(ctx.hi), ctx.lo_ptr,
                                            _lib.dt_code(cd), Tq * D, D, ptr(key_pad), ptr(prob_mask),
                                            ptr(probs_out), B, H, Tq, Tk, D // H, causal, q_pos0,
                                            1.0 / math.sqrt(D // H), kv_group, _lib.stream_ptr()), "mha_small")
        return ctx

    def _mha_decode(self, q, q_sb, k, k_sb, k_st, v, rows, Tk, kv_rows, kv_group, dev):
        """Single-query attention over the KV cache / image memory (ccx_mha_decode)."""
        D, H, cd = self.embed_dim, self.num_heads, self.compute_dtype
        ctx = Operand.empty((rows, D), cd, dev)
        _lib.check(_lib.lib().ccx_mha_decode(q, q_sb, k, k_sb, k_st, v, k_sb, k_st, ptr(ctx.hi), ctx.lo_ptr,
                                             _lib.dt_code(cd), D, ptr(kv_rows),
                                             0 if kv_rows is None else kv_rows.stride(0), rows, H, Tk, D // H,
                                             kv_group, 1.0 / math.sqrt(D // H), _lib.stream_ptr()), "mha_decode")
        return ctx

    def _linear_op(self, a, w, bias, act=_lib.ACT_NONE):
        """GEMM whose output is directly the next GEMM's operand."""
        if self.compute_dtype == torch.bfloat16:
            return Operand(_lib.linear(a, w, bias=bias, act=act, out_dtype=torch.bfloat16), None, torch.bfloat16)
        return _lib.linear(a, w, bias=bias, act=act, split=True)

    def _cross_kv(self, Pw, mem):
        """K/V projections of the image memory for every layer (time-invariant; the reference recomputes them in
        every layer call of every greedy step)."""
        return [_lib.linear(mem, lw["ca_kv"], bias=lw["ca_kv_b"]) for lw in Pw["layers"]]

    def _head_mean(self, probs, prob_mask, alphas, a_sb, a_st, B, Tq, Tk, layer, row_active=None, over="heads"):
        """alphas (+)= one layer's share of the averaged cross-attention map (ccx_attn_head_mean), probs (B,H,Tq,Tk):
        over="heads":     alphas[b, t, :] = mean over layers and heads  (greedy, transformerDecoderAttVis.py:223-226)
        over="positions": alphas[h, b, :] = mean over layers and target positions — what the reference's teacher-forced
                          path really computes: (L,B,H,T,P).mean(dim=(0,3)).permute(1,0,2), :163-165."""
        H, L = self.num_heads, self.num_layers
        if over == "heads":
            args = (H * Tq * Tk, Tq * Tk, Tk, ptr(prob_mask), ptr(row_active), alphas, a_sb, a_st, B, H, Tq, Tk,
                    1.0 / (H * L))
        else:      # summed axis = positions (stride Tk), kept axis = heads (stride Tq*Tk)
            args = (H * Tq * Tk, Tk, Tq * Tk, ptr(prob_mask), ptr(row_active), alphas, a_sb, a_st, B, Tq, H, Tk,
                    1.0 / (Tq * L))
        _lib.check(_lib.lib().ccx_attn_head_mean(ptr(probs), *args, 1 if layer > 0 else 0, _lib.stream_ptr()),
                   "attn_head_mean")

    def _run_layers(self, Pw, x_plain, x_op, mem, B, T, Pn, causal, kpm, alphas=None):
        """The 6 post-norm layers on batch-first rows (eval-mode math): torch/nn/modules/transformer.py:1089-1199.
        alphas (H, B, Pn), optional: receives the AttVis teacher-forcing map (see _head_mean over="positions")."""
        D = self.embed_dim
        dev = x_plain.device
        kvs = self._cross_kv(Pw, mem)
        probs = None
        if alphas is not None:
            probs = torch.empty((B, self.num_heads, T, Pn), dtype=torch.float32, device=dev)
        for li, (lw, kv) in enumerate(zip(Pw["layers"], kvs)):
            qkv = _lib.linear(x_op, lw["sa_in"], bias=lw["sa_in_b"])
            ctx = self._mha(ptr(qkv), T * 3 * D, 3 * D, qkv.data_ptr() + 4 * D, T * 3 * D, 3 * D,
                            qkv.data_ptr() + 8 * D, B, T, T, causal, 0, kpm, None, 1, dev)
            y = _lib.linear(ctx, lw["sa_out"], bias=lw["sa_out_b"], residual=x_plain)
            x_plain, x_op = self._ln(y, lw["n"][0], B * T)
            q = _lib.linear(x_op, lw["ca_q"], bias=lw["ca_q_b"])
            ctx = self._mha(ptr(q), T * D, D, ptr(kv), Pn * 2 * D, 2 * D, kv.data_ptr() + 4 * D, B, T, Pn, 0, 0,
                            None, None, 1, dev, probs_out=probs)
            if alphas is not None:
                self._head_mean(probs, None, ptr(alphas), Pn, B * Pn, B, T, Pn, li, over="positions")
            y = _lib.linear(ctx, lw["ca_out"], bias=lw["ca_out_b"], residual=x_plain)
            x_plain, x_op = self._ln(y, lw["n"][1], B * T)
            h = self._linear_op(x_op, lw["l1"], lw["l1_b"], act=_lib.ACT_RELU)
            y = _lib.linear(h, lw["l2"], bias=lw["l2_b"], residual=x_plain)
            x_plain, x_op = self._ln(y, lw["n"][2], B * T)
        return x_plain, x_op

    def _drop(self, name, shape, dev):
        if not self.training or self.dropout_p == 0:
            return None
        if self.inject_dropout is not None:
            return self.inject_dropout[name].to(device=dev, dtype=torch.float32).contiguous()
        keep = 1.0 - self.dropout_p
        return (torch.bernoulli(torch.full(shape, keep, device=dev)) / keep).contiguous()

    def enable_cuda_graph(self, enabled=True):
        """Training option (not in the reference): replay the teacher-forced forward / backward as CUDA graphs;
        see transformer_train.enable_cuda_graph for the static-buffer contract."""
        from .transformer_train import enable_cuda_graph
        return enable_cuda_graph(self, enabled)

    # ---- reference API ------------------------------------------------------------------------------------------
    def forwardWithTeacherForcing(self, encoder_out, encoded_captions, caption_lengths, tgt_key_padding_mask):
        """models/transformerDecoder.py:88-108."""
        if torch.is_grad_enabled() and any_requires_grad(self):
            from .transformer_train import transformer_teacher_forcing_with_grad
            return transformer_teacher_forcing_with_grad(self, encoder_out, encoded_captions, caption_lengths,
                                                         tgt_key_padding_mask)
        if self.training and self.dropout_p > 0:
            raise RuntimeError("train-mode dropout without autograd is not a reference use case; call .eval() "
                               "for inference or enable grad for training")
        _lib.require_cuda(encoder_out, "encoder_out")
        B = encoder_out.size(0)
        decode_lengths = (host_copy(caption_lengths).reshape(-1) - 1).tolist()
        stash_device_twin(decode_lengths, caption_lengths.reshape(-1) - 1)     # the same values, on the device
        dev = encoder_out.device
        D, V, cd = self.embed_dim, self.vocab_size, self.compute_dtype
        T = encoded_captions.size(1)
        Pw = self._cache.get()
        L, st, code = _lib.lib(), _lib.stream_ptr(), _lib.dt_code(cd)
        mem, Pn = self._memory(Pw, encoder_out)
        caps = encoded_captions.contiguous()
        x_plain = torch.empty((B * T, D), dtype=torch.float32, device=dev)
        x_op = Operand.empty((B * T, D), cd, dev)
        _lib.check(L.ccx_embed_rows(ptr(caps), T, 0, ptr(self.embedding.weight), V, D, ptr(Pw["pe"]), None,
                                    ptr(x_plain), T * D, D, ptr(x_op.hi), x_op.lo_ptr, code, T * D, D, B, T, st),
                   "embed_rows")
        kpm = None
        if tgt_key_padding_mask is not None:
            kpm = tgt_key_padding_mask.to(torch.uint8).contiguous()
        alphas = (torch.empty((self.num_heads, B, Pn), dtype=torch.float32, device=dev) if self._want_alphas
                  else None)
        x_plain, x_op = self._run_layers(Pw, x_plain, x_op, mem, B, T, Pn, 1, kpm, alphas=alphas)
        predictions = torch.empty((B, T, V), dtype=torch.float32, device=dev)
        _lib.linear(x_op, Pw["fc"], bias=Pw["fc_b"], out=predictions.view(B * T, V))
        self._last_alphas = alphas
        return predictions, encoded_captions, decode_lengths

    @torch.no_grad()
    def decode_step_cached(self, Pw, state, t, rows):
        """One KV-cache decode step for `rows` sequences whose token at position t is in state['tokens'][:, t].
        Returns the GEMM operand of the last layer's output for position t ([rows, D])."""
        L, st = _lib.lib(), _lib.stream_ptr()
        D, V, cd = self.embed_dim, self.vocab_size, self.compute_dtype
        code = _lib.dt_code(cd)
        dev = state["tokens"].device
        Tm, Pn, g = state["Tmax"], state["Pn"], state["kv_group"]
        x_plain = torch.empty((rows, D), dtype=torch.float32, device=dev)
        x_op = Operand.empty((rows, D), cd, dev)
        tok = state["tokens"]
        _lib.check(L.ccx_embed_rows(ptr(tok), tok.stride(0), t, ptr(self.embedding.weight), V, D, ptr(Pw["pe"]), None,
                                    ptr(x_plain), D, 0, ptr(x_op.hi), x_op.lo_ptr, code, D, 0, rows, 1, st),
                   "embed_rows")
        for lw, kv, cache in zip(Pw["layers"], state["cross_kv"], state["cache"]):
            # qkv of the new token is written straight into the cache row (b, t, :)
            _lib.linear(x_op, lw["sa_in"], bias=lw["sa_in_b"], out=cache[:rows, t])
            base = cache.data_ptr()
            ctx = self._mha_decode(base + 4 * (t * 3 * D), Tm * 3 * D, base + 4 * D, Tm * 3 * D, 3 * D, base + 8 * D,
                                   rows, t + 1, state.get("kv_rows"), 1, dev)
            y = _lib.linear(ctx, lw["sa_out"], bias=lw["sa_out_b"], residual=x_plain)
            x_plain, x_op = self._ln(y, lw["n"][0], rows)
            q = _lib.linear(x_op, lw["ca_q"], bias=lw["ca_q_b"])
            if g > 1:
                # beam search: the g beams of an image are g query rows of ONE (image, head) CTA, so the image's K/V
                # head slice is staged once instead of being re-read by every beam
                ctx = self._mha(ptr(q), g * D, D, ptr(kv), Pn * 2 * D, 2 * D, kv.data_ptr() + 4 * D, rows // g, g, Pn,
                                0, 0, None, None, 1, dev)
            elif state.get("alphas") is not None:
                # attention-map variant: the staged kernel can also write the probabilities of the new token
                probs = state["probs"]
                ctx = self._mha(ptr(q), D, D, ptr(kv), Pn * 2 * D, 2 * D, kv.data_ptr() + 4 * D, rows, 1, Pn, 0, 0,
                                None, None, 1, dev, probs_out=probs)
                al = state["alphas"]
                self._head_mean(probs, None, al.data_ptr() + 4 * t * Pn, al.stride(0), 0, rows, 1, Pn,
                                state["layer_index"][id(cache)], row_active=state["active"])
            else:
                ctx = self._mha_decode(ptr(q), D, ptr(kv), Pn * 2 * D, 2 * D, kv.data_ptr() + 4 * D, rows, Pn, None,
                                       1, dev)
            y = _lib.linear(ctx, lw["ca_out"], bias=lw["ca_out_b"], residual=x_plain)
            x_plain, x_op = self._ln(y, lw["n"][1], rows)
            h = self._linear_op(x_op, lw["l1"], lw["l1_b"], act=_lib.ACT_RELU)
            y = _lib.linear(h, lw["l2"], bias=lw["l2_b"], residual=x_plain)
            x_plain, x_op = self._ln(y, lw["n"][2], rows)
        return x_op

    def new_decode_state(self, Pw, encoder_out, rows, Tmax, kv_group=1):
        """Per-batch decode state: projected memory K/V per layer, empty KV caches, token buffer."""
        dev = encoder_out.device
        mem, Pn = self._memory(Pw, encoder_out)
        D = self.embed_dim
        return {"cross_kv": self._cross_kv(Pw, mem), "Pn": Pn, "Tmax": Tmax, "kv_group": kv_group,
                "cache": [torch.empty((rows, Tmax, 3 * D), dtype=torch.float32, device=dev)
                          for _ in range(self.num_layers)],
                "tokens": torch.zeros((rows, Tmax + 1), dtype=torch.long, device=dev)}

    def forwardWithoutTeacherForcing(self, encoder_out, wordMap, maxDecodeLen):
        """models/transformerDecoder.py:110-160 (greedy), with a KV cache instead of prefix recomputation.  With
        autograd enabled (trainWithoutTeacherForcing, trainMultiGPU.py:444-460) the predictions carry gradients."""
        if torch.is_grad_enabled() and (encoder_out.requires_grad or any_requires_grad(self)):
            from .transformer_train import transformer_free_running_with_grad
            return transformer_free_running_with_grad(self, encoder_out, wordMap, maxDecodeLen)
        return self._greedy(encoder_out, wordMap, maxDecodeLen)

    @torch.no_grad()
    def _greedy(self, encoder_out, wordMap, maxDecodeLen, dropout_free=False):
        _lib.require_cuda(encoder_out, "encoder_out")
        if self.training and self.dropout_p > 0 and not dropout_free:
            raise RuntimeError("free-running decode in train mode (dropout live) is not built; call .eval()")
        B, T, V = encoder_out.size(0), int(maxDecodeLen), self.vocab_size
        if T > self.maxLen:
            raise ValueError(f"maxDecodeLen {T} exceeds the positional-encoding table ({self.maxLen})")
        dev = encoder_out.device
        Pw = self._cache.get()
        L, st = _lib.lib(), _lib.stream_ptr()
        state = self.new_decode_state(Pw, encoder_out, B, T)
        tokens = state["tokens"]
        tokens[:, 0] = wordMap['<start>']
        predictions = torch.zeros((B, T, V), dtype=torch.float32, device=dev)
        sequences = torch.zeros((B, T), dtype=torch.long, device=dev)
        active = torch.ones(B, dtype=torch.float32, device=dev)
        self._last_alphas = None
        if self._want_alphas:
            Pn = state["Pn"]
            state["alphas"] = self._last_alphas = torch.zeros((B, T, Pn), dtype=torch.float32, device=dev)
            state["probs"] = torch.empty((B, self.num_heads, 1, Pn), dtype=torch.float32, device=dev)
            state["active"] = active
            state["layer_index"] = {id(c): i for i, c in enumerate(state["cache"])}
        for t in range(T):
            x_op = self.decode_step_cached(Pw, state, t, B)
            p_t = predictions[:, t]
            _lib.linear(x_op, Pw["fc"], bias=Pw["fc_b"], rowscale=active, rows_per_group=1, out=p_t)
            _lib.check(L.ccx_greedy_next(ptr(p_t), T * V, B, V, t, T, ptr(sequences), ptr(active),
                                         tokens.data_ptr() + 8 * (t + 1), T + 1, wordMap['<end>'], st), "greedy_next")
            if t % 8 == 7 and not bool(active.any()):
                break
        return predictions, sequences

    def forward(self, teacherForcing, encoder_out, encoded_captions=None, caption_lengths=None,
                tgt_key_padding_mask=None, wordMap=None, maxDecodeLen=None):
        """models/transformerDecoder.py:162-168."""
        if teacherForcing is True:
            return self.forwardWithTeacherForcing(encoder_out, encoded_captions, caption_lengths,
                                                  tgt_key_padding_mask)
        return self.forwardWithoutTeacherForcing(encoder_out, wordMap, maxDecodeLen)


class TransformerDecoderForAttentionViz(TransformerDecoder):
    """models/transformerDecoderAttVis.py:105-236: the TransformerDecoder that also returns ``alphas`` for
    attention-map visualisation — greedy: the new token's cross-attention weights averaged over the 6 layers and 8
    heads, (B, maxDecodeLen, num_pixels); teacher forcing: what the reference computes there, the average over layers
    and target positions, (num_heads, B, num_pixels).  Same arithmetic as TransformerDecoder (its CustomTransformerDecoderLayer restates the post-norm
    nn.TransformerDecoderLayer, :63-97); the maps come out of the fused attention kernel's optional probability
    output and one head/layer-mean kernel per layer.  state_dict keys follow the reference class
    (``decoder_layers.{i}.…``).  The maps are returned detached (the reference only plots them; its alpha
    regulariser for this decoder is commented out, trainMultiGPU.py:377,458)."""

    def __init__(self, embed_dim, decoder_dim, vocab_size, maxLen, device, dropout=0.5, encoder_dim=1024, num_heads=8,
                 num_layers=6, compute_dtype=torch.float32):
        super().__init__(embed_dim, decoder_dim, vocab_size, maxLen, device, None, None, True, dropout=dropout,
                         encoder_dim=encoder_dim, num_heads=num_heads, num_layers=num_layers,
                         compute_dtype=compute_dtype)
        self._want_alphas = True
        self._register_state_dict_hook(self._rename_out)
        self._register_load_state_dict_pre_hook(self._rename_in)

    _OURS, _THEIRS = "transformer_decoder.layers.", "decoder_layers."

    @property
    def decoder_layers(self):
        return self.transformer_decoder.layers

    @staticmethod
    def _rename_out(module, state_dict, prefix, local_metadata):
        a, b = prefix + module._OURS, prefix + module._THEIRS
        for k in [k for k in state_dict if k.startswith(a)]:
            state_dict[b + k[len(a):]] = state_dict.pop(k)
        return state_dict

    def _rename_in(self, state_dict, prefix, *args):
        a, b = prefix + self._THEIRS, prefix + self._OURS
        for k in [k for k in state_dict if k.startswith(a)]:
            state_dict[b + k[len(a):]] = state_dict.pop(k)

    def forwardWithTeacherForcing(self, encoder_out, encoded_captions, caption_lengths, tgt_key_padding_mask):
        """models/transformerDecoderAttVis.py:133-167 -> (predictions, encoded_captions, decode_lengths, alphas)."""
        out = super().forwardWithTeacherForcing(encoder_out, encoded_captions, caption_lengths, tgt_key_padding_mask)
        return out + (self._last_alphas,)

    def forwardWithoutTeacherForcing(self, encoder_out, wordMap, maxDecodeLen):
        """models/transformerDecoderAttVis.py:170-228 -> (predictions, sequences, alphas)."""
        out = super().forwardWithoutTeacherForcing(encoder_out, wordMap, maxDecodeLen)
        return out + (self._last_alphas,)

    def forward(self, teacherForcing, encoder_out, encoded_captions=None, caption_lengths=None,
                tgt_key_padding_mask=None, wordMap=None, maxDecodeLen=None):
        """models/transformerDecoderAttVis.py:231-236."""
        if teacherForcing is True:
            return self.forwardWithTeacherForcing(encoder_out, encoded_captions, caption_lengths,
                                                  tgt_key_padding_mask)
        return self.forwardWithoutTeacherForcing(encoder_out, wordMap, maxDecodeLen)
