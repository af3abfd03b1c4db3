"""``Attention`` / ``DecoderWithAttention`` — drop-in for the reference's ``models/decoder.py:16-172``.

Same constructor signature, attribute names and ``state_dict`` keys (``attention.{encoder_att,decoder_att,full_att}``,
``embedding``, ``decode_step.{weight_ih,weight_hh,bias_ih,bias_hh}``, ``init_h``, ``init_c``, ``f_beta``, ``fc``); same
``forward(teacherForcing, encoder_out, encoded_captions, caption_lengths, wordMap, maxDecodeLen)`` return tuples.

What runs underneath (all libccx kernels, no torch compute on the path):
  * ``encoder_att`` is time-invariant and is hoisted out of the step loop (the reference recomputes it 51x);
  * per step: ONE tcgen05 GEMM for [decoder_att | f_beta](h), one fused attention kernel
    (add+ReLU+dot+softmax-over-pixels+weighted-sum+sigmoid gate), ONE tcgen05 GEMM for the LSTM gates over the
    concatenated operand [emb_t | awe | h] x [W_ih | W_hh]^T (no torch.cat), one point-wise LSTM kernel;
  * teacher forcing: the vocabulary projection ``fc`` is hoisted into ONE GEMM over all (b, t) rows whose epilogue
    zeroes the rows past each caption's length (the reference allocates zeros on the host and scatters);
  * greedy: per-step fc GEMM with finished-row masking in the epilogue + an argmax/bookkeeping kernel; no
    ``nonzero()`` host sync per step.
Extra ctor kwarg: ``compute_dtype`` (float32 = 3xTF32, bfloat16).  There is no CPU path.
"""
import ctypes
import os

import torch
from torch import nn

from . import _lib
from ._host import (CcxEmbedding, CcxLinear, PreparedCache, RefreshPlan, any_requires_grad, host_copy,
                    stash_device_twin)
from ._lib import Operand, ptr


_INT_ARRAYS = {}


def _int_array_type(n):
    """ctypes array types are cached: creating `c_int32 * n` per call leaves a new (cyclic) type object behind."""
    t = _INT_ARRAYS.get(n)
    if t is None:
        t = _INT_ARRAYS[n] = ctypes.c_int32 * n
    return t


class Attention(nn.Module):
    """models/decoder.py:16-31."""

    def __init__(self, encoder_dim, decoder_dim, attention_dim, compute_dtype=torch.float32):
        super().__init__()
        self.encoder_att = CcxLinear(encoder_dim, attention_dim, compute_dtype=compute_dtype)
        self.decoder_att = CcxLinear(decoder_dim, attention_dim, compute_dtype=compute_dtype)
        self.full_att = CcxLinear(attention_dim, 1, compute_dtype=compute_dtype)
        self.relu = nn.ReLU()
        self.softmax = nn.Softmax(dim=1)

    def forward(self, encoder_out, decoder_hidden):
        """(b,P,E), (b,D) -> (awe (b,E), alpha (b,P)); un-hoisted form for external callers (caption.py:98)."""
        b, P, E = encoder_out.shape
        enc = encoder_out.contiguous().float()
        att1 = self.encoder_att(enc.view(b * P, E))
        att2 = self.decoder_att(decoder_hidden)
        A = att2.shape[1]
        alpha = torch.empty((b, P), dtype=torch.float32, device=enc.device)
        awe = torch.empty((b, E), dtype=torch.float32, device=enc.device)
        w_f = self.full_att.weight.detach().view(-1)
        _lib.check(_lib.lib().ccx_bahdanau_attention(
            ptr(att1), ptr(att2), A, ptr(w_f), ptr(self.full_att.bias.detach()), ptr(enc), None, ptr(alpha), P,
            ptr(awe), None, _lib.CCX_F32, E, b, P, A, E, 0, 1, _lib.stream_ptr()), "attention")
        return awe, alpha


class _LSTMCellParams(nn.Module):
    """Parameter holder with nn.LSTMCell's names/shapes/init (uniform +-1/sqrt(hidden)); forward on libccx."""

    def __init__(self, input_size, hidden_size, compute_dtype):
        super().__init__()
        self.input_size, self.hidden_size, self.compute_dtype = input_size, hidden_size, compute_dtype
        k = 1.0 / hidden_size ** 0.5
        self.weight_ih = nn.Parameter(torch.empty(4 * hidden_size, input_size).uniform_(-k, k))
        self.weight_hh = nn.Parameter(torch.empty(4 * hidden_size, hidden_size).uniform_(-k, k))
        self.bias_ih = nn.Parameter(torch.empty(4 * hidden_size).uniform_(-k, k))
        self.bias_hh = nn.Parameter(torch.empty(4 * hidden_size).uniform_(-k, k))

    def forward(self, x, state):
        """(x (b,In), (h, c)) -> (h', c'), as nn.LSTMCell (external callers: caption.py:102)."""
        h, c = state
        cd = self.compute_dtype
        w = Operand.prepare(torch.cat([self.weight_ih.detach(), self.weight_hh.detach()], dim=1), cd)
        a = Operand.prepare(torch.cat([x.float(), h.float()], dim=1), cd)
        gates = _lib.linear(a, w, bias=(self.bias_ih + self.bias_hh).detach())
        b, D = h.shape
        c2 = torch.empty_like(c, dtype=torch.float32)
        h2 = torch.empty_like(c2)
        _lib.check(_lib.lib().ccx_lstm_pointwise(ptr(gates), 4 * D, ptr(c.float().contiguous()), ptr(c2), None, None, 0,
                                                 None, None, 0, _lib.CCX_F32, None, 0, ptr(h2), D, b, D,
                                                 _lib.stream_ptr()), "lstm_pointwise")
        return h2, c2


class DecoderWithAttention(nn.Module):
    def __init__(self, attention_dim, embed_dim, decoder_dim, vocab_size, device, encoder_dim=1024, dropout=0.5,
                 compute_dtype=torch.float32):
        super().__init__()
        _lib.dt_code(compute_dtype)
        self.encoder_dim = encoder_dim
        self.attention_dim = attention_dim
        self.embed_dim = embed_dim
        self.decoder_dim = decoder_dim
        self.vocab_size = vocab_size
        self.compute_dtype = compute_dtype
        self.dropout_p = dropout
        cd = compute_dtype
        self.attention = Attention(encoder_dim, decoder_dim, attention_dim, cd)
        self.embedding = CcxEmbedding(vocab_size, embed_dim)
        self.dropout = nn.Dropout(p=dropout)                      # mask source only (see _dropout_mask)
        self.decode_step = _LSTMCellParams(embed_dim + encoder_dim, decoder_dim, cd)
        self.init_h = CcxLinear(encoder_dim, decoder_dim, compute_dtype=cd)
        self.init_c = CcxLinear(encoder_dim, decoder_dim, compute_dtype=cd)
        self.f_beta = CcxLinear(decoder_dim, encoder_dim, compute_dtype=cd)
        self.sigmoid = nn.Sigmoid()
        self.fc = CcxLinear(decoder_dim, vocab_size, compute_dtype=cd)
        self.init_weights()
        self.device = device
        self._cache = PreparedCache(self)
        self.use_persist = os.environ.get("CCX_LSTM_PERSIST", "1") != "0"
        self._persist_dbg = None
        self.fixed_T = False          # True: step buffers always span captions.shape[1] - 1 steps (CUDA-graph replay)
        self.inject_dropmask = None   # tests: (B, T, decoder_dim) multiplier used instead of a fresh Bernoulli draw

    def init_weights(self):
        """models/decoder.py:58-61."""
        self.embedding.weight.data.uniform_(-0.1, 0.1)
        self.fc.bias.data.fill_(0)
        self.fc.weight.data.uniform_(-0.1, 0.1)

    # ---- prepared (kernel-side) weights ------------------------------------------------------------------------
    def _prepare(self):
        cd = self.compute_dtype
        d = lambda p: p.detach()
        att = self.attention
        P = {}
        P["w_enc_att"] = Operand.prepare(d(att.encoder_att.weight), cd)
        P["b_enc_att"] = d(att.encoder_att.bias).contiguous()
        P["w_h"] = Operand.prepare(torch.cat([d(att.decoder_att.weight), d(self.f_beta.weight)], dim=0), cd)
        P["b_h"] = torch.cat([d(att.decoder_att.bias), d(self.f_beta.bias)]).contiguous()
        P["w_f"] = d(att.full_att.weight).reshape(-1).contiguous()
        P["b_f"] = d(att.full_att.bias).contiguous()
        P["w_init_h"] = Operand.prepare(d(self.init_h.weight), cd)
        P["w_init_c"] = Operand.prepare(d(self.init_c.weight), cd)
        ds = self.decode_step
        P["w_lstm"] = Operand.prepare(torch.cat([d(ds.weight_ih), d(ds.weight_hh)], dim=1), cd)
        P["b_lstm"] = (d(ds.bias_ih) + d(ds.bias_hh)).contiguous()
        P["w_fc"] = Operand.prepare(d(self.fc.weight), cd)
        if self._persist_dims_ok():
            # persistent recurrence kernel (csrc/lstm_persist.cu): gate rows re-ordered so that each CTA's 32 rows are
            # {i,f,g,o} x 8 hidden units: new row 32*(j//8) + 8*gate + j%8  <-  torch row gate*D + j
            D, Emb = self.decoder_dim, self.embed_dim
            perm = torch.arange(4 * D, device=ds.weight_ih.device).view(4, D // 8, 8).permute(1, 0, 2).reshape(-1)
            P["w2p"] = Operand.prepare(torch.cat([d(ds.weight_hh), d(ds.weight_ih)[:, Emb:]], dim=1)[perm], cd)
            P["w2p_emb"] = Operand.prepare(d(ds.weight_ih)[:, :Emb][perm], cd)
            P["b_perm"] = P["b_lstm"][perm].contiguous()
            # backward kernel: W^T with the contraction (gate) index innermost
            P["wx"] = Operand.prepare(torch.cat([d(ds.weight_ih)[:, Emb:], d(ds.weight_hh)], dim=1).t().contiguous(), cd)
            P["wht"] = Operand.prepare(torch.cat([d(att.decoder_att.weight), d(self.f_beta.weight)], dim=0)
                                       .t().contiguous(), cd)
            P["w_emb_t"] = Operand.prepare(d(ds.weight_ih)[:, :Emb].t().contiguous(), cd)      # [Emb, 4D]
        return P

    def _refresh_plan(self, P):
        """The refresh of every entry of ``_prepare()`` that is not a view of its parameter, as one launch (see
        _host.RefreshPlan); bf16 compute only (the tf32 hi/lo split keeps the per-operand path)."""
        if self.compute_dtype != torch.bfloat16:
            return None
        att, ds = self.attention, self.decode_step
        A, Emb, D = self.attention_dim, self.embed_dim, self.decoder_dim
        w_ih, w_hh = ds.weight_ih, ds.weight_hh
        plan = RefreshPlan()
        plan.add(P["w_enc_att"].hi, att.encoder_att.weight)
        plan.add(P["w_h"].hi[:A], att.decoder_att.weight).add(P["w_h"].hi[A:], self.f_beta.weight)
        plan.add(P["b_h"][:A], att.decoder_att.bias).add(P["b_h"][A:], self.f_beta.bias)
        plan.add(P["w_init_h"].hi, self.init_h.weight).add(P["w_init_c"].hi, self.init_c.weight)
        K1 = w_ih.shape[1]
        plan.add(P["w_lstm"].hi[:, :K1], w_ih).add(P["w_lstm"].hi[:, K1:], w_hh)
        plan.add(P["b_lstm"], ds.bias_ih, src2=ds.bias_hh)
        plan.add(P["w_fc"].hi, self.fc.weight)
        if "w2p" in P:
            perm = torch.arange(4 * D, device=w_ih.device).view(4, D // 8, 8).permute(1, 0, 2).reshape(-1)
            plan.add(P["w2p"].hi[:, :D], w_hh, row_map=perm).add(P["w2p"].hi[:, D:], w_ih[:, Emb:], row_map=perm)
            plan.add(P["w2p_emb"].hi, w_ih[:, :Emb], row_map=perm)
            plan.add(P["b_perm"].view(-1, 1), ds.bias_ih.view(-1, 1), src2=ds.bias_hh.view(-1, 1), row_map=perm)
            plan.add(P["wx"].hi[:K1 - Emb], w_ih[:, Emb:], transpose=True).add(P["wx"].hi[K1 - Emb:], w_hh, transpose=True)
            plan.add(P["wht"].hi[:, :A], att.decoder_att.weight, transpose=True)
            plan.add(P["wht"].hi[:, A:], self.f_beta.weight, transpose=True)
            plan.add(P["w_emb_t"].hi, w_ih[:, :Emb], transpose=True)
        return plan

    def _persist_dims_ok(self):
        return (self.compute_dtype == torch.bfloat16 and self.decoder_dim == 512 and self.attention_dim == 512 and
                self.embed_dim == 512 and self.encoder_dim == 1024)

    def _persist_ok(self, B, Pn):
        """The one-kernel recurrence (ccx_lstm_tf_forward_persist) serves this call?  Otherwise the per-step launch
        loop of the same library runs (fp32 / 3xTF32 mode, B > 32, 14x14 feature maps)."""
        return (self.use_persist and self._persist_dims_ok() and bool(_lib.lib().ccx_lstm_persist_supported(
            B, Pn, self.encoder_dim, self.attention_dim, self.decoder_dim, self.embed_dim,
            _lib.dt_code(self.compute_dtype))))

    def init_hidden_state(self, encoder_out):
        """models/decoder.py:63-67 (external callers: caption.py:94)."""
        b, Pn, E = encoder_out.shape
        Pw = self._cache.get()
        m = Operand.empty((b, E), self.compute_dtype, encoder_out.device)
        _lib.check(_lib.lib().ccx_mean_pixels(ptr(encoder_out.contiguous()), b, Pn, E, ptr(m.hi), m.lo_ptr,
                                              _lib.dt_code(m.dtype), E, _lib.stream_ptr()), "mean_pixels")
        return (_lib.linear(m, Pw["w_init_h"], bias=self.init_h.bias.detach()),
                _lib.linear(m, Pw["w_init_c"], bias=self.init_c.bias.detach()))

    # ---- shared set-up -----------------------------------------------------------------------------------------
    def _setup(self, enc, T):
        """Hoisted work: att1 = encoder_att(enc), h0/c0, and the per-step operand buffer XH[t] = [emb | awe | h]."""
        B, Pn, E = enc.shape
        cd, dev = self.compute_dtype, enc.device
        Pw = self._cache.get()
        L = _lib.lib()
        st = _lib.stream_ptr()
        enc_op = Operand.prepare(enc.view(B * Pn, E), cd)
        att1 = _lib.linear(enc_op, Pw["w_enc_att"], bias=Pw["b_enc_att"])            # (B*P, A) fp32
        K = self.embed_dim + E + self.decoder_dim
        XH = Operand.zeros((T + 1, B, K), cd, dev)
        C_all = torch.empty((T + 1, B, self.decoder_dim), dtype=torch.float32, device=dev)
        m = Operand.empty((B, E), cd, dev)
        _lib.check(L.ccx_mean_pixels(ptr(enc), B, Pn, E, ptr(m.hi), m.lo_ptr, _lib.dt_code(cd), E, st), "mean_pixels")
        hoff = self.embed_dim + E
        h0 = XH.map(lambda t: t[0, :, hoff:])
        if cd == torch.bfloat16:
            _lib.linear(m, Pw["w_init_h"], bias=self.init_h.bias.detach(), out=h0.hi)
        else:
            _lib.linear(m, Pw["w_init_h"], bias=self.init_h.bias.detach(), out=h0, split=True)
        _lib.linear(m, Pw["w_init_c"], bias=self.init_c.bias.detach(), out=C_all[0])
        self._setup_extras = (m, enc_op)        # kept for the backward pass (decoder_train.py)
        return Pw, att1, XH, C_all

    def _step(self, Pw, enc, att1, XH, C_all, HG, G, t, bt, alphas_t, alpha_ld, active, h_all, ha_ld, dm, dm_ld,
              enc_group=1):
        """One decode step for rows [0, bt): models/decoder.py:102-108."""
        L, st = _lib.lib(), _lib.stream_ptr()
        Pn, E, A, D = enc.shape[1], enc.shape[2], self.attention_dim, self.decoder_dim
        hoff = self.embed_dim + E
        code = _lib.dt_code(self.compute_dtype)
        xh_t = XH.map(lambda x: x[t, :bt])
        h_prev = XH.map(lambda x: x[t, :bt, hoff:])
        hg = _lib.linear(h_prev, Pw["w_h"], bias=Pw["b_h"], out=HG[t, :bt])                  # [att2 | gate pre-act]
        awe = XH.map(lambda x: x[t, :bt, self.embed_dim:hoff])
        _lib.check(L.ccx_bahdanau_attention(ptr(att1), ptr(hg), hg.stride(0), ptr(Pw["w_f"]), ptr(Pw["b_f"]),
                                            ptr(enc), ptr(active), ptr(alphas_t), alpha_ld, ptr(awe.hi), awe.lo_ptr,
                                            code, awe.hi.stride(0), bt, Pn, A, E, 1, enc_group, st), "attention")
        gates = _lib.linear(xh_t, Pw["w_lstm"], bias=Pw["b_lstm"], out=G[t, :bt])
        h_next = XH.map(lambda x: x[t + 1, :bt, hoff:])
        _lib.check(L.ccx_lstm_pointwise(ptr(gates), gates.stride(0), ptr(C_all[t]), ptr(C_all[t + 1]),
                                        ptr(h_next.hi), h_next.lo_ptr, h_next.hi.stride(0),
                                        ptr(h_all.hi), h_all.lo_ptr, ha_ld, code, ptr(dm), dm_ld, None, 0, bt, D,
                                        st), "lstm_pointwise")

    def _loop_desc(self, Pw, enc, att1, XH, C_all, HG, G, alphas, H_all, dm, bts):
        """ccx_lstm_tf descriptor over the step buffers (kept alive by the caller)."""
        d = _lib.LstmTF()
        d.XH_hi, d.XH_lo = ptr(XH.hi), XH.lo_ptr
        d.C_all, d.HG, d.G, d.alphas = ptr(C_all), ptr(HG), ptr(G), ptr(alphas)
        d.H_all_hi, d.H_all_lo = ptr(H_all.hi), H_all.lo_ptr
        d.dropmask, d.att1, d.enc = ptr(dm), ptr(att1), ptr(enc)
        d.w_h, d.w_h_lo, d.b_h = ptr(Pw["w_h"].hi), Pw["w_h"].lo_ptr, ptr(Pw["b_h"])
        d.w_f, d.b_f = ptr(Pw["w_f"]), ptr(Pw["b_f"])
        d.w_lstm, d.w_lstm_lo, d.b_lstm = ptr(Pw["w_lstm"].hi), Pw["w_lstm"].lo_ptr, ptr(Pw["b_lstm"])
        arr = _int_array_type(max(len(bts), 1))(*bts)
        d.bts_host = arr
        d._keep = arr                                   # ctypes does not keep the array alive by itself
        d.B, d.T, d.P, d.E = enc.shape[0], len(bts), enc.shape[1], enc.shape[2]
        d.A, d.D, d.Emb = self.attention_dim, self.decoder_dim, self.embed_dim
        d.compute_dtype = _lib.dt_code(self.compute_dtype)
        return d

    def _dropout_mask(self, B, T, dev):
        """nn.Dropout(p) of models/decoder.py:109 as a multiplier tensor (None in eval mode)."""
        if not self.training or self.dropout_p == 0:
            return None
        if self.inject_dropmask is not None:
            return self.inject_dropmask.to(device=dev, dtype=torch.float32).contiguous()
        keep = 1.0 - self.dropout_p
        return torch.empty((B, T, self.decoder_dim), dtype=torch.float32, device=dev).bernoulli_(keep).div_(keep)

    # ---- reference API -----------------------------------------------------------------------------------------
    def forwardWithTeacherForcing(self, encoder_out, encoded_captions, caption_lengths):
        """models/decoder.py:69-113."""
        if torch.is_grad_enabled() and any_requires_grad(self):
            from .decoder_train import lstm_teacher_forcing_with_grad
            return lstm_teacher_forcing_with_grad(self, encoder_out, encoded_captions, caption_lengths)
        return self._tf_forward(encoder_out, encoded_captions, caption_lengths)[:5]

    def _tf_forward(self, encoder_out, encoded_captions, caption_lengths, dropmask_unsorted=None):
        """dropmask_unsorted: (B, >=T, D) dropout multipliers in the CALLER's row order (the free-running training
        path re-uses the masks of its greedy pass); default = a fresh draw / the injected mask, sorted order."""
        _lib.require_cuda(encoder_out, "encoder_out")
        B, E = encoder_out.size(0), encoder_out.size(-1)
        V, D = self.vocab_size, self.decoder_dim
        dev = encoder_out.device
        enc = encoder_out.reshape(B, -1, E)
        Pn = enc.size(1)
        host_lengths = host_copy(caption_lengths).reshape(-1)
        caption_lengths, sort_ind = caption_lengths.squeeze(1).sort(dim=0, descending=True)
        enc = enc[sort_ind].float().contiguous()
        encoded_captions = encoded_captions[sort_ind].contiguous()
        decode_lengths = sorted((host_lengths - 1).tolist(), reverse=True)      # = (sorted lengths - 1).tolist()
        decode_lengths_dev = caption_lengths - 1                                # the same values, on the device
        stash_device_twin(decode_lengths, decode_lengths_dev)
        T = max(decode_lengths)
        if self.fixed_T:
            # static shapes for CUDA-graph replay: buffers span every possible step; the kernels stop at the longest
            # caption of the batch (read from the device) and rows / steps past a caption's length stay zero
            if not self._persist_ok(B, Pn):
                raise ValueError("fixed_T needs the persistent recurrence kernel (bf16, B <= 32, 7x7 features)")
            T = encoded_captions.shape[1] - 1
        L, st, cd = _lib.lib(), _lib.stream_ptr(), self.compute_dtype
        code = _lib.dt_code(cd)

        Pw, att1, XH, C_all = self._setup(enc, T)
        HG = torch.empty((T, B, self.attention_dim + E), dtype=torch.float32, device=dev)
        G = torch.empty((T, B, 4 * D), dtype=torch.float32, device=dev)
        alphas = torch.zeros((B, T, Pn), dtype=torch.float32, device=dev)
        H_all = Operand.zeros((B, T, D), cd, dev)
        if dropmask_unsorted is not None:
            dm = dropmask_unsorted[sort_ind][:, :T].contiguous()
        else:
            dm = self._dropout_mask(B, T, dev)
        # all embeddings at once: XH[t, b, 0:embed] = embedding[caps[b, t]]
        K = XH.hi.shape[2]
        _lib.check(L.ccx_embed_rows(ptr(encoded_captions), encoded_captions.stride(0), 0, ptr(self.embedding.weight),
                                    V, self.embed_dim, None, None, None, 0, 0, ptr(XH.hi), XH.lo_ptr, code,
                                    K, B * K, B, T, st), "embed_rows")
        bts = [sum(l > t for l in decode_lengths) for t in range(T)]
        # the whole time loop in ONE FFI call (csrc/lstm_runner.cu): same kernels/order as self._step per step
        desc = self._loop_desc(Pw, enc, att1, XH, C_all, HG, G, alphas, H_all, dm, bts)
        if self._persist_ok(B, Pn):
            # ONE cooperative kernel for all T steps (csrc/lstm_persist.cu).  Hoisted out of it: the embedding part of
            # the gate GEMM for every (t, b) at once, bias included, columns in the kernel's permuted gate order.
            E_all = _lib.linear(XH.map(lambda x: x[:T].view(T * B, K)[:, :self.embed_dim]), Pw["w2p_emb"],
                                bias=Pw["b_perm"])
            pd = _lib.LstmPersist()
            att1_bf = _lib.cast_bf16(att1)
            counters = torch.empty(4 * (T + 1), dtype=torch.int32, device=dev)
            pd.E_all, pd.w2p, pd.att1_bf, pd.enc_bf = ptr(E_all), ptr(Pw["w2p"].hi), ptr(att1_bf), ptr(
                self._setup_extras[1].hi)
            pd.decode_len, pd.counters = ptr(decode_lengths_dev), ptr(counters)
            awe_all = torch.empty((T, B, E), dtype=torch.float32, device=dev)
            scratch = torch.empty((T + 1) * 32768 + T * 65536 + 2 * 4 * 32 * 1536 * 4, dtype=torch.uint8, device=dev)
            pd.awe_all, pd.scratch = ptr(awe_all), ptr(scratch)
            if self._persist_dbg is not None:          # tools/bench_lstm_persist.py: in-kernel clock64 stamps
                self._persist_dbg = torch.zeros((3, T, 8), dtype=torch.int64, device=dev)
                pd.dbg = ptr(self._persist_dbg)
            _lib.check(L.ccx_lstm_tf_forward_persist(ctypes.byref(desc), ctypes.byref(pd), st), "lstm_tf_forward_persist")
        else:
            awe_all = att1_bf = None
            _lib.check(L.ccx_lstm_tf_forward(ctypes.byref(desc), st), "lstm_tf_forward")
        valid = (torch.arange(T, device=dev).unsqueeze(0) <
                 decode_lengths_dev.unsqueeze(1)).to(torch.float32).reshape(-1).contiguous()
        predictions = torch.empty((B, T, V), dtype=torch.float32, device=dev)
        _lib.linear(H_all.map(lambda x: x.view(B * T, D)), Pw["w_fc"], bias=self.fc.bias.detach(), rowscale=valid,
                    rows_per_group=1, out=predictions.view(B * T, V))
        saved = dict(enc=enc, att1=att1, XH=XH, C_all=C_all, HG=HG, G=G, H_all=H_all, dm=dm, valid=valid, bts=bts,
                     sort_ind=sort_ind, caps=encoded_captions, Pw=Pw, m_op=self._setup_extras[0],
                     enc_op=self._setup_extras[1], alphas=alphas.detach(), T=T, awe_all=awe_all, att1_bf=att1_bf,
                     decode_lengths_dev=decode_lengths_dev)   # an alias: the returned
        # tensor becomes an autograd OUTPUT (grad_fn -> node -> saved -> tensor would be a reference cycle that keeps
        # every activation of the step alive until the cyclic GC runs)
        self._setup_extras = None
        return predictions, encoded_captions, decode_lengths, alphas, sort_ind, saved

    def forwardWithoutTeacherForcing(self, encoder_out, wordMap, maxDecodeLen):
        """models/decoder.py:119-163 (greedy).  With autograd enabled (trainWithoutTeacherForcing,
        trainMultiGPU.py:444-460) the outputs carry gradients to every parameter and to encoder_out."""
        if torch.is_grad_enabled() and (encoder_out.requires_grad or any_requires_grad(self)):
            from .decoder_train import lstm_free_running_with_grad
            return lstm_free_running_with_grad(self, encoder_out, wordMap, maxDecodeLen)
        return self._greedy(encoder_out, wordMap, maxDecodeLen)

    @torch.no_grad()
    def _greedy(self, encoder_out, wordMap, maxDecodeLen, dropmask=False):
        """Finished rows are masked on the device instead of being compacted with nonzero(): their
        predictions/alphas/sequences stay zero exactly as in the reference.  dropmask: False = draw (train mode) or
        none (eval); None / tensor (B, T, D) = use exactly this."""
        _lib.require_cuda(encoder_out, "encoder_out")
        B, E = encoder_out.size(0), encoder_out.size(-1)
        V, D, T = self.vocab_size, self.decoder_dim, int(maxDecodeLen)
        dev = encoder_out.device
        enc = encoder_out.reshape(B, -1, E).float().contiguous()
        Pn = enc.size(1)
        L, st, cd = _lib.lib(), _lib.stream_ptr(), self.compute_dtype
        code = _lib.dt_code(cd)
        Pw, att1, XH, C_all = self._setup(enc, T)
        HG = torch.empty((T, B, self.attention_dim + E), dtype=torch.float32, device=dev)
        G = torch.empty((T, B, 4 * D), dtype=torch.float32, device=dev)
        predictions = torch.zeros((B, T, V), dtype=torch.float32, device=dev)
        alphas = torch.zeros((B, T, Pn), dtype=torch.float32, device=dev)
        sequences = torch.zeros((B, T), dtype=torch.long, device=dev)
        tokens = torch.zeros((B, T + 1), dtype=torch.long, device=dev)
        tokens[:, 0] = wordMap['<start>']
        active = torch.ones(B, dtype=torch.float32, device=dev)
        h_cur = Operand.empty((B, D), cd, dev)
        dm = self._dropout_mask(B, T, dev) if dropmask is False else dropmask
        K = XH.hi.shape[2]
        for t in range(T):
            _lib.check(L.ccx_embed_rows(ptr(tokens), T + 1, t, ptr(self.embedding.weight), V, self.embed_dim, None,
                                        None, None, 0, 0, ptr(XH.hi[t]), None if XH.lo is None else ptr(XH.lo[t]),
                                        code, K, 0, B, 1, st), "embed_rows")
            dm_t = None if dm is None else dm[:, t]
            self._step(Pw, enc, att1, XH, C_all, HG, G, t, B, alphas[:, t], T * Pn, active, h_cur, D, dm_t, T * D)
            p_t = predictions[:, t]
            _lib.linear(h_cur, Pw["w_fc"], bias=self.fc.bias.detach(), rowscale=active, rows_per_group=1, out=p_t)
            _lib.check(L.ccx_greedy_next(ptr(p_t), T * V, B, V, t, T, ptr(sequences), ptr(active),
                                         tokens.data_ptr() + 8 * (t + 1), T + 1, wordMap['<end>'], st), "greedy_next")
            if t % 8 == 7 and not bool(active.any()):   # the reference breaks when every row has finished
                break
        return predictions, alphas, sequences

    def forward(self, teacherForcing, encoder_out, encoded_captions=None, caption_lengths=None, wordMap=None,
                maxDecodeLen=None):
        """models/decoder.py:165-172."""
        if teacherForcing is True:
            return self.forwardWithTeacherForcing(encoder_out, encoded_captions, caption_lengths)
        return self.forwardWithoutTeacherForcing(encoder_out, wordMap, maxDecodeLen)
