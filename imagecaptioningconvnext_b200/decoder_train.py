"""Teacher-forced DecoderWithAttention forward WITH autograd: explicit back-propagation through time on libccx
(reference: models/decoder.py:69-113 under trainMultiGPU.py:363-384).

Per backward step (t = T-1 .. 0): LSTM point-wise backward, ONE dgrad GEMM dgates . [W_ih | W_hh] giving
[d emb_t | d awe_t | d h_{t-1}], the fused attention backward, ONE dgrad GEMM through [decoder_att | f_beta].
All weight gradients are batched over time into single GEMMs after the loop (the reference's autograd runs
~5,100 small kernels for this).
"""
import ctypes

import torch

from . import _lib
from ._host import named_params, params_of
from ._lib import Operand, ptr
from .train_ops import deliver, grads_out, linear_bwd, to_operand, weight_t, zero_grads_like, zeros_many


def _bptt_launch_loop(dec, S, dH_all, dalphas, need_enc, cd, dev):
    """BPTT as the per-step launch loop (csrc/lstm_runner.cu: 5 launches per step): fp32 / 3xTF32 mode, B > 32 or
    feature maps that do not fit the persistent kernel.  Returns the same pieces as the persistent path."""
    L, st = _lib.lib(), _lib.stream_ptr()
    enc, att1, XH, C_all, HG, G, H_all = S["enc"], S["att1"], S["XH"], S["C_all"], S["HG"], S["G"], S["H_all"]
    dm, bts, T, alphas, Pw = S["dm"], S["bts"], S["T"], S["alphas"], S["Pw"]
    B, Pn, E = enc.shape
    D, A, Emb = dec.decoder_dim, dec.attention_dim, dec.embed_dim
    K = Emb + E + D
    # (the C runner takes W^T [K, pad8(N)] operands: explicit transposed copies, whatever weight_t would pick)
    w_lstm_t = to_operand(torch.cat([dec.decode_step.weight_ih.detach(), dec.decode_step.weight_hh.detach()], 1), cd,
                          transpose=True)
    w_h_t = to_operand(torch.cat([dec.attention.decoder_att.weight.detach(), dec.f_beta.weight.detach()], 0), cd,
                       transpose=True)
    (dG_all, dHG_all, dXH_all, d_att1, d_enc, d_wf, dc, dh, dawe_all, dalpha_all) = zeros_many(
        [(T, B, 4 * D), (T, B, A + E), (T, B, K), (B * Pn, A), (B, Pn, E), (A,), (B, D), (B, D), (T, B, E),
         (2, T, B, Pn)], dev)
    if not need_enc:
        d_enc = None
    dal = None if dalphas is None else dalphas.contiguous()
    scratch = Operand.empty((B, 4 * D), cd, dev)
    scratch2 = Operand.empty((B, A + E), cd, dev)
    fwd = dec._loop_desc(Pw, enc, att1, XH, C_all, HG, G, alphas, H_all, dm, bts)
    bd = _lib.LstmTFBwd()
    bd.dH_all, bd.dalphas = ptr(dH_all), ptr(dal)
    bd.dG_all, bd.dHG_all, bd.dXH_all = ptr(dG_all), ptr(dHG_all), ptr(dXH_all)
    bd.dh, bd.dc, bd.d_att1, bd.d_enc, bd.d_wf = ptr(dh), ptr(dc), ptr(d_att1), ptr(d_enc), ptr(d_wf)
    bd.w_lstm_t, bd.w_lstm_t_lo = ptr(w_lstm_t.hi), w_lstm_t.lo_ptr
    bd.w_h_t, bd.w_h_t_lo = ptr(w_h_t.hi), w_h_t.lo_ptr
    bd.scratch_hi, bd.scratch_lo = ptr(scratch.hi), scratch.lo_ptr
    bd.scratch2_hi, bd.scratch2_lo = ptr(scratch2.hi), scratch2.lo_ptr
    # deferred accumulation (see ccx_lstm_tf_bwd): per-step records, summed over time once after the loop
    bd.dawe_all, bd.dalpha_all, bd.de_all = ptr(dawe_all), ptr(dalpha_all[0]), ptr(dalpha_all[1])
    _lib.check(L.ccx_lstm_tf_backward(ctypes.byref(fwd), ctypes.byref(bd), st), "lstm_tf_backward")
    return dG_all, dHG_all, d_att1, d_enc, d_wf, dc, dh, (dXH_all, K, B * K)


class _LstmTF(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dec, holder, encoder_out, caps, lens, *params):
        preds, caps_sorted, decode_lengths, alphas, sort_ind, saved = dec._tf_forward(
            encoder_out, caps, lens, dropmask_unsorted=holder.get("dropmask_unsorted"))
        holder["caps_sorted"], holder["decode_lengths"], holder["sort_ind"] = caps_sorted, decode_lengths, sort_ind
        ctx.dec, ctx.saved = dec, saved
        ctx.enc_needs_grad = encoder_out.requires_grad
        ctx.enc_shape = encoder_out.shape
        return preds, alphas

    @staticmethod
    def backward(ctx, dpred, dalphas):
        dec, S = ctx.dec, ctx.saved
        L, st = _lib.lib(), _lib.stream_ptr()
        cd = dec.compute_dtype
        enc, att1, XH, C_all, HG, G, H_all = S["enc"], S["att1"], S["XH"], S["C_all"], S["HG"], S["G"], S["H_all"]
        dm, bts, T, alphas, Pw = S["dm"], S["bts"], S["T"], S["alphas"], S["Pw"]
        B, Pn, E = enc.shape
        D, A, V, Emb = dec.decoder_dim, dec.attention_dim, dec.vocab_size, dec.embed_dim
        K, hoff = Emb + E + D, Emb + E
        dev = enc.device
        f32 = dict(dtype=torch.float32, device=dev)
        names = [n for n, _ in named_params(dec)]
        params = dict(named_params(dec))
        grads = zero_grads_like(params.items())
        g = lambda n: grads.get(n)
        need_enc = ctx.enc_needs_grad

        # hoisted fc: dH_all[b,t,:] = dpred[b,t,:] . W_fc   (rows past a caption's length carry zero gradient)
        dpred = dpred.contiguous().view(B * T, V)
        dH_all = linear_bwd(dpred, H_all.map(lambda x: x.view(B * T, D)), weight_t(dec.fc.weight, cd, Pw.get("w_fc")), cd,
                            g("fc.weight"), g("fc.bias"))
        persist = S.get("awe_all") is not None
        if persist:
            # BPTT as ONE cooperative kernel (csrc/lstm_persist.cu) + the two deferred attention sums
            (dG_all, dHG_all, d_att1, d_enc, d_wf, dawe_all, de_all) = zeros_many(
                [(T, B, 4 * D), (T, B, A + E), (B * Pn, A), (B, Pn, E), (A,), (T, B, E), (T, B, Pn)], dev)
            dc, dh = torch.empty((B, D), **f32), torch.empty((B, D), **f32)
            dG_bf = torch.zeros((T, B, 4 * D), dtype=torch.bfloat16, device=dev)
            scratch = torch.empty(T * 229376, dtype=torch.uint8, device=dev)
            Xp = torch.empty(2 * 8 * B * E + 2 * 4 * B * D, **f32)
            counters = torch.empty(4 * (T + 1), dtype=torch.int32, device=dev)
            if not need_enc:
                d_enc = None
            dal = None if dalphas is None else dalphas.contiguous()
            fwd = dec._loop_desc(Pw, enc, att1, XH, C_all, HG, G, alphas, H_all, dm, bts)
            bd = _lib.LstmTFBwd()
            bd.dH_all, bd.dalphas = ptr(dH_all), ptr(dal)
            bd.dG_all, bd.dHG_all = ptr(dG_all), ptr(dHG_all)
            bd.dh, bd.dc, bd.d_att1, bd.d_enc, bd.d_wf = ptr(dh), ptr(dc), ptr(d_att1), ptr(d_enc), ptr(d_wf)
            bd.dawe_all, bd.de_all = ptr(dawe_all), ptr(de_all)
            pb = _lib.LstmPersistBwd()
            pb.wx, pb.wht, pb.dG_bf, pb.scratch, pb.Xp = ptr(Pw["wx"].hi), ptr(Pw["wht"].hi), ptr(dG_bf), ptr(
                scratch), ptr(Xp)
            pb.awe_all, pb.att1_bf, pb.enc_bf = ptr(S["awe_all"]), ptr(S["att1_bf"]), ptr(S["enc_op"].hi)
            pb.decode_len, pb.counters = ptr(S["decode_lengths_dev"]), ptr(counters)
            if dec._persist_dbg is not None:
                dec._persist_dbg = torch.zeros((3, T, 8), dtype=torch.int64, device=dev)
                pb.dbg = ptr(dec._persist_dbg)
            _lib.check(L.ccx_lstm_tf_backward_persist(ctypes.byref(fwd), ctypes.byref(bd), ctypes.byref(pb), st),
                       "lstm_tf_backward_persist")
            # d emb_t = dgates_t . W_ih[:, :Emb], all steps in one GEMM
            d_emb = _lib.linear(Operand(dG_bf.view(T * B, 4 * D), None, torch.bfloat16), Pw["w_emb_t"])
            emb_grad_src = (d_emb, Emb, B * Emb)
        else:
            dG_all, dHG_all, d_att1, d_enc, d_wf, dc, dh, emb_grad_src = _bptt_launch_loop(
                dec, S, dH_all, dalphas, need_enc, cd, dev)
        # ---- weight gradients, batched over time -------------------------------------------------------------
        gw_lstm, gb_lstm, gw_h, gb_h = zeros_many([(4 * D, K), (4 * D,), (A + E, D), (A + E,)], dev)
        TB = T * B
        x_all = XH.map(lambda x: x[:T].view(TB, K))
        if g("decode_step.weight_ih") is not None:
            gw, gb = gw_lstm, gb_lstm
            linear_bwd(dG_all.view(TB, 4 * D), x_all, None, cd, gw, gb, need_dx=False)
            ds = dec.decode_step
            deliver(grads, "decode_step.weight_ih", ds.weight_ih, gw[:, :hoff])
            deliver(grads, "decode_step.weight_hh", ds.weight_hh, gw[:, hoff:])
            deliver(grads, "decode_step.bias_ih", ds.bias_ih, gb)
            deliver(grads, "decode_step.bias_hh", ds.bias_hh, gb.clone() if grads.get("decode_step.bias_ih") is gb else gb)
        if g("f_beta.weight") is not None:
            gw, gb = gw_h, gb_h
            h_prev_all = XH.map(lambda x: x[:T].view(TB, K)[:, hoff:])
            linear_bwd(dHG_all.view(TB, A + E), h_prev_all, None, cd, gw, gb, need_dx=False)
            deliver(grads, "attention.decoder_att.weight", dec.attention.decoder_att.weight, gw[:A])
            deliver(grads, "f_beta.weight", dec.f_beta.weight, gw[A:])
            deliver(grads, "attention.decoder_att.bias", dec.attention.decoder_att.bias, gb[:A])
            deliver(grads, "f_beta.bias", dec.f_beta.bias, gb[A:])
        if g("attention.full_att.weight") is not None:
            deliver(grads, "attention.full_att.weight", dec.attention.full_att.weight, d_wf.view(1, A))
            # d/d b_f of softmax(e + b_f) is identically zero (softmax is shift invariant)
        ge = g("embedding.weight")
        if ge is not None:
            caps = S["caps"]
            dx, sb, stt = emb_grad_src
            _lib.check(L.ccx_embedding_bwd(ptr(caps), caps.stride(0), 0, ptr(dx), sb, stt, None, ptr(ge), V,
                                           Emb, B, T, st), "embedding_bwd")
        # ---- initial state and hoisted encoder_att -------------------------------------------------------------
        dmean = linear_bwd(dh, S["m_op"], weight_t(dec.init_h.weight, cd, Pw.get("w_init_h")), cd, g("init_h.weight"), g("init_h.bias"),
                           need_dx=need_enc)
        dmean = linear_bwd(dc, S["m_op"], weight_t(dec.init_c.weight, cd, Pw.get("w_init_c")), cd, g("init_c.weight"), g("init_c.bias"),
                           need_dx=need_enc, dx_residual=dmean)
        if need_enc:
            _lib.check(L.ccx_bcast_add_rows(ptr(d_enc), ptr(dmean), 1.0 / Pn, B, Pn, E, st), "bcast_add_rows")
        d_enc_flat = linear_bwd(d_att1, S["enc_op"], weight_t(dec.attention.encoder_att.weight, cd, Pw.get("w_enc_att")), cd,
                                g("attention.encoder_att.weight"), g("attention.encoder_att.bias"), need_dx=need_enc,
                                dx_residual=None if d_enc is None else d_enc.view(B * Pn, E))
        d_encoder_out = None
        if need_enc:
            inv = torch.empty_like(S["sort_ind"])
            inv[S["sort_ind"]] = torch.arange(B, device=dev)          # un-sort (index plumbing)
            inv32 = inv.to(torch.int32)
            d_unsorted = torch.empty_like(d_enc_flat)
            _lib.check(L.ccx_gather_rows(ptr(d_enc_flat), Pn * E * 4, ptr(d_unsorted), Pn * E * 4, ptr(inv32),
                                         Pn * E * 4, B, st), "gather_rows")
            d_encoder_out = d_unsorted.view(ctx.enc_shape)
        return (None, None, d_encoder_out, None, None) + grads_out(grads, named_params(dec))


def lstm_teacher_forcing_with_grad(dec, encoder_out, encoded_captions, caption_lengths, dropmask_unsorted=None):
    holder = {"dropmask_unsorted": dropmask_unsorted}
    params = params_of(dec)
    preds, alphas = _LstmTF.apply(dec, holder, encoder_out, encoded_captions, caption_lengths, *params)
    return preds, holder["caps_sorted"], holder["decode_lengths"], alphas, holder["sort_ind"]


def generated_captions(sequences, start_token, end_token, max_len):
    """Greedy output (B, T) -> the caption tensor a teacher-forced pass would be fed to reproduce the same steps:
    row = [<start>, generated ids ...] (B, T+1) and caption_lengths (B, 1) = decode length + 1, where the decode
    length is the position of the first <end> + 1, or T (utils/utils.py:270-276)."""
    B, T = sequences.shape
    is_end = sequences == end_token
    first_end = torch.where(is_end.any(dim=1), is_end.to(torch.int64).argmax(dim=1) + 1,
                            torch.full((B,), T, dtype=torch.int64, device=sequences.device))
    start = torch.full((B, 1), start_token, dtype=torch.int64, device=sequences.device)
    return torch.cat([start, sequences], dim=1), (first_end + 1).unsqueeze(1)


def lstm_free_running_with_grad(dec, encoder_out, wordMap, maxDecodeLen):
    """Free-running TRAINING forward (models/decoder.py:119-163 with autograd, trainMultiGPU.py:444-460).

    No gradient flows through argmax, and rows never interact, so the reference's 51-step autograd graph equals:
    (1) the greedy pass (inference kernels, no graph kept) that fixes the generated ids and each row's finish step;
    (2) a teacher-forced pass over [<start>, generated ids] with the SAME dropout masks, whose explicit BPTT
    (_LstmTF) then yields exactly the gradients of the reference loop — h/c recurrences, attention, the embedding
    rows of the sampled ids, encoder_out.  Outputs are returned in the caller's row order, zero past each row's
    finish step like the reference's."""
    T = int(maxDecodeLen)
    B = encoder_out.size(0)
    dev = encoder_out.device
    dm = dec._dropout_mask(B, T, dev)                  # one draw, shared by both passes (None in eval mode)
    _, _, sequences = dec._greedy(encoder_out.detach(), wordMap, T, dropmask=dm)
    caps, lens = generated_captions(sequences, wordMap['<start>'], wordMap['<end>'], T)
    preds_s, _, decode_lengths, alphas_s, sort_ind = lstm_teacher_forcing_with_grad(dec, encoder_out, caps, lens,
                                                                                    dropmask_unsorted=dm)
    inv = torch.empty_like(sort_ind)
    inv[sort_ind] = torch.arange(B, device=dev)
    pad_t = T - preds_s.size(1)
    preds, alphas = preds_s[inv], alphas_s[inv]
    if pad_t:
        preds = torch.nn.functional.pad(preds, (0, 0, 0, pad_t))
        alphas = torch.nn.functional.pad(alphas, (0, 0, 0, pad_t))
    return preds, alphas, sequences
