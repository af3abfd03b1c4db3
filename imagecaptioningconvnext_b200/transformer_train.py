"""Teacher-forced TransformerDecoder forward WITH autograd (train mode): one ``torch.autograd.Function`` whose forward
and backward are libccx launches only (reference: models/transformerDecoder.py:88-108 under trainMultiGPU.py:371-384).

Dropout (p = 0.5 everywhere the reference has it: embedding, both attention-probability dropouts, dropout1/2/3 and
the FFN hidden dropout — torch/nn/modules/transformer.py:1158-1199) uses multiplier tensors drawn with
torch.bernoulli (or injected by tests) and applied inside the GEMM epilogues / attention kernel.
"""
import math
import os
import weakref

import torch

from . import _lib
from ._host import host_copy, named_params, no_gc, params_of, stash_device_twin
from ._lib import Operand, ptr
from .train_ops import grads_out, linear_bwd, ln_bwd, to_operand, weight_t, zero_grads_like

_SITES = ("sa_p", "d1", "ca_p", "d2", "ff", "d3")


def _masks(dec, B, T, Pn, dev):
    """Dropout multipliers for one forward (None entries in eval mode)."""
    D, H, Dff, L = dec.embed_dim, dec.num_heads, dec.decoder_dim, dec.num_layers
    if not dec.training or dec.dropout_p == 0:
        return None
    if dec.inject_dropout is not None:
        return {k: v.to(device=dev, dtype=torch.float32).contiguous() for k, v in dec.inject_dropout.items()}
    keep = 1.0 - dec.dropout_p
    if os.environ.get("CCX_DROPOUT_FLAT", "1") == "0":          # A/B switch: one draw per site (3 launches each)
        def draw(*shape):
            return (torch.bernoulli(torch.full(shape, keep, device=dev)) / keep).contiguous()
        m = {"emb": draw(B * T, D)}
        for l in range(L):
            m[(l, "sa_p")], m[(l, "d1")], m[(l, "ca_p")] = draw(B, H, T, T), draw(B * T, D), draw(B, H, T, Pn)
            m[(l, "d2")], m[(l, "ff")], m[(l, "d3")] = draw(B * T, D), draw(B * T, Dff), draw(B * T, D)
        return m
    # all 6 L + 1 masks of a forward are views of ONE buffer drawn by one bernoulli_ / one div_ (a draw per site was
    # 3 launches x 37 sites per train step); every view starts on a 256-byte boundary (the kernels read them as float4)
    shapes = [("emb", (B * T, D))]
    for l in range(L):
        shapes += [((l, "sa_p"), (B, H, T, T)), ((l, "d1"), (B * T, D)), ((l, "ca_p"), (B, H, T, Pn)),
                   ((l, "d2"), (B * T, D)), ((l, "ff"), (B * T, Dff)), ((l, "d3"), (B * T, D))]
    offs, total = [], 0
    for _, sh in shapes:
        offs.append(total)
        total += (torch.Size(sh).numel() + 63) // 64 * 64
    flat = torch.empty(total, dtype=torch.float32, device=dev).bernoulli_(keep).div_(keep)
    return {k: flat[o:o + torch.Size(sh).numel()].view(sh) for (k, sh), o in zip(shapes, offs)}


class _Payload:
    """What the forward keeps for the backward (plain attribute bag; also used as the CUDA-graph static state)."""


def _forward_body(dec, encoder_out, caps, kpm):
    """Train-mode forward on libccx; returns (predictions, payload)."""
    if True:
        ctx = _Payload()
        L, st = _lib.lib(), _lib.stream_ptr()
        cd = dec.compute_dtype
        code = _lib.dt_code(cd)
        B = encoder_out.size(0)
        D, V, H, T = dec.embed_dim, dec.vocab_size, dec.num_heads, caps.size(1)
        dev = encoder_out.device
        Pw = dec._cache.get()
        E = encoder_out.size(-1)
        enc = encoder_out.reshape(B, -1, E).float().contiguous()
        Pn = enc.size(1)
        enc_op = Operand.prepare(enc.view(B * Pn, E), cd)
        mem = dec._linear_op(enc_op, Pw["proj"], Pw["proj_b"]) if Pw["proj"] is not None else enc_op
        masks = _masks(dec, B, T, Pn, dev)
        mk = (lambda k: None) if masks is None else (lambda k: masks.get(k))
        scale = 1.0 / math.sqrt(D // H)

        x_plain = torch.empty((B * T, D), dtype=torch.float32, device=dev)
        x_op = Operand.empty((B * T, D), cd, dev)
        _lib.check(L.ccx_embed_rows(ptr(caps), T, 0, ptr(dec.embedding.weight), V, D, ptr(Pw["pe"]), ptr(mk("emb")),
                                    ptr(x_plain), T * D, D, ptr(x_op.hi), x_op.lo_ptr, code, T * D, D, B, T, st),
                   "embed_rows")
        saved = []
        alphas = torch.empty((H, B, Pn), dtype=torch.float32, device=dev) if dec._want_alphas else None
        for li, lw in enumerate(Pw["layers"]):
            S = {"x0_op": x_op}
            qkv = _lib.linear(x_op, lw["sa_in"], bias=lw["sa_in_b"])
            probs1 = torch.empty((B, H, T, T), dtype=torch.float32, device=dev)
            ctx1 = dec._mha(ptr(qkv), T * 3 * D, 3 * D, qkv.data_ptr() + 4 * D, T * 3 * D, 3 * D,
                            qkv.data_ptr() + 8 * D, B, T, T, 1, 0, kpm, mk((li, "sa_p")), 1, dev, probs_out=probs1)
            y1 = _lib.linear(ctx1, lw["sa_out"], bias=lw["sa_out_b"], residual=x_plain, emask=mk((li, "d1")))
            x1_plain, x1_op = dec._ln(y1, lw["n"][0], B * T)
            q2 = _lib.linear(x1_op, lw["ca_q"], bias=lw["ca_q_b"])
            kv = _lib.linear(mem, lw["ca_kv"], bias=lw["ca_kv_b"])
            probs2 = torch.empty((B, H, T, Pn), dtype=torch.float32, device=dev)
            ctx2 = dec._mha(ptr(q2), T * D, D, ptr(kv), Pn * 2 * D, 2 * D, kv.data_ptr() + 4 * D, B, T, Pn, 0, 0,
                            None, mk((li, "ca_p")), 1, dev, probs_out=probs2)
            if alphas is not None:          # what need_weights=True returns in train mode: dropped-out weights
                dec._head_mean(probs2, mk((li, "ca_p")), ptr(alphas), Pn, B * Pn, B, T, Pn, li, over="positions")
            y2 = _lib.linear(ctx2, lw["ca_out"], bias=lw["ca_out_b"], residual=x1_plain, emask=mk((li, "d2")))
            x2_plain, x2_op = dec._ln(y2, lw["n"][1], B * T)
            h_plain = _lib.linear(x2_op, lw["l1"], bias=lw["l1_b"], act=_lib.ACT_RELU, emask=mk((li, "ff")))
            h_op = to_operand(h_plain, cd)
            y3 = _lib.linear(h_op, lw["l2"], bias=lw["l2_b"], residual=x2_plain, emask=mk((li, "d3")))
            x_plain, x_op = dec._ln(y3, lw["n"][2], B * T)
            S.update(qkv=qkv, probs1=probs1, ctx1=ctx1, y1=y1, x1_op=x1_op, q2=q2, kv=kv, probs2=probs2, ctx2=ctx2,
                     y2=y2, x2_op=x2_op, h_plain=h_plain, h_op=h_op, y3=y3)
            saved.append(S)
        predictions = torch.empty((B, T, V), dtype=torch.float32, device=dev)
        _lib.linear(x_op, Pw["fc"], bias=Pw["fc_b"], out=predictions.view(B * T, V))
        ctx.dec, ctx.saved, ctx.masks, ctx.kpm = dec, saved, masks, kpm
        ctx.x_last_op, ctx.mem, ctx.enc_op, ctx.caps = x_op, mem, enc_op, caps
        ctx.dims = (B, T, Pn, E)
        ctx.enc_shape = encoder_out.shape
        ctx.alphas = alphas
        return predictions, ctx


def _backward_body(ctx, dpred, enc_needs_grad):
    """Explicit backward on libccx; returns (d encoder_out or None, {parameter name: gradient})."""
    if True:
        dec = ctx.dec
        L, st = _lib.lib(), _lib.stream_ptr()
        cd = dec.compute_dtype
        B, T, Pn, E = ctx.dims
        D, V, H, Dff = dec.embed_dim, dec.vocab_size, dec.num_heads, dec.decoder_dim
        dev = dpred.device
        masks = ctx.masks
        mk = (lambda k: None) if masks is None else (lambda k: masks.get(k))
        mode = 0 if masks is None else 1
        keep_scale = 1.0 if masks is None else 1.0 / (1.0 - dec.dropout_p)
        scale = 1.0 / math.sqrt(D // H)
        # bf16 compute mode at the decoder's shapes (<= 64 tokens / pixels, heads of 64): attention backward on the
        # tensor cores (csrc/mha_tc.cu); otherwise the fp32 SIMT kernel
        mha_bwd_fn = (L.ccx_mha_bwd_tc if (cd == torch.bfloat16 and T <= 64 and Pn <= 64 and D // H == 64 and
                                           os.environ.get("CCX_MHA_TC", "1") != "0") else L.ccx_mha_bwd)
        names = [n for n, _ in named_params(dec)]
        params = dict(named_params(dec))
        grads = zero_grads_like(params.items())
        g = lambda n: grads.get(n)
        M = B * T

        dpred = dpred.contiguous().view(M, V)
        dx = linear_bwd(dpred, ctx.x_last_op, weight_t(dec.fc_out.weight, cd), cd, g("fc_out.weight"),
                        g("fc_out.bias"))
        dmem = torch.zeros((B * Pn, D), dtype=torch.float32, device=dev)
        for li in reversed(range(dec.num_layers)):
            S, lyr = ctx.saved[li], dec.transformer_decoder.layers[li]
            pre = f"transformer_decoder.layers.{li}."
            # LN3 / FFN
            dy3 = ln_bwd(dx, S["y3"], lyr.norm3.weight.detach(), g(pre + "norm3.weight"), g(pre + "norm3.bias"), 1e-5)
            dh = linear_bwd(dy3, S["h_op"], weight_t(lyr.linear2.weight, cd), cd, g(pre + "linear2.weight"),
                            g(pre + "linear2.bias"), mul=mk((li, "d3")), mul_mode=mode)
            dx2 = linear_bwd(dh, S["x2_op"], weight_t(lyr.linear1.weight, cd), cd, g(pre + "linear1.weight"),
                             g(pre + "linear1.bias"), dx_residual=dy3, mul=S["h_plain"], mul_mode=2,
                             mul_scale=keep_scale)
            # LN2 / cross attention
            dy2 = ln_bwd(dx2, S["y2"], lyr.norm2.weight.detach(), g(pre + "norm2.weight"), g(pre + "norm2.bias"), 1e-5)
            ca = lyr.multihead_attn
            dctx2 = linear_bwd(dy2, S["ctx2"], weight_t(ca.out_proj.weight, cd), cd, g(pre + "multihead_attn.out_proj.weight"),
                               g(pre + "multihead_attn.out_proj.bias"), mul=mk((li, "d2")), mul_mode=mode)
            dq2 = torch.empty((M, D), dtype=torch.float32, device=dev)
            dkv = torch.empty((B * Pn, 2 * D), dtype=torch.float32, device=dev)
            q2, kv = S["q2"], S["kv"]
            _lib.check(mha_bwd_fn(ptr(q2), T * D, D, ptr(kv), Pn * 2 * D, 2 * D, kv.data_ptr() + 4 * D, Pn * 2 * D,
                                     2 * D, ptr(dctx2), T * D, D, ptr(S["probs2"]), ptr(mk((li, "ca_p"))),
                                     ptr(dq2), T * D, D, ptr(dkv), Pn * 2 * D, 2 * D, dkv.data_ptr() + 4 * D,
                                     Pn * 2 * D, 2 * D, B, H, T, Pn, D // H, scale, st), "mha_bwd")
            gw, gb = g(pre + "multihead_attn.in_proj_weight"), g(pre + "multihead_attn.in_proj_bias")
            w_in = ca.in_proj_weight.detach()
            dx1 = linear_bwd(dq2, S["x1_op"], weight_t(w_in[:D], cd), cd, None if gw is None else gw[:D],
                             None if gb is None else gb[:D], dx_residual=dy2)
            dmem = linear_bwd(dkv, ctx.mem, weight_t(w_in[D:], cd), cd, None if gw is None else gw[D:],
                              None if gb is None else gb[D:], dx_residual=dmem)
            # LN1 / self attention
            dy1 = ln_bwd(dx1, S["y1"], lyr.norm1.weight.detach(), g(pre + "norm1.weight"), g(pre + "norm1.bias"), 1e-5)
            sa = lyr.self_attn
            dctx1 = linear_bwd(dy1, S["ctx1"], weight_t(sa.out_proj.weight, cd), cd, g(pre + "self_attn.out_proj.weight"),
                               g(pre + "self_attn.out_proj.bias"), mul=mk((li, "d1")), mul_mode=mode)
            qkv = S["qkv"]
            dqkv = torch.empty((M, 3 * D), dtype=torch.float32, device=dev)
            _lib.check(mha_bwd_fn(ptr(qkv), T * 3 * D, 3 * D, qkv.data_ptr() + 4 * D, T * 3 * D, 3 * D,
                                     qkv.data_ptr() + 8 * D, T * 3 * D, 3 * D, ptr(dctx1), T * D, D,
                                     ptr(S["probs1"]), ptr(mk((li, "sa_p"))), ptr(dqkv), T * 3 * D, 3 * D,
                                     dqkv.data_ptr() + 4 * D, T * 3 * D, 3 * D, dqkv.data_ptr() + 8 * D, T * 3 * D,
                                     3 * D, B, H, T, T, D // H, scale, st), "mha_bwd")
            dx = linear_bwd(dqkv, S["x0_op"], weight_t(sa.in_proj_weight, cd), cd, g(pre + "self_attn.in_proj_weight"),
                            g(pre + "self_attn.in_proj_bias"), dx_residual=dy1)
        # embedding (dense gradient, times the embedding-dropout multiplier; the PE add has no parameters)
        ge = g("embedding.weight")
        if ge is not None:
            _lib.check(L.ccx_embedding_bwd(ptr(ctx.caps), T, 0, ptr(dx), T * D, D, ptr(mk("emb")), ptr(ge), V, D, B, T,
                                           st), "embedding_bwd")
        # encoder_proj
        denc = None
        if isinstance(dec.encoder_proj, torch.nn.Identity):
            denc = dmem if enc_needs_grad else None
        else:
            denc = linear_bwd(dmem, ctx.enc_op, weight_t(dec.encoder_proj.weight, cd), cd, g("encoder_proj.weight"),
                              g("encoder_proj.bias"), need_dx=enc_needs_grad)
        if denc is not None:
            denc = denc.view(ctx.enc_shape)
        return denc, grads


class _GraphState:
    """One captured (forward, backward) pair for a fixed shape signature: static input buffers, the payload whose
    tensors live in the graphs' private memory pool, static gradient outputs."""

    def __init__(self):
        self.calls = 0
        self.fwd = self.bwd = None
        self.owner = None          # weakref to the token of the forward whose backward has not run yet

    @property
    def pending(self):
        """True while a graphed forward still waits for its backward (its autograd node is alive): the static
        buffers belong to it, so another forward has to run eagerly."""
        return self.owner is not None and self.owner() is not None


class _Token:
    pass


def _graph_key(dec, encoder_out, caps, kpm):
    return (tuple(encoder_out.shape), tuple(caps.shape), kpm is not None, dec.training, encoder_out.requires_grad,
            dec._want_alphas,
            dec.compute_dtype, encoder_out.device, dec._cache.storage_key(),
            tuple(p.requires_grad for p in params_of(dec)))


class _TransformerTF(torch.autograd.Function):
    """Eager path, or — with ``dec.enable_cuda_graph()`` — the forward body and the backward body each replayed as
    ONE CUDA graph (the step is ~580 tiny launches; replay removes the launch gaps and the host from the loop).
    The first two calls of a shape signature run eagerly (warm-up), the third captures."""

    @staticmethod
    def forward(ctx, dec, encoder_out, caps, kpm, use_graph, *params):
        ctx.enc_needs_grad = encoder_out.requires_grad
        ctx.names = [n for n, _ in named_params(dec)]
        graphs = getattr(dec, "_train_graphs", None)
        st = None
        if graphs is not None and use_graph:
            dec._cache.get()                                   # weight copies refreshed in place, outside the graph
            key = _graph_key(dec, encoder_out, caps, kpm)
            st = graphs.get(key)
            if st is None:
                st = graphs[key] = _GraphState()
            st.calls += 1
            if st.pending or st.calls <= 2:
                st = None                                      # buffers busy (two forwards in flight) or warming up
        ctx.gstate = st
        if st is None:
            predictions, ctx.payload = _forward_body(dec, encoder_out, caps, kpm)
            dec._last_alphas = ctx.payload.alphas
            return predictions
        if st.fwd is None:
            st.enc_in, st.caps_in = encoder_out.detach().clone(), caps.clone()
            st.kpm_in = None if kpm is None else kpm.clone()
            torch.cuda.synchronize()
            st.fwd = torch.cuda.CUDAGraph()
            with no_gc(), torch.cuda.graph(st.fwd):
                st.predictions, st.payload = _forward_body(dec, st.enc_in, st.caps_in, st.kpm_in)
            st.pool = st.fwd.pool()
        st.enc_in.copy_(encoder_out.detach())
        st.caps_in.copy_(caps)
        if kpm is not None:
            st.kpm_in.copy_(kpm)
        st.fwd.replay()
        ctx.token = _Token()
        st.owner = weakref.ref(ctx.token)
        ctx.payload = st.payload
        dec._last_alphas = st.payload.alphas
        return st.predictions.detach()          # fresh alias: autograd attaches this call's node to it

    @staticmethod
    def backward(ctx, dpred):
        st = ctx.gstate
        if st is None:
            denc, grads = _backward_body(ctx.payload, dpred, ctx.enc_needs_grad)
            return (None, denc, None, None, None) + grads_out(grads, named_params(ctx.payload.dec))
        if st.owner is None or st.owner() is not ctx.token:
            raise RuntimeError("graphed TransformerDecoder backward ran twice, or after its static buffers were reused")
        if st.bwd is None:
            st.dpred_in = dpred.detach().contiguous().clone()
            torch.cuda.synchronize()
            st.bwd = torch.cuda.CUDAGraph()
            with no_gc(), torch.cuda.graph(st.bwd, pool=st.pool):
                st.denc, st.grads = _backward_body(st.payload, st.dpred_in, ctx.enc_needs_grad)
        st.dpred_in.copy_(dpred)
        st.bwd.replay()
        st.owner = None
        return (None, st.denc, None, None, None) + grads_out(st.grads, named_params(ctx.payload.dec))


def transformer_teacher_forcing_with_grad(dec, encoder_out, encoded_captions, caption_lengths, tgt_key_padding_mask):
    _lib.require_cuda(encoder_out, "encoder_out")
    decode_lengths = (host_copy(caption_lengths).reshape(-1) - 1).tolist()
    if caption_lengths.is_cuda:
        stash_device_twin(decode_lengths, caption_lengths.reshape(-1) - 1)
    caps = encoded_captions.contiguous()
    kpm = None if tgt_key_padding_mask is None else tgt_key_padding_mask.to(torch.uint8).contiguous()
    params = params_of(dec)
    use_graph = torch.is_grad_enabled() and (encoder_out.requires_grad or any(p.requires_grad for p in params))
    predictions = _TransformerTF.apply(dec, encoder_out, caps, kpm, use_graph, *params)
    return predictions, encoded_captions, decode_lengths


def transformer_free_running_with_grad(dec, encoder_out, wordMap, maxDecodeLen):
    """Free-running TRAINING forward (models/transformerDecoder.py:110-160 with autograd).

    The reference re-runs the whole prefix at every step and keeps all 51 graphs (O(T^2) forward and backward).

    Dropout-free modules (eval mode, dropout = 0): under the causal mask the step-t logits depend only on tokens 0..t,
    a finished row is never computed again and an active row's prefix holds no <pad>, so — no gradient passing through
    argmax — the same loss gradient comes from (1) the KV-cached greedy pass that fixes the generated ids and (2) ONE
    causal teacher-forced pass over [<start>, generated ids] with autograd, its logits zeroed past each row's finish
    step (parity-tested against the reference's gradients).

    Live dropout (decoder.train(), trainMultiGPU.py:425-426): the reference draws a FRESH dropout realisation at every
    step's prefix recomputation (models/transformerDecoder.py:129-130 and the layers' own dropouts), which both picks
    the generated ids and defines the gradient; a KV cache cannot reproduce that (the keys / values of earlier
    positions change with every realisation).  ``_free_running_exact`` therefore does what the reference does — one
    differentiable causal pass over the prefix per step, 51 autograd nodes — on the libccx kernels.
    ``dec.free_running_fast = True`` opts into the cheap estimator instead (ids generated dropout-free, one mask set
    for the single differentiable pass): same expectation, ~12x faster, not the reference's sampling."""
    if dec.training and dec.dropout_p > 0 and not getattr(dec, "free_running_fast", False):
        return _free_running_exact(dec, encoder_out, wordMap, maxDecodeLen)
    from .decoder_train import generated_captions
    T = int(maxDecodeLen)
    _, sequences = dec._greedy(encoder_out.detach(), wordMap, T, dropout_free=True)
    greedy_alphas = dec._last_alphas          # attention-map variant: maps of the generation pass (zero once finished)
    caps, lens = generated_captions(sequences, wordMap['<start>'], wordMap['<end>'], T)
    params = params_of(dec)
    preds = _TransformerTF.apply(dec, encoder_out, caps[:, :T].contiguous(), None, True, *params)
    dec._last_alphas = greedy_alphas
    valid = torch.arange(T, device=preds.device).unsqueeze(0) < (lens - 1)
    return preds * valid.unsqueeze(-1).to(preds.dtype), sequences


def _free_running_exact(dec, encoder_out, wordMap, maxDecodeLen):
    """models/transformerDecoder.py:124-158 step by step under autograd: at step t the whole prefix (t + 1 tokens, the
    finished rows padded with <pad>) goes through the decoder with that step's own dropout masks; the last position's
    logits give the step's predictions (kept in the autograd graph) and, by argmax, the next token.
    ``dec.inject_dropout_steps`` (tests): list over t of mask dicts in ``_masks``' format for T = t + 1."""
    T = int(maxDecodeLen)
    B = encoder_out.size(0)
    dev = encoder_out.device
    start, end, pad = wordMap['<start>'], wordMap['<end>'], wordMap['<pad>']
    params = params_of(dec)
    inputs = torch.full((B, T), pad, dtype=torch.long, device=dev)
    inputs[:, 0] = start
    finished = torch.zeros(B, dtype=torch.bool, device=dev)
    sequences = torch.zeros((B, T), dtype=torch.long, device=dev)
    steps = getattr(dec, "inject_dropout_steps", None)
    saved_inject = dec.inject_dropout
    preds = []
    try:
        for t in range(T):
            if steps is not None:
                dec.inject_dropout = steps[t]
            active = ~finished
            p_all = _TransformerTF.apply(dec, encoder_out, inputs[:, :t + 1].contiguous(), None, False, *params)
            p_t = p_all[:, t] * active.unsqueeze(1).to(p_all.dtype)      # finished rows stay zero (:119-121)
            ids = p_t.detach().argmax(dim=1)
            sequences[:, t] = torch.where(active, ids, sequences[:, t])
            finished = finished | (active & (ids == end))
            if t + 1 < T:
                inputs[:, t + 1] = torch.where(active, ids, torch.full_like(ids, pad))   # :150-158
            preds.append(p_t)
            if t % 8 == 7 and bool(finished.all()):                      # the reference breaks when all rows finished
                break
    finally:
        dec.inject_dropout = saved_inject
    out = torch.stack(preds, dim=1)
    if out.size(1) < T:
        out = torch.nn.functional.pad(out, (0, 0, 0, T - out.size(1)))
    dec._last_alphas = None
    return out, sequences


def enable_cuda_graph(dec, enabled=True):
    """Training option: replay the teacher-forced forward and backward as CUDA graphs (see _TransformerTF).
    The returned predictions and the gradients handed to autograd are then STATIC buffers, overwritten by the next
    graphed step of the same shape: consume them (loss, optimizer step, all-reduce) before the next forward, as a
    normal training loop does.  A second forward issued before the first one's backward runs eagerly."""
    dec._train_graphs = {} if enabled else None
    return dec
