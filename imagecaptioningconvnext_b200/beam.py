"""Batched beam search — the reference's ``caption.py:39-155`` (LSTM) and ``caption.py:160-255`` (Transformer)
semantics, per image, for NI images x k beams in one launch sequence (BASELINE.json config 5).

The reference decodes ONE image on the CPU with Python list bookkeeping and, for the Transformer, re-runs the whole
prefix every step.  Here every image keeps its own (scores, seqs, k_remaining) state on the device
(csrc/beam_kernels.cu); the Transformer uses a KV cache that is re-ordered with the beams.  Image file I/O,
PIL resize and matplotlib (caption.py:54-65,258-383) are out of scope: callers pass encoder features.

Returned per image: the completed sequence with the highest cumulative log-probability (first maximum, no length
normalisation — caption.py:151-152), as a list of token ids including <start> and <end>; ``None`` where no beam
completed (the reference raises ValueError there).
"""
import torch

from . import _lib
from ._host import no_gc
from ._lib import Operand, ptr


class _BeamState:
    def __init__(self, NI, k, Tcap, start_tok, dev):
        i32, f32, i64 = torch.int32, torch.float32, torch.long
        self.NI, self.k, self.Tcap = NI, k, Tcap
        self.seqs = [torch.zeros((NI, k, Tcap), dtype=i64, device=dev) for _ in range(2)]
        self.seqs[0][:, :, 0] = start_tok
        self.top = torch.zeros((NI, k), dtype=f32, device=dev)
        self.k_rem = torch.full((NI,), k, dtype=i32, device=dev)
        self.done_seqs = torch.zeros((NI, k, Tcap), dtype=i64, device=dev)
        self.done_scores = torch.full((NI, k), float("-inf"), dtype=f32, device=dev)
        self.done_len = torch.zeros((NI, k), dtype=i32, device=dev)
        self.n_done = torch.zeros((NI,), dtype=i32, device=dev)
        self.src_row = torch.arange(NI * k, dtype=i32, device=dev)
        self.cand_s = torch.empty((NI, k), dtype=f32, device=dev)
        self.cand_p = torch.empty((NI, k), dtype=i32, device=dev)
        self.cand_w = torch.empty((NI, k), dtype=i32, device=dev)
        self.cur = 0
        # attention-map tracking (beam_search_lstm(return_alphas=True)): per step the maps of every row and the
        # parent of every surviving row; a completed sequence records the row it grew from
        self.done_parent = None
        self.alpha_hist = self.parent_hist = None

    def advance(self, logits, V, step, end_tok, next_tok_ptr, ld_next, trace):
        """top-k + bookkeeping for one decode step (`step` = tokens per sequence so far, 1 at the first step)."""
        L, st = _lib.lib(), _lib.stream_ptr()
        _lib.check(L.ccx_beam_topk(ptr(logits), logits.stride(0), self.NI, self.k, V, ptr(self.top), ptr(self.k_rem),
                                   1 if step == 1 else 0, ptr(self.cand_s), ptr(self.cand_p), ptr(self.cand_w), st),
                   "beam_topk")
        if trace is not None:
            trace.append((self.k_rem.clone(), self.cand_s.clone(), self.cand_p.clone(), self.cand_w.clone()))
        _lib.check(L.ccx_beam_update(self.NI, self.k, self.Tcap, step, end_tok, ptr(self.cand_s), ptr(self.cand_p),
                                     ptr(self.cand_w), ptr(self.seqs[self.cur]), ptr(self.seqs[1 - self.cur]),
                                     ptr(self.top), ptr(self.k_rem), ptr(self.done_seqs), ptr(self.done_scores),
                                     ptr(self.done_len), ptr(self.n_done), ptr(self.src_row), next_tok_ptr, ld_next,
                                     ptr(self.done_parent), st), "beam_update")
        if self.parent_hist is not None:
            self.parent_hist[step - 1].copy_(self.src_row)
        self.cur = 1 - self.cur

    def track_alphas(self, steps, Pn):
        dev = self.top.device
        self.done_parent = torch.zeros((self.NI, self.k), dtype=torch.int32, device=dev)
        self.alpha_hist = torch.zeros((steps, self.NI * self.k, Pn), dtype=torch.float32, device=dev)
        self.parent_hist = torch.zeros((steps, self.NI * self.k), dtype=torch.int32, device=dev)

    def results_with_alphas(self, return_all=False):
        """caption.py:151-155 per image: (best completed sequence, its alphas (len(seq), s, s)) or None.  alphas[0]
        is the all-ones map the reference starts seqsAlpha with (caption.py:85); alphas[t] is the attention of the
        step that produced token t, followed back through the beam re-orderings.  return_all: additionally, per
        image, the (seq, alphas) of every completed sequence in completion order."""
        n_done = self.n_done.cpu()
        scores, lens, seqs = self.done_scores.cpu(), self.done_len.cpu(), self.done_seqs.cpu()
        parent, ahist, phist = self.done_parent.cpu(), self.alpha_hist.cpu(), self.parent_hist.cpu()
        Pn = ahist.shape[-1]
        side = int(round(Pn ** 0.5))

        def chain(i, j):
            L = int(lens[i, j])
            row, back = int(parent[i, j]), []
            for s in range(L - 1, 0, -1):            # decode steps s = L-1 .. 1 (1-based)
                back.append(ahist[s - 1, row])
                if s > 1:
                    row = int(phist[s - 2, row])
            maps = [torch.ones(Pn)] + back[::-1]
            return seqs[i, j, :L].tolist(), torch.stack(maps).view(L, side, side)

        best, every = [], []
        for i in range(self.NI):
            n = int(n_done[i])
            every.append([chain(i, j) for j in range(n)] if return_all else None)
            best.append(chain(i, int(torch.argmax(scores[i, :n]))) if n else None)
        return (best, every) if return_all else best

    def all_done(self):
        """Per image: (list of completed sequences in completion order, list of their scores)."""
        n_done = self.n_done.cpu()
        scores, lens, seqs = self.done_scores.cpu(), self.done_len.cpu(), self.done_seqs.cpu()
        return [([seqs[i, j, :int(lens[i, j])].tolist() for j in range(int(n_done[i]))],
                 scores[i, :int(n_done[i])].tolist()) for i in range(self.NI)]

    def results(self):
        """caption.py:151-152 per image: completed sequence with max score (first on ties), or None."""
        n_done = self.n_done.cpu()
        scores, lens, seqs = self.done_scores.cpu(), self.done_len.cpu(), self.done_seqs.cpu()
        out = []
        for i in range(self.NI):
            n = int(n_done[i])
            if n == 0:
                out.append(None)
                continue
            j = int(torch.argmax(scores[i, :n]))   # first maximum, like list.index(max(...))
            out.append(seqs[i, j, :int(lens[i, j])].tolist())
        return out


def _gather(src, dst, src_row, rows):
    """dst[r] = src[src_row[r]] over the leading dimension (both contiguous in the trailing dims)."""
    row_bytes = src[0].numel() * src.element_size()
    _lib.check(_lib.lib().ccx_gather_rows(ptr(src), src.stride(0) * src.element_size(), ptr(dst),
                                          dst.stride(0) * dst.element_size(), ptr(src_row), row_bytes, rows,
                                          _lib.stream_ptr()), "gather_rows")


@torch.no_grad()
def beam_search_lstm(decoder, encoder_out, wordMap, beamSize=3, max_steps=50, trace=None, return_all=False,
                     _state_only=False, return_alphas=False):
    """caption.py:39-155 for a batch of images.  encoder_out (NI, s, s, E) from ``Encoder``; eval-mode decoder.
    return_alphas: per image (seq, alphas) like the reference's ``return seq, alphas`` (caption.py:155)."""
    _lib.require_cuda(encoder_out, "encoder_out")
    NI, E = encoder_out.size(0), encoder_out.size(-1)
    k, V, D, A = int(beamSize), decoder.vocab_size, decoder.decoder_dim, decoder.attention_dim
    dev, cd = encoder_out.device, decoder.compute_dtype
    code = _lib.dt_code(cd)
    L, st = _lib.lib(), _lib.stream_ptr()
    enc = encoder_out.reshape(NI, -1, E).float().contiguous()
    Pn = enc.size(1)
    rows = NI * k
    Tcap = max_steps + 2
    Pw, att1, XH0, C0 = decoder._setup(enc, 0)          # hoisted encoder_att + h0/c0 per IMAGE
    K = decoder.embed_dim + E + D
    hoff = decoder.embed_dim + E
    XH = [Operand.zeros((rows, K), cd, dev) for _ in range(2)]
    Cs = [torch.empty((rows, D), dtype=torch.float32, device=dev) for _ in range(2)]
    expand = (torch.arange(rows, dtype=torch.int32, device=dev) // k).contiguous()   # row -> image
    h0 = XH0.map(lambda x: x[0, :, hoff:])
    es = h0.hi.element_size()
    for src, dst in ((h0.hi, XH[0].hi), (h0.lo, XH[0].lo)):
        if src is not None:
            _lib.check(L.ccx_gather_rows(ptr(src), src.stride(0) * es, dst.data_ptr() + hoff * es, K * es,
                                         ptr(expand), D * es, rows, st), "gather_rows")
    _gather(C0[0], Cs[0], expand, rows)
    bs = _BeamState(NI, k, Tcap, wordMap['<start>'], dev)
    if return_alphas:
        bs.track_alphas(max_steps + 1, Pn)
    tokens = torch.full((rows, 1), wordMap['<start>'], dtype=torch.long, device=dev)
    HG = torch.empty((rows, A + E), dtype=torch.float32, device=dev)
    G = torch.empty((rows, 4 * D), dtype=torch.float32, device=dev)
    h_new = Operand.empty((rows, D), cd, dev)
    c_new = torch.empty((rows, D), dtype=torch.float32, device=dev)
    logits = torch.empty((rows, V), dtype=torch.float32, device=dev)
    cur = 0
    for step in range(1, max_steps + 2):                 # caption.py:147: at most 51 decode steps
        x = XH[cur]
        _lib.check(L.ccx_embed_rows(ptr(tokens), 1, 0, ptr(decoder.embedding.weight), V, decoder.embed_dim, None,
                                    None, None, 0, 0, ptr(x.hi), x.lo_ptr, code, K, 0, rows, 1, st), "embed_rows")
        h_prev = x.map(lambda t: t[:, hoff:])
        _lib.linear(h_prev, Pw["w_h"], bias=Pw["b_h"], out=HG)
        awe = x.map(lambda t: t[:, decoder.embed_dim:hoff])
        a_out = None if bs.alpha_hist is None else bs.alpha_hist[step - 1]
        _lib.check(L.ccx_bahdanau_attention(ptr(att1), ptr(HG), A + E, ptr(Pw["w_f"]), ptr(Pw["b_f"]), ptr(enc), None,
                                            ptr(a_out), Pn, ptr(awe.hi), awe.lo_ptr, code, K, rows, Pn, A, E, 1, k,
                                            st), "attention")
        _lib.linear(x, Pw["w_lstm"], bias=Pw["b_lstm"], out=G)
        _lib.check(L.ccx_lstm_pointwise(ptr(G), 4 * D, ptr(Cs[cur]), ptr(c_new), None, None, 0, ptr(h_new.hi),
                                        h_new.lo_ptr, D, code, None, 0, None, 0, rows, D, st), "lstm_pointwise")
        _lib.linear(h_new, Pw["w_fc"], bias=decoder.fc.bias.detach(), out=logits)
        bs.advance(logits, V, step, wordMap['<end>'], tokens.data_ptr(), 1, trace)
        nxt = XH[1 - cur]
        for src, dst in ((h_new.hi, nxt.hi), (h_new.lo, nxt.lo)):
            if src is not None:
                _lib.check(L.ccx_gather_rows(ptr(src), D * es, dst.data_ptr() + hoff * es, K * es, ptr(bs.src_row),
                                             D * es, rows, st), "gather_rows")
        _gather(c_new, Cs[1 - cur], bs.src_row, rows)
        cur = 1 - cur
        if not _state_only and step % 8 == 0 and not bool(bs.k_rem.any()):   # caption.py:135-136: k == 0 -> break
            break
    if _state_only:
        return bs
    if return_alphas:
        return bs.results_with_alphas(return_all)
    return (bs.results(), bs.all_done()) if return_all else bs.results()


@torch.no_grad()
def beam_search_transformer(decoder, encoder_out, wordMap, beamSize=3, max_decode_len=51, trace=None,
                            return_all=False, _state_only=False):
    """caption.py:160-255 for a batch of images, with a KV cache re-ordered along the surviving beams."""
    _lib.require_cuda(encoder_out, "encoder_out")
    NI = encoder_out.size(0)
    k, V, D = int(beamSize), decoder.vocab_size, decoder.embed_dim
    dev = encoder_out.device
    if max_decode_len > decoder.maxLen:
        raise ValueError("max_decode_len exceeds the positional-encoding table")
    rows = NI * k
    Pw = decoder._cache.get()
    state = decoder.new_decode_state(Pw, encoder_out, rows, max_decode_len, kv_group=k)
    # beams are re-ordered by rewriting a (rows x positions) map of physical cache rows, not by copying the caches:
    # logical row r reads position j from cache row kv_rows[r, j]; its new token's K/V go to cache row r.
    Tm = max_decode_len
    Tpad = (Tm + 3) // 4 * 4
    ident = torch.arange(rows, dtype=torch.int32, device=dev)
    maps = [ident.view(-1, 1).expand(rows, Tpad).contiguous() for _ in range(2)]
    state["kv_rows"] = maps[0]
    tokens = state["tokens"]
    tokens[:, 0] = wordMap['<start>']
    bs = _BeamState(NI, k, max_decode_len + 1, wordMap['<start>'], dev)
    logits = torch.empty((rows, V), dtype=torch.float32, device=dev)
    L, st = _lib.lib(), _lib.stream_ptr()
    cur = 0
    for t in range(max_decode_len):                       # caption.py:249: `if step + 1 >= max_decode_len: break`
        x_op = decoder.decode_step_cached(Pw, state, t, rows)
        _lib.linear(x_op, Pw["fc"], bias=Pw["fc_b"], out=logits)
        bs.advance(logits, V, t + 1, wordMap['<end>'], tokens.data_ptr() + 8 * (t + 1), tokens.stride(0), trace)
        # new_map[r, :] = old_map[parent(r), :]; position t+1 of every row will live in its own cache row
        src, dst = maps[cur], maps[1 - cur]
        _lib.check(L.ccx_gather_rows(ptr(src), Tpad * 4, ptr(dst), Tpad * 4, ptr(bs.src_row), Tpad * 4, rows, st),
                   "gather_rows")
        if t + 1 < Tpad:
            dst[:, t + 1] = ident
        cur = 1 - cur
        state["kv_rows"] = maps[cur]
        if not _state_only and t % 8 == 7 and not bool(bs.k_rem.any()):
            break
    if _state_only:
        return bs
    return (bs.results(), bs.all_done()) if return_all else bs.results()


class CapturedBeamSearch:
    """The whole fixed-shape decode loop (51 steps x ~75 launches for the Transformer) captured ONCE into a CUDA graph
    and replayed per batch: the per-step work is launch-bound at these sizes (640 rows), so removing the host from
    the loop is worth more than any single kernel.  Shapes are fixed at construction (n_images, beamSize, pixels);
    the first call runs eagerly once (warm-up: kernel attributes, weight preparation) and then captures.
    Results equal the eager functions' (same kernels, same order); early exit is replaced by running all steps, which
    cannot change the outcome (finished images have k_rem == 0 and are skipped by the kernels)."""

    def __init__(self, decoder, wordMap, kind, beamSize=3, max_len=None):
        self.decoder, self.wordMap, self.kind, self.k = decoder, wordMap, kind, int(beamSize)
        self.max_len = max_len
        self.graph, self.static_in, self.state = None, None, None

    def _run(self, feats):
        if self.kind == "lstm":
            return beam_search_lstm(self.decoder, feats, self.wordMap, self.k,
                                    max_steps=50 if self.max_len is None else self.max_len, _state_only=True)
        return beam_search_transformer(self.decoder, feats, self.wordMap, self.k,
                                       max_decode_len=51 if self.max_len is None else self.max_len, _state_only=True)

    @torch.no_grad()
    def __call__(self, encoder_out, return_all=False):
        _lib.require_cuda(encoder_out, "encoder_out")
        if self.graph is None or self.static_in.shape != encoder_out.shape:
            self.static_in = encoder_out.detach().clone().float().contiguous()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._run(self.static_in)                       # eager warm-up on the side stream
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with no_gc(), torch.cuda.graph(self.graph):
                self.state = self._run(self.static_in)
        self.static_in.copy_(encoder_out)
        self.graph.replay()
        return (self.state.results(), self.state.all_done()) if return_all else self.state.results()
