"""One teacher-forced train step — the body of the reference's hot loop (trainMultiGPU.py:357-394 = train.py:257-291):
encoder -> decoder(teacherForcing=True) -> packed cross-entropy (+ doubly-stochastic attention regulariser for the
LSTM decoder) -> zero_grad -> backward -> clip_gradient(+-5) -> Adam.  Everything heavy runs on libccx; wrap the
modules in ``torch.nn.parallel.DistributedDataParallel`` (as the reference does, trainMultiGPU.py:233-235) and the
gradient all-reduce over NCCL overlaps with the explicit backward.
"""

import os

import torch

from . import _lib
from ._host import device_twin, named_params, no_gc, stash_host_copy
from .decoder import DecoderWithAttention
from .losses import free_running_cross_entropy, packed_cross_entropy
from .optim import ClampAdam
from .train_ops import direct_grads


def _unwrap(m):
    return m.module if hasattr(m, "module") else m


def make_optimizers(encoder, decoder, decoder_lr=1e-4, encoder_lr=1e-4, grad_clip=5.0):
    """trainMultiGPU.py:191-198,254: Adam over the trainable parameters; clip_gradient fused into the update."""
    dec_opt = ClampAdam([p for p in decoder.parameters() if p.requires_grad], lr=decoder_lr, grad_clip=grad_clip)
    enc_params = [p for p in encoder.parameters() if p.requires_grad]
    enc_opt = ClampAdam(enc_params, lr=encoder_lr, grad_clip=grad_clip) if enc_params else None
    return dec_opt, enc_opt


def caption_train_step(encoder, decoder, imgs, caps, caplens, decoder_optimizer, encoder_optimizer=None, pad_token=0,
                       alpha_c=1.0, teacher_forcing=True, wordMap=None, max_decode_len=51, caplens_host=None):
    """Returns the loss tensor (no host sync).  imgs (B,3,256,256) fp32, caps (B,52) int64, caplens (B,1) int64.
    teacher_forcing=False: the free-running step of trainWithoutTeacherForcing (trainMultiGPU.py:423-460; needs
    wordMap for <start>/<end>).  caplens_host: optional host copy of caplens (what the DataLoader yielded before
    `.to(device)`, trainMultiGPU.py:355-359) — saves the step's only device-to-host read."""
    if teacher_forcing:
        # the decoders need the lengths on the host: read them before the encoder is queued — or not at all when the
        # caller still has the data loader's host tensor (caplens_host), which leaves the step free of host syncs
        stash_host_copy(caplens, caplens_host)
    feats = encoder(imgs)                                                              # trainMultiGPU.py:361
    if not teacher_forcing:
        out = decoder(teacherForcing=False, encoder_out=feats, wordMap=wordMap, maxDecodeLen=max_decode_len)  # :445,:451
        scores, sequences = out[0], out[-1]
        loss, _, _ = free_running_cross_entropy(scores, sequences, caps, wordMap['<end>'], pad_token)   # :446-450
        if isinstance(_unwrap(decoder), DecoderWithAttention):
            loss = loss + alpha_c * ((1.0 - out[1].sum(dim=1)) ** 2).mean()           # :448
    elif isinstance(_unwrap(decoder), DecoderWithAttention):
        scores, caps_sorted, decode_lengths, alphas, _ = decoder(teacherForcing=True, encoder_out=feats,
                                                                 encoded_captions=caps, caption_lengths=caplens)
        loss = packed_cross_entropy(scores, caps_sorted, decode_lengths, unit_grad=True)   # :364-367
        loss = loss + alpha_c * ((1.0 - alphas.sum(dim=1)) ** 2).mean()               # :369
    else:
        kpm = caps == pad_token                                                        # :371
        scores, caps_out, decode_lengths = decoder(teacherForcing=True, encoder_out=feats, encoded_captions=caps,
                                                   caption_lengths=caplens, tgt_key_padding_mask=kpm)
        loss = packed_cross_entropy(scores, caps_out, decode_lengths, unit_grad=True)  # :373-377
    if encoder_optimizer is not None:
        encoder_optimizer.zero_grad(set_to_none=False)
    decoder_optimizer.zero_grad(set_to_none=False)                                     # :381-383
    loss.backward()                                                                    # :384 (DDP all-reduce inside)
    if encoder_optimizer is not None:
        encoder_optimizer.step()                                                       # :387-394 (clamp fused)
    decoder_optimizer.step()
    return loss.detach()


class CapturedTrainStep:
    """The teacher-forced train step (``caption_train_step``: trainMultiGPU.py:357-394) recorded ONCE into a CUDA
    graph and replayed — encoder forward, decoder forward, loss, backward, gradient all-reduce, clamp + Adam — so
    that neither Python nor the ~500 launch calls of a step sit between the GPU and its work (the eager step is
    host-bound: 7.4 ms of enqueue time for ~6 ms of kernels).  An addition: the reference call sites keep working
    through ``caption_train_step``.

        step = CapturedTrainStep(encoder, decoder, decoder_optimizer, encoder_optimizer)
        for imgs, caps, caplens in loader:                 # device tensors of a FIXED shape
            loss = step(imgs, caps, caplens)               # device scalar, valid until the next call

    What makes the step replayable:
      * shapes never depend on the batch: the step buffers span all ``caps.shape[1] - 1`` time steps
        (``decoder.fixed_T``); the recurrence kernels read the longest caption of the batch from the device and stop
        there, rows / steps past a caption's length stay zero, the loss normaliser (number of scored tokens) is a
        device scalar — nothing about the caption lengths is needed on the host;
      * the optimizers count their steps on the device (``ClampAdam.make_capturable``);
      * the kernel-side bf16 copies of the weights are refreshed inside the graph at the top of every step.
    Multi-GPU (``torch.distributed`` initialised, one process per GPU): gradients live in two flat fp32 buckets
    (decoder, fine-tuned encoder stage).  The decoder bucket is all-reduced (NCCL, average) as soon as the decoder's
    backward is done, and each CNBlock's slice of the encoder bucket as soon as that block's backward is done, while
    the blocks before it still back-propagate; each optimizer waits only for its own bucket.
    This replaces DistributedDataParallel's reducer (trainMultiGPU.py:233-236) and is part of the same graph.
    Parameters are broadcast from rank 0 at construction, as DDP does.
    """

    def __init__(self, encoder, decoder, decoder_optimizer, encoder_optimizer=None, pad_token=0, alpha_c=1.0,
                 warmup_steps=3, process_group=None, skip_allreduce=False):
        import torch.distributed as dist
        self.encoder, self.decoder = encoder, decoder
        self.d_opt, self.e_opt = decoder_optimizer, encoder_optimizer
        self.pad_token, self.alpha_c = pad_token, alpha_c
        self.warmup_left = max(int(warmup_steps), 2)     # the first eager steps also build every lazily created cache
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self.lstm = isinstance(decoder, DecoderWithAttention)
        # SMs the persistent GEMMs leave to NCCL while an all-reduce runs next to the encoder backward (0 = none).
        # Measured on 2 x B200 (gpurun_out / profiles/r02_allreduce_attribution.txt): 0, 16 and 32 give the same step
        # time — the all-reduce is not the one waiting for SMs — so the default stays 0.
        self.comm_sms = int(os.environ.get("CCX_COMM_SMS", "0"))
        self.skip_allreduce = skip_allreduce   # diagnostics only (bench.py: how much of the all-reduce is exposed)
        self.graph, self.static, self.loss = None, None, None
        self._shape_key = None
        if self.lstm:
            decoder.fixed_T = True
        self._trainable = [p for _, p in named_params(decoder) if p.requires_grad] + \
                          [p for _, p in named_params(encoder) if p.requires_grad]
        # optional second stream (CCX_STEP_OVERLAP bit 0: the re-cast of the updated weights into their kernel-side copies
        # runs next to the frozen encoder stages; bit 1: the decoder's gradient all-reduce + Adam run next to the encoder's
        # backward).  Off by default: measured on one B200 (bench.py, 30 steps) 5.79 / 5.81 / 5.90 / 5.79 ms per step
        # for modes 0 / 1 / 2 / 3 — the persistent GEMM / dwconv CTAs own every SM, so there is nothing to run next to.
        self.overlap = int(os.environ.get("CCX_STEP_OVERLAP", "0"))
        self._side = None
        self._buckets = []
        for opt in (decoder_optimizer, encoder_optimizer):
            if opt is None:
                continue
            ps = [p for g in opt.param_groups for p in g["params"] if p.requires_grad]
            if not ps:
                continue
            pad = lambda n: (n + 63) // 64 * 64          # every slice starts on a 256-byte boundary
            flat = torch.zeros(sum(pad(p.numel()) for p in ps), dtype=torch.float32, device=ps[0].device)
            off = 0
            for p in ps:                       # gradients are views of one flat bucket: one all-reduce, no copies
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += pad(p.numel())
            opt.make_capturable()
            self._buckets.append(flat)
            if opt is encoder_optimizer:
                # gradient slice of every trainable unit of the encoder (a CNBlock or a downsample child): contiguous
                # inside the bucket because the bucket follows the module order
                where, off = {}, 0
                for p in ps:
                    where[id(p)] = (off, off + pad(p.numel()))
                    off += pad(p.numel())
                self._enc_units = {}
                for child, mod in enumerate(encoder.convnext.children()):
                    units = [(None, mod)] if child % 2 == 0 else list(enumerate(mod))
                    for i, unit in units:
                        spans = [where[id(p)] for p in unit.parameters() if id(p) in where]
                        if spans:
                            self._enc_units[(child, i)] = (min(a for a, _ in spans), max(b for _, b in spans))
        if self.world > 1:
            with torch.no_grad():
                for p in list(decoder.parameters()) + list(encoder.parameters()):
                    dist.broadcast(p.data, src=dist.get_global_rank(process_group, 0) if process_group else 0,
                                   group=process_group)

    def eager_step(self, imgs, caps, caplens):
        """The same step without the graph (debugging, per-kernel timing through the library's event hooks)."""
        if self._shape_key is None:
            return self(imgs, caps, caplens)
        loss = self._body(imgs, caps, caplens)
        for p in self._trainable:
            p._ccx_epoch = getattr(p, "_ccx_epoch", 0) + 1
        return loss

    # ---- one step, eagerly; also the body that is captured ----------------------------------------------------
    def _body(self, imgs, caps, caplens):
        import torch.distributed as dist
        # the only use of the lengths on the host is to size the step buffers; they always span every step here
        fake = self._fake_lens
        stash_host_copy(caplens, fake)
        main = torch.cuda.current_stream()
        if self.overlap and self._side is None:
            self._side = torch.cuda.Stream(device=imgs.device)
        side = self._side if (self.overlap & 1) else None
        if side is not None:
            # the weights the last step moved are re-cast on the side stream while the frozen stages run; the first
            # trainable encoder unit (and everything after it) waits for them
            side.wait_stream(main)
            with torch.cuda.stream(side):
                self.encoder.prepared()
                self.decoder._cache.get()
            self.encoder._before_trainable = lambda: main.wait_stream(side)
        try:
            feats = self.encoder(imgs)                                                     # trainMultiGPU.py:361
        finally:
            self.encoder._before_trainable = None
        if side is not None:
            main.wait_stream(side)         # (frozen encoder: no trainable unit fired the hook)
        enc_trains = feats.requires_grad
        feats_in = feats.detach().requires_grad_() if enc_trains else feats
        if self.lstm:
            scores, caps_s, decode_lengths, alphas, _ = self.decoder(
                teacherForcing=True, encoder_out=feats_in, encoded_captions=caps, caption_lengths=caplens)
        else:
            scores, caps_s, decode_lengths = self.decoder(
                teacherForcing=True, encoder_out=feats_in, encoded_captions=caps, caption_lengths=caplens,
                tgt_key_padding_mask=(caps == self.pad_token))
        n_valid = device_twin(decode_lengths, scores.device).sum().to(torch.float32).reshape(1)
        loss = packed_cross_entropy(scores, caps_s, decode_lengths, n_valid_dev=n_valid, unit_grad=True)
        if self.lstm:
            loss = loss + self.alpha_c * ((1.0 - alphas.sum(dim=1)) ** 2).mean()
        for flat in self._buckets:
            flat.zero_()
        # the backward kernels accumulate straight into the buckets (no returned gradient tensors, no AccumulateGrad adds)
        with direct_grads():
            return self._backward_and_update(loss, feats, feats_in, enc_trains)

    def _backward_and_update(self, loss, feats, feats_in, enc_trains):
        import torch.distributed as dist
        loss.backward()                                   # decoder part: parameter gradients + d features
        dec_work, enc_works, reduced = None, [], set()
        multi = self.world > 1 and not self.skip_allreduce

        def reduce_slice(flat, a, b):
            return dist.all_reduce(flat[a:b], op=dist.ReduceOp.AVG, group=self.group, async_op=True)
        main = torch.cuda.current_stream()
        side = self._side if ((self.overlap & 2) and enc_trains) else None
        if side is not None:
            # the decoder's gradients are final: their all-reduce and the decoder's Adam run on the side stream,
            # under the encoder's backward
            side.wait_stream(main)
            with torch.cuda.stream(side):
                if multi:
                    reduce_slice(self._buckets[0], 0, self._buckets[0].numel()).wait()
                self.d_opt.step()
        elif multi:
            dec_work = reduce_slice(self._buckets[0], 0, self._buckets[0].numel())
        if enc_trains:
            if multi and len(self._buckets) > 1:
                # a unit's slice goes on the wire as soon as its backward is done, under the backward of the units
                # before it (same overlap DDP gets from one autograd node per CNBlock, without its reducer)
                def ready(child, i):
                    if (child, i) in self._enc_units and (child, i) not in reduced:
                        reduced.add((child, i))
                        enc_works.append(reduce_slice(self._buckets[1], *self._enc_units[(child, i)]))
                self.encoder._unit_grads_ready = ready
            if multi and self.comm_sms > 0:
                # the persistent GEMMs leave `comm_sms` SMs to the NCCL kernels that run next to this backward
                _lib.lib().ccx_set_sm_limit(max(_lib.lib().ccx_num_sms() - self.comm_sms, 1))
            try:
                feats.backward(feats_in.grad)             # fine-tuned encoder stage; overlaps the decoder all-reduce
            finally:
                if multi and self.comm_sms > 0:
                    _lib.lib().ccx_set_sm_limit(0)
            self.encoder._unit_grads_ready = None
            if multi and len(self._buckets) > 1:
                for key, (a, b) in self._enc_units.items():       # the first trainable unit (its input has no gradient)
                    if key not in reduced:
                        enc_works.append(reduce_slice(self._buckets[1], a, b))
        if side is None:
            if dec_work is not None:
                dec_work.wait()
            self.d_opt.step()
        if self.e_opt is not None:
            for w in enc_works:
                w.wait()
            self.e_opt.step()
        if side is not None:
            main.wait_stream(side)
        return loss.detach()

    def __call__(self, imgs, caps, caplens):
        key = (tuple(imgs.shape), imgs.dtype, tuple(caps.shape), tuple(caplens.shape))
        if self._shape_key is None:
            self._shape_key = key
            self._fake_lens = torch.full(tuple(caplens.shape), caps.shape[1], dtype=caplens.dtype)
        elif key != self._shape_key:
            raise ValueError(f"CapturedTrainStep was built for inputs {self._shape_key}, got {key}: shapes must not "
                             "change (pad the last batch or use caption_train_step for it)")
        if self.warmup_left > 0:
            self.warmup_left -= 1
            return self._body(imgs, caps, caplens)
        if self.graph is None:
            self.static = (imgs.clone(), caps.clone(), caplens.clone())
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with no_gc(), torch.cuda.graph(self.graph):          # records the step; nothing runs yet
                self.loss = self._body(*self.static)
        for dst, src in zip(self.static, (imgs, caps, caplens)):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        for p in self._trainable:          # the masters moved without this Python code running: kernel-side copies
            p._ccx_epoch = getattr(p, "_ccx_epoch", 0) + 1     # are stale for any eager call that follows
        return self.loss
