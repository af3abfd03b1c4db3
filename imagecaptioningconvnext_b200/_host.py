"""Host-side helpers shared by the decoder modules: parameter holders whose ``forward`` runs on libccx, and the
prepared-weight cache (bf16 / tf32-split copies of the fp32 master weights, rebuilt when a parameter changes)."""
import os
import weakref

import torch
from torch import nn

from . import _lib
from ._lib import Operand, ptr


class CcxLinear(nn.Linear):
    """nn.Linear's parameters / init / state_dict keys; ``forward`` is the tcgen05 GEMM (no autograd: modules
    that train call the autograd Functions in *_train.py instead)."""

    def __init__(self, in_features, out_features, bias=True, compute_dtype=torch.float32):
        super().__init__(in_features, out_features, bias=bias)
        self.compute_dtype = compute_dtype
        self._key, self._w = None, None

    def operand(self):
        key = (self.weight.data_ptr(), self.weight._version + getattr(self.weight, "_ccx_epoch", 0), self.compute_dtype)
        if key != self._key:
            self._w, self._key = Operand.prepare(self.weight.detach(), self.compute_dtype), key
        return self._w

    def forward(self, x):
        _lib.require_cuda(x, "input")
        shp = x.shape
        a = Operand.prepare(x.reshape(-1, shp[-1]).float().contiguous(), self.compute_dtype)
        y = _lib.linear(a, self.operand(), bias=None if self.bias is None else self.bias.detach())
        return y.view(*shp[:-1], self.out_features)


class CcxEmbedding(nn.Embedding):
    """nn.Embedding's parameters / init; ``forward`` is the gather kernel."""

    def forward(self, tokens):
        _lib.require_cuda(tokens, "tokens")
        tok = tokens.reshape(-1, 1).contiguous().long()
        n, D = tok.shape[0], self.embedding_dim
        out = torch.empty((n, D), dtype=torch.float32, device=tokens.device)
        _lib.check(_lib.lib().ccx_embed_rows(ptr(tok), 1, 0, ptr(self.weight.detach()), self.num_embeddings, D, None,
                                             None, ptr(out), D, 0, None, None, _lib.CCX_F32, 0, 0, n, 1,
                                             _lib.stream_ptr()), "embed_rows")
        return out.view(*tokens.shape, D)


class no_gc:
    """No cyclic garbage collection while a stream captures (``with no_gc(), torch.cuda.graph(g): ...``).  The modules
    here sit in reference cycles (``_owner`` back-references), so a decoder that went out of scope — with the CUDA graphs
    and private memory pools it captured — is only freed by the cyclic collector; if that runs in the middle of a later
    capture, the graph's destructor (cudaGraphExecDestroy, cudaFree of its pool) invalidates the capture
    ("operation failed due to a previous error during capture"; torch.cuda.graph only collects by itself when
    torch.compiler.config.force_cudagraph_gc is set).
    Collect once before the capture begins, then keep the collector off until it ends."""

    def __enter__(self):
        import gc
        self._was = gc.isenabled()
        gc.collect()
        gc.disable()
        return self

    def __exit__(self, *exc):
        import gc
        if self._was:
            gc.enable()


PLAN_REFRESH = [os.environ.get("CCX_PLAN_REFRESH", "1") != "0"]    # A/B switch: 0 = re-run _prepare() on every refresh


def _refresh_into(old, new):
    """Copy freshly prepared values into the buffers of a previous preparation (same structure), so that device
    pointers stay stable across optimizer steps (weight tables, captured CUDA graphs)."""
    if isinstance(old, dict):
        for k in old:
            _refresh_into(old[k], new[k])
    elif isinstance(old, (list, tuple)):
        for a, b in zip(old, new):
            _refresh_into(a, b)
    elif isinstance(old, Operand):
        if old.hi.data_ptr() != new.hi.data_ptr():      # (refreshed in place by Operand.prepare: nothing to copy)
            old.hi.copy_(new.hi)
            if old.lo is not None:
                old.lo.copy_(new.lo)
    elif torch.is_tensor(old):
        if old.data_ptr() != new.data_ptr():      # views of the parameter storage need no copy
            old.copy_(new)


def _same_structure(a, b):
    if type(a) is not type(b):
        return False
    if isinstance(a, dict):
        return a.keys() == b.keys() and all(_same_structure(a[k], b[k]) for k in a)
    if isinstance(a, (list, tuple)):
        return len(a) == len(b) and all(_same_structure(x, y) for x, y in zip(a, b))
    if isinstance(a, Operand):
        return a.hi.shape == b.hi.shape and a.dtype == b.dtype
    if torch.is_tensor(a):
        return a.shape == b.shape and a.dtype == b.dtype
    return True


class RefreshPlan:
    """The weight refresh of a module as ONE ``ccx_cast_segments`` launch: a table of rectangular pieces
    ``dst <- src`` (optionally row-permuted, transposed, or the sum of two sources) built once against the buffers of
    a preparation, replayed after every optimizer step.  ``add`` takes 2-D views (unit inner stride) — slices of the
    prepared buffers on the left, slices of the fp32 parameters on the right — so concatenations are just several
    pieces with different destination views."""

    def __init__(self):
        self.segs, self.keep, self.tiles, self.bytes = [], [], 0, 0.0
        self._table = None

    @staticmethod
    def _2d(t):
        return t.unsqueeze(0) if t.dim() == 1 else t

    def add(self, dst, src, src2=None, row_map=None, transpose=False):
        dst, src = self._2d(dst), self._2d(src.detach())
        rows = src.shape[0] if row_map is None else int(row_map.numel())
        cols = src.shape[1]
        want = (cols, rows) if transpose else (rows, cols)
        if tuple(dst.shape) != want or dst.stride(1) != 1 or (src.stride(1) != 1 and cols > 1) or \
                src.dtype != torch.float32 or dst.dtype not in (torch.bfloat16, torch.float32):
            raise ValueError(f"RefreshPlan.add: dst {tuple(dst.shape)} {dst.dtype} vs src block {want}")
        seg = _lib.CastSeg()
        seg.src, seg.dst = src.data_ptr(), dst.data_ptr()
        seg.src_ld, seg.dst_ld = src.stride(0), dst.stride(0)
        if src2 is not None:
            src2 = self._2d(src2.detach())
            if tuple(src2.shape) != tuple(src.shape) or src2.stride(0) != src.stride(0) or src2.dtype != torch.float32:
                raise ValueError("RefreshPlan.add: src2 must mirror src")
            seg.src2 = src2.data_ptr()
        if row_map is not None:
            row_map = row_map.to(device=src.device, dtype=torch.int32).contiguous()
            seg.row_map = row_map.data_ptr()
        seg.rows, seg.cols = rows, cols
        seg.flags = (1 if transpose else 0) | (2 if dst.dtype == torch.float32 else 0)
        seg.tile0 = self.tiles
        self.tiles += ((rows + 63) // 64) * ((cols + 63) // 64)
        self.bytes += rows * cols * (4.0 * (2 if src2 is not None else 1) + dst.element_size())
        self.segs.append(seg)
        self.keep.append((dst, src, src2, row_map))
        self._table = None
        return self

    def sources(self):
        return [(k[1].data_ptr(), k[0].data_ptr()) for k in self.keep]

    def run(self):
        if not self.segs:
            return
        if self._table is None:
            arr = (_lib.CastSeg * len(self.segs))(*self.segs)
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            self._table = host.to(self.keep[0][0].device)
        _lib.check(_lib.lib().ccx_cast_segments(self._table.data_ptr(), len(self.segs), self.tiles, self.bytes,
                                                _lib.stream_ptr()), "cast_segments")


class PreparedCache:
    """Caches ``owner._prepare()`` (dict of kernel-side weight tensors).  When a parameter changes, the existing
    buffers are refreshed in place (pointer-stable: weight tables and captured CUDA graphs hold their addresses),
    unless storage / dtype / shapes changed: by the owner's one-launch ``_refresh_plan`` (bf16), else by re-running
    ``_prepare()`` with every ``Operand.prepare`` converting straight into the operand it made last time."""

    def __init__(self, owner):
        self._owner = [owner]   # list: keep the module out of nn.Module's attribute registration
        self._key, self._val, self._made, self._plan, self._plan_src = None, None, None, None, None

    def get(self):
        owner = self._owner[0]
        params = params_of(owner)
        key = (owner.compute_dtype, params[0].data_ptr(), params[-1].data_ptr(),
               sum(p._version + getattr(p, "_ccx_epoch", 0) for p in params))
        if key != self._key:
            for p in params:
                if not p.is_cuda or p.dtype != torch.float32:
                    raise ValueError("parameters must be float32 CUDA tensors (call .cuda()); there is no CPU path")
            refresh = self._val is not None and self._key is not None and self._key[:3] == key[:3]
            if refresh and self._run_plan(owner):
                self._key = key
                return self._val
            # a refresh converts the changed masters straight into the operands of the previous preparation (same
            # call sequence), so the bf16 / tf32 copies are written once and their device pointers never move
            with _lib.prepare_log(reuse=self._made if refresh else None) as log:
                fresh = owner._prepare()
            if refresh and _same_structure(self._val, fresh):
                _refresh_into(self._val, fresh)
            else:
                self._val = fresh
                self._made = log["made"]
                self._plan = None
            self._key = key
        return self._val

    def _run_plan(self, owner):
        """One-launch refresh (``owner._refresh_plan(prepared dict) -> RefreshPlan``) of the current buffers; False
        when the module has no plan for this compute dtype or a parameter's storage moved since the plan was built.
        The plan (and its device table: one host-to-device copy) is built at the FIRST refresh, which must not happen
        inside a stream capture — every capturing caller (CapturedTrainStep, enable_cuda_graph) runs at least two eager
        steps first, so the first refresh is the second eager step's."""
        make = getattr(owner, "_refresh_plan", None)
        if make is None or not PLAN_REFRESH[0]:
            return False
        if self._plan is None:
            self._plan = make(self._val)
            if self._plan is None:
                self._plan = False
            else:
                self._plan_src = [(p, p.data_ptr()) for p in params_of(owner)]
        if self._plan is False:
            return False
        if any(p.data_ptr() != a for p, a in self._plan_src):
            self._plan = None
            return False
        self._plan.run()
        return True

    def storage_key(self):
        return None if self._key is None else self._key[:3]

    def invalidate(self):
        self._key = None


# ---- host copies of small device tensors --------------------------------------------------------------------------
# The reference reads the caption lengths back in the middle of the step (`.tolist()`, models/decoder.py:91,
# models/transformerDecoder.py:92): the host then waits for the encoder forward before it can enqueue the decoder.
# A caller that knows the lengths early (train_step.caption_train_step) stashes a host copy BEFORE launching the
# encoder — or hands over the host tensor it still has from the data loader, in which case the step has no
# host<->device synchronisation at all and the host can run ahead of the GPU across steps.
# Entries are bound to the tensor OBJECT (weak reference) and its version counter, so a recycled address or an
# in-place update can never produce a stale answer.
_HOST_COPIES = {}


def stash_host_copy(t, host=None):
    """Remember a host copy of device tensor `t` (made now, synchronising, unless `host` is supplied)."""
    if len(_HOST_COPIES) >= 8:
        _HOST_COPIES.clear()
    h = t.detach().cpu() if host is None else host.detach()
    if tuple(h.shape) != tuple(t.shape):
        raise ValueError("host copy has a different shape")
    _HOST_COPIES[id(t)] = (weakref.ref(t), t._version, h)


def host_copy(t):
    """Host copy of a device tensor: the stashed one if `t` is that very tensor, unmodified; else a D2H read."""
    if not t.is_cuda:
        return t
    e = _HOST_COPIES.get(id(t))
    if e is not None and e[0]() is t and e[1] == t._version:
        return e[2]
    return t.detach().cpu()


# ---- cached parameter lists ------------------------------------------------------------------------------------------
# nn.Module.named_parameters() walks the module tree (named_modules + a dedup set) on every call; the train step asked
# for it ~1,500 times per step (requires_grad checks, gradient dicts), about a quarter of the host time of a step that
# is host-bound.  The list is cached on the module and re-validated by resolving its first and last entry.
def _resolve(module, dotted):
    obj = module
    for part in dotted.split("."):
        obj = obj._modules[part] if part in obj._modules else obj._parameters[part]
    return obj


def named_params(module):
    """[(name, parameter)] of `module`, cached."""
    c = module.__dict__.get("_ccx_named_params")
    if c is not None:
        try:
            if not c or (_resolve(module, c[0][0]) is c[0][1] and _resolve(module, c[-1][0]) is c[-1][1]):
                return c
        except (KeyError, AttributeError):
            pass
    c = list(module.named_parameters())
    module.__dict__["_ccx_named_params"] = c
    return c


def params_of(module):
    return [p for _, p in named_params(module)]


def any_requires_grad(module):
    for _, p in named_params(module):
        if p.requires_grad:
            return True
    return False


# ---- device twins of small host lists --------------------------------------------------------------------------------
# The reference API hands `decode_lengths` around as a Python list (models/decoder.py:91).  Turning such a list into a
# device tensor with torch.tensor(list, device="cuda") is a BLOCKING copy — it drains the stream, twice per train
# step.  Whoever produced the list from a device tensor registers that tensor here; consumers ask for it back.
_DEVICE_TWINS = {}


def stash_device_twin(lst, tensor):
    if len(_DEVICE_TWINS) >= 8:
        _DEVICE_TWINS.clear()
    _DEVICE_TWINS[id(lst)] = (lst, tensor)


def device_twin(lst, device, dtype=torch.int64):
    """Device tensor holding `lst`: the registered twin if `lst` is that very list, else an asynchronous upload
    from pinned memory (never a blocking copy)."""
    e = _DEVICE_TWINS.get(id(lst))
    if e is not None and e[0] is lst and e[1].device == torch.device(device):
        return e[1].to(dtype)
    if isinstance(lst, torch.Tensor):
        return lst.to(device=device, dtype=dtype)
    return torch.tensor(lst, dtype=dtype).pin_memory().to(device, non_blocking=True)
