"""Host-side helpers shared by the decoder modules: parameter holders whose ``forward`` runs on libccx, and the
prepared-weight cache (bf16 / tf32-split copies of the fp32 master weights, rebuilt when a parameter changes)."""
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


class PreparedCache:
    """Caches ``owner._prepare()`` (dict of kernel-side weight tensors).  When a parameter changes the new values are
    copied INTO the existing buffers (pointer-stable), unless storage / dtype / shapes changed."""

    def __init__(self, owner):
        self._owner = [owner]   # list: keep the module out of nn.Module's attribute registration
        self._key, self._val = None, None

    def get(self):
        owner = self._owner[0]
        params = params_of(owner)
        key = (owner.compute_dtype, params[0].data_ptr(), params[-1].data_ptr(),
               sum(p._version + getattr(p, "_ccx_epoch", 0) for p in params))
        if key != self._key:
            for p in params:
                if not p.is_cuda or p.dtype != torch.float32:
                    raise ValueError("parameters must be float32 CUDA tensors (call .cuda()); there is no CPU path")
            fresh = owner._prepare()
            if self._val is not None and self._key is not None and self._key[:3] == key[:3] and \
                    _same_structure(self._val, fresh):
                _refresh_into(self._val, fresh)
            else:
                self._val = fresh
            self._key = key
        return self._val

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
