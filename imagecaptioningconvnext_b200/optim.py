"""``ClampAdam`` — ``clip_gradient`` (utils/utils.py:183-192: grad.clamp_(-c, c)) fused with ``torch.optim.Adam``'s
update (trainMultiGPU.py:387-394) in ONE multi-tensor kernel launch per parameter group (``ccx_adam_clamp``).

State layout and ``state_dict`` keys equal torch.optim.Adam's (``step``, ``exp_avg``, ``exp_avg_sq``), so optimizer
states from reference checkpoints (``torch.optim.Adam.state_dict()``, trainMultiGPU.py:218,223) load unchanged:
``load_state_dict`` fills in the ``grad_clip`` key those lack and refuses ``weight_decay`` / ``amsgrad`` states, which
this kernel does not implement (the reference uses neither).
"""
import math

import torch

from . import _lib
from ._lib import ptr

_CHUNK = 16384


class ClampAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, grad_clip=None):
        # foreach=True: Optimizer.zero_grad(set_to_none=False) then clears all gradients with multi-tensor launches
        # instead of one fill kernel per parameter (40 launches per step for the fine-tuned encoder + LSTM decoder)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, grad_clip=grad_clip, foreach=True))
        self._tables = {}
        self._dev_steps = None     # capturable mode: {group index: float32 [1] device tensor holding the step count}

    def make_capturable(self):
        """Keep the step count on the DEVICE (like torch.optim.Adam(capturable=True)) so that ``step()`` can be
        recorded into a CUDA graph: the bias corrections are then computed inside the kernel from that tensor and
        nothing host-side is baked into the launch.  All parameters of a group must share one step count.  The
        per-parameter ``state['step']`` entries are brought up to date by ``sync_step_counts()`` (state_dict() calls
        it), since a graph replay does not run this Python code."""
        self._dev_steps = {}
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.requires_grad]
            steps = {int(self.state[p]["step"]) for p in ps if len(self.state[p])}
            if len(steps) > 1:
                raise ValueError("ClampAdam.make_capturable: parameters of one group have different step counts")
            dev = ps[0].device if ps else torch.device("cuda")
            self._dev_steps[gi] = torch.full((1,), float(steps.pop() if steps else 0), dtype=torch.float32, device=dev)
        return self

    def sync_step_counts(self):
        """capturable mode: copy the device step counts into the torch.optim.Adam-style per-parameter state."""
        if self._dev_steps is None:
            return
        for gi, group in enumerate(self.param_groups):
            t = float(self._dev_steps[gi].item())
            for p in group["params"]:
                if len(self.state[p]):
                    self.state[p]["step"] = torch.tensor(t)

    def state_dict(self):
        self.sync_step_counts()
        return super().state_dict()

    # ---- checkpoint compatibility ------------------------------------------------------------------------------
    def _normalise_groups(self):
        """A torch.optim.Adam state_dict carries no 'grad_clip' and may carry options this kernel does not have."""
        for group in self.param_groups:
            group.setdefault("grad_clip", self.defaults["grad_clip"])
            group.setdefault("foreach", True)
            if group.get("weight_decay", 0) not in (0, 0.0, None):
                raise ValueError("ClampAdam: weight_decay is not implemented (the reference trains without it)")
            if group.get("amsgrad", False):
                raise ValueError("ClampAdam: amsgrad is not implemented (the reference trains without it)")
            if group.get("maximize", False):
                raise ValueError("ClampAdam: maximize is not implemented")

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._normalise_groups()
        self._tables = {}            # the moment tensors were replaced: cached device pointers are stale
        if self._dev_steps is not None:
            self.make_capturable()

    def __setstate__(self, state):
        super().__setstate__(state)
        self._tables = {}            # Optimizer.__getstate__ only keeps defaults / state / param_groups
        self._dev_steps = None
        self._normalise_groups()

    def _table(self, gi, ps):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                     self.state[p]["exp_avg_sq"].data_ptr()) for p in ps)
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == key:
            return cached[1:]
        rows, be, bo = [], [], []
        for i, p in enumerate(ps):
            stt = self.state[p]
            n = p.numel()
            rows.append([p.data_ptr(), p.grad.data_ptr(), stt["exp_avg"].data_ptr(), stt["exp_avg_sq"].data_ptr(), n])
            for off in range(0, n, _CHUNK):
                be.append(i)
                bo.append(off)
        dev = ps[0].device
        table = torch.tensor(rows, dtype=torch.int64).to(dev)
        block_entry = torch.tensor(be, dtype=torch.int32).to(dev)
        block_offset = torch.tensor(bo, dtype=torch.int64).to(dev)
        total = float(sum(p.numel() for p in ps))
        self._tables[gi] = (key, table, block_entry, block_offset, len(be), total)
        return self._tables[gi][1:]

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for gi, group in enumerate(self.param_groups):
            all_ps = [p for p in group["params"] if p.grad is not None]
            if not all_ps:
                continue
            for p in all_ps:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise ValueError("ClampAdam needs contiguous float32 CUDA parameters and gradients")
                stt = self.state[p]
                if len(stt) == 0:
                    stt["step"] = torch.tensor(0.0)
                    stt["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    stt["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                if not torch.is_tensor(stt["step"]):           # very old checkpoints store a python int
                    stt["step"] = torch.tensor(float(stt["step"]))
                if self._dev_steps is None:
                    stt["step"] += 1
            b1, b2 = group["betas"]
            clip = group.get("grad_clip")
            clip = clip if clip is not None else 0.0
            if self._dev_steps is not None:
                # capturable: count on the device, bias corrections in the kernel (see make_capturable)
                dstep = self._dev_steps[gi]
                dstep += 1
                table, be, bo, nblk, total = self._table((gi, -1), all_ps)
                _lib.check(_lib.lib().ccx_adam_clamp_dev(ptr(table), ptr(be), ptr(bo), nblk, group["lr"], b1, b2,
                                                         group["eps"], ptr(dstep), clip, _CHUNK, total,
                                                         _lib.stream_ptr()), "adam_clamp_dev")
                for p in all_ps:
                    p._ccx_epoch = getattr(p, "_ccx_epoch", 0) + 1
                continue
            # torch.optim.Adam keeps one step count per parameter (bias correction): parameters whose first gradient
            # came later (fine_tune() switched on mid-run) are launched apart, one launch per distinct step value
            by_step = {}
            for p in all_ps:
                by_step.setdefault(int(self.state[p]["step"]), []).append(p)
            b1, b2 = group["betas"]
            clip = group.get("grad_clip")
            clip = clip if clip is not None else 0.0
            for step, ps in by_step.items():
                table, be, bo, nblk, total = self._table((gi, step if len(by_step) > 1 else -1), ps)
                _lib.check(_lib.lib().ccx_adam_clamp(ptr(table), ptr(be), ptr(bo), nblk, group["lr"], b1, b2,
                                                     group["eps"], 1.0 - b1 ** step, math.sqrt(1.0 - b2 ** step), clip,
                                                     _CHUNK, total, _lib.stream_ptr()), "adam_clamp")
            for p in all_ps:      # the weights changed behind torch's back: invalidate kernel-side copies
                p._ccx_epoch = getattr(p, "_ccx_epoch", 0) + 1
        return loss
