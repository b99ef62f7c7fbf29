"""torch.optim.Adam with the update on ONE launch of `uorf_adam_step` (csrc/adam.cu).

The training drivers step Adam over ~80 small fp32 tensors (reference models/uorf_nogan_model.py:59-61, 229-245).  `Adam` below IS
torch.optim.Adam - same constructor, same per-parameter state (`step`, `exp_avg`, `exp_avg_sq`), same `state_dict()` - whose `step()`
hands the update to the kernel when the group is plain (fp32 CUDA tensors, dense gradients, amsgrad / maximize / differentiable off,
step counters on the device: `capturable=True` or `fused=True`) and falls back to torch's own step otherwise.  The chunk table
(device pointers of every 1,024-element piece) is cached on the tensors' addresses: with the flat gradient bucket they never move, so
the table is built once and the step is graph-capturable.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
from torch import optim

from ._lib import lib, check

_CHUNK = 1024
_CHUNK_DTYPE = np.dtype([("param", "<u8"), ("grad", "<u8"), ("exp_avg", "<u8"), ("exp_avg_sq", "<u8"), ("step", "<u8"), ("n", "<i4"),
                         ("reserved", "<i4")])      # struct uorf_adam_chunk, include/uorf_b200.h


class Adam(optim.Adam):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._tables = {}            # group index -> (key, device table, n_chunks)
        self.kernel_steps = 0        # updates done by uorf_adam_step (tests / launch accounting)

    def _kernel_ok(self, group, params, grads, exp_avgs, exp_avg_sqs, steps):
        if group["amsgrad"] or group["maximize"] or group.get("differentiable", False) or not params:
            return False
        dev = params[0].device
        if dev.type != "cuda":
            return False
        for ts in (params, grads, exp_avgs, exp_avg_sqs):
            for t in ts:
                if t.dtype != torch.float32 or t.device != dev or not t.is_contiguous() or t.is_sparse:
                    return False
        return all(torch.is_tensor(s) and s.device == dev and s.dtype == torch.float32 for s in steps)

    def _table(self, gi, params, grads, exp_avgs, exp_avg_sqs, steps):
        key = tuple((p_.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), s.data_ptr(), p_.numel())
                    for p_, g, m, v, s in zip(params, grads, exp_avgs, exp_avg_sqs, steps))
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == key:
            return cached
        if torch.cuda.is_current_stream_capturing():
            return None              # a new table needs a host-to-device copy: not under capture (torch's step runs instead)
        rows = []
        for pp, gp, mp, vp, sp, n in key:
            for off in range(0, n, _CHUNK):
                rows.append((pp + 4 * off, gp + 4 * off, mp + 4 * off, vp + 4 * off, sp, min(_CHUNK, n - off), 0))
        host = np.array(rows, dtype=_CHUNK_DTYPE)
        table = torch.from_numpy(host.view(np.uint8).reshape(-1).copy()).to(params[0].device)
        cached = (key, table, len(rows))
        self._tables[gi] = cached
        return cached

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if getattr(self, "grad_scale", None) is not None or getattr(self, "found_inf", None) is not None:
            super().step()           # a GradScaler drives this step (unscale / skip logic lives in torch's implementation)
            return loss
        plan = []
        for gi, group in enumerate(self.param_groups):
            params, grads, exp_avgs, exp_avg_sqs, max_sqs, steps = [], [], [], [], [], []
            self._init_group(group, params, grads, exp_avgs, exp_avg_sqs, max_sqs, steps)
            if not params:
                continue
            tab = None
            if self._kernel_ok(group, params, grads, exp_avgs, exp_avg_sqs, steps):
                tab = self._table(gi, params, grads, exp_avgs, exp_avg_sqs, steps)
            if tab is None:              # nothing has been updated yet: torch's own step does all groups
                super().step()
                return loss
            plan.append((group, steps, tab, params[0].device))
        for group, steps, (_, table, n_chunks), dev in plan:
            torch._foreach_add_(steps, 1)
            lr = group["lr"]
            beta1, beta2 = group["betas"]
            with torch.cuda.device(dev):
                stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
                lr_dev = lr.data_ptr() if torch.is_tensor(lr) else None
                if torch.is_tensor(lr) and (lr.device != dev or lr.dtype != torch.float32):
                    lr_dev, lr = None, float(lr)
                check(lib().uorf_adam_step(table.data_ptr(), n_chunks, lr_dev, 0.0 if lr_dev else float(lr), float(beta1), float(beta2),
                                           float(group["eps"]), float(group["weight_decay"]), stream), "uorf_adam_step")
            self.kernel_steps += 1
        return loss
