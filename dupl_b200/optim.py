"""PolyWarmupAdamW (utils/optimizer.py:38-68) as ONE fused multi-tensor kernel over both students.

`FusedPolyWarmupAdamW` has the optimizer surface the reference script uses — `step()`, `zero_grad()`,
`param_groups[i]["lr"]` (read for logging at train_final_voc.py:478), `global_step`, `state_dict()` — and torch.optim.AdamW's
arithmetic (decoupled weight decay, bias correction from each parameter's own step count; a parameter whose .grad is None is
skipped entirely), but issues two launches per step whatever the number of tensors, keeps every scalar the kernels read on
the device (CUDA-graph replayable) and refreshes the split-bf16 planes of the GEMM weights in the same pass.
"""
import ctypes as C

import torch

from . import _lib as L


class FusedPolyWarmupAdamW:
    def __init__(self, params, lr, weight_decay, betas, warmup_iter=None, max_iter=None, warmup_ratio=None, power=None,
                 plane_lookup=None, **kwargs):
        """params: list of param-group dicts {"params": [...], "lr": ..., "weight_decay": ...} as utils/train_helper.py builds
        them.  plane_lookup(param) -> (hi, lo) bf16 planes with the parameter's element order, or None."""
        self.param_groups = [dict(g) for g in params]
        for g in self.param_groups:
            g.setdefault("lr", lr)
            g.setdefault("weight_decay", weight_decay)
            g["params"] = list(g["params"])
            if g["weight_decay"] != weight_decay:
                raise ValueError("the fused kernel applies one weight decay to every group (as the reference does)")
        self._init_lr = [g["lr"] for g in self.param_groups]
        self.defaults = dict(lr=lr, betas=betas, weight_decay=weight_decay, eps=1e-8)
        self.betas, self.eps, self.weight_decay = betas, 1e-8, weight_decay
        self.global_step = 0
        self.warmup_iter, self.warmup_ratio, self.max_iter, self.power = warmup_iter, warmup_ratio, max_iter, power
        self._plane_lookup = plane_lookup
        self._built = False

    # ------------------------------------------------------------------ tables on the device
    def _build(self):
        ps, lrs = [], []
        for g, lr0 in zip(self.param_groups, self._init_lr):
            for p in g["params"]:
                if p.requires_grad:
                    ps.append(p)
                    lrs.append(lr0)
        if not ps:
            raise ValueError("no trainable parameters")
        dev = ps[0].device
        L.require_cuda(*ps)
        self._params, self._lrs, self._dev = ps, lrs, dev
        n = len(ps)
        numel = (C.c_int64 * n)(*[p.numel() for p in ps])
        cnt = C.c_int64(0)
        L.check(L.lib().dupl_adamw_items(numel, n, None, 0, C.byref(cnt)), "dupl_adamw_items")
        items = torch.empty(cnt.value, 2, dtype=torch.int32)
        L.check(L.lib().dupl_adamw_items(numel, n, C.cast(items.data_ptr(), L.c_i32p), cnt.value, C.byref(cnt)), "dupl_adamw_items")
        self._items = items.to(dev)
        # moments in two flat buffers (views per tensor, 64-element aligned like the gradient arenas)
        offs, o = [], 0
        for p in ps:
            offs.append(o)
            o += (p.numel() + 63) // 64 * 64
        self._m = torch.zeros(o, dtype=torch.float32, device=dev)
        self._v = torch.zeros(o, dtype=torch.float32, device=dev)
        self.state = {p: {"exp_avg": self._m[f:f + p.numel()].view(p.shape), "exp_avg_sq": self._v[f:f + p.numel()].view(p.shape)}
                      for p, f in zip(ps, offs)}
        self._offs = offs
        self._steps = torch.zeros(n, dtype=torch.int32, device=dev)
        self._coef = torch.zeros(n, 2, dtype=torch.float32, device=dev)
        self._active = torch.zeros(n, dtype=torch.int32, device=dev)
        self._active_host = None
        self._lr_scale = torch.ones((), dtype=torch.float32, device=dev)
        self._table = torch.zeros(n * C.sizeof(L.AdamwParam), dtype=torch.uint8, device=dev)
        self._grad_ptrs = None
        self._built = True

    def _refresh_table(self):
        """(Re)writes the pointer table when a gradient / parameter / plane pointer or the set of parameters with a gradient
        changed (first step, phase change).  Host work only when something moved."""
        ps = self._params
        grad_ptrs = tuple(0 if p.grad is None else p.grad.data_ptr() for p in ps)
        data_ptrs = tuple(p.data_ptr() for p in ps)
        key = (grad_ptrs, data_ptrs)
        if key == self._grad_ptrs:
            return
        tab = (L.AdamwParam * len(ps))()
        active = []
        for i, (p, f, lr0) in enumerate(zip(ps, self._offs, self._lrs)):
            e = tab[i]
            g = p.grad
            active.append(0 if g is None else 1)
            if g is not None and (not g.is_contiguous() or g.dtype != torch.float32 or g.data_ptr() % 16):
                raise RuntimeError("fused AdamW needs contiguous 16-byte aligned fp32 gradients")
            e.param, e.grad = p.data_ptr(), (0 if g is None else g.data_ptr())
            e.exp_avg, e.exp_avg_sq = self._m.data_ptr() + 4 * f, self._v.data_ptr() + 4 * f
            planes = self._plane_lookup(p) if self._plane_lookup is not None else None
            if planes is not None:
                e.plane_hi, e.plane_lo = planes[0].data_ptr(), planes[1].data_ptr()
            e.numel, e.lr = p.numel(), lr0
        raw = torch.frombuffer(bytearray(bytes(tab)), dtype=torch.uint8)
        self._table.copy_(raw)
        self._active.copy_(torch.tensor(active, dtype=torch.int32))
        self._grad_ptrs = key

    # ------------------------------------------------------------------ schedule (utils/optimizer.py:52-66)
    def advance_schedule(self):
        if not self._built:
            self._build()
        if self.global_step < self.warmup_iter:
            m = 1 - (1 - self.global_step / self.warmup_iter) * (1 - self.warmup_ratio)
        elif self.global_step < self.max_iter:
            m = (1 - self.global_step / self.max_iter) ** self.power
        else:
            m = None
        if m is not None:
            for g, lr0 in zip(self.param_groups, self._init_lr):
                g["lr"] = lr0 * m
            self._lr_scale.fill_(m)
        self.global_step += 1

    def step_captured(self, closure=None):
        """The parameter update alone (what a CUDA graph captures); the caller advances the schedule."""
        if not self._built:
            self._build()
        self._refresh_table()
        a = L.AdamwArgs()
        a.params, a.items, a.active = self._table.data_ptr(), self._items.data_ptr(), self._active.data_ptr()
        a.steps, a.coef, a.lr_scale = self._steps.data_ptr(), self._coef.data_ptr(), self._lr_scale.data_ptr()
        a.n_params, a.n_items = len(self._params), self._items.shape[0]
        a.beta1, a.beta2, a.eps, a.weight_decay = self.betas[0], self.betas[1], self.eps, self.weight_decay
        L.check(L.lib().dupl_adamw_step(C.byref(a), L.stream_ptr(self._dev)), "dupl_adamw_step")

    def step(self, closure=None):
        self.advance_schedule()
        self.step_captured()

    # ------------------------------------------------------------------ what a captured graph of one phase depends on
    def export_tables(self):
        """The device tables as they are now (pointer table, active flags): a CUDA graph captured in this phase reads them at
        replay time, so a driver that keeps graphs of several phases restores them before replaying another phase's graph."""
        if not self._built:
            self._build()
        return dict(table=self._table.clone(), active=self._active.clone(), key=self._grad_ptrs)

    def import_tables(self, t):
        self._table.copy_(t["table"])
        self._active.copy_(t["active"])
        self._grad_ptrs = t["key"]

    def snapshot(self):
        if not self._built:
            self._build()
        return dict(m=self._m.clone(), v=self._v.clone(), steps=self._steps.clone(), global_step=self.global_step,
                    lrs=[g["lr"] for g in self.param_groups])

    def restore(self, snap):
        self._m.copy_(snap["m"])
        self._v.copy_(snap["v"])
        self._steps.copy_(snap["steps"])
        self.global_step = snap["global_step"]
        for g, lr in zip(self.param_groups, snap["lrs"]):
            g["lr"] = lr

    def zero_grad(self, set_to_none=True):
        for g in self.param_groups:
            for p in g["params"]:
                if set_to_none:
                    p.grad = None
                elif p.grad is not None:
                    p.grad.zero_()

    def state_dict(self):
        if not self._built:
            self._build()
        return {"global_step": self.global_step, "steps": self._steps.cpu(), "exp_avg": self._m.cpu(), "exp_avg_sq": self._v.cpu()}

    def load_state_dict(self, sd):
        if not self._built:
            self._build()
        self.global_step = sd["global_step"]
        self._steps.copy_(sd["steps"])
        self._m.copy_(sd["exp_avg"])
        self._v.copy_(sd["exp_avg_sq"])
