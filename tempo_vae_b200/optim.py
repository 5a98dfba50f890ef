"""Fused clip-by-global-norm + AdamW over flat fp32 buffers (tvae_sumsq + tvae_adamw).

Replaces `torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)` + `torch.optim.AdamW.step()` of
src/train_utils.py:175-177 / src/model.py:756-758. Parameters are re-pointed into one flat fp32 buffer (each
start aligned to 64 B); gradients, exp_avg and exp_avg_sq live in matching flat buffers, so one step is two
kernels regardless of the number of tensors, and a data-parallel all-reduce works on contiguous ranges.

state_dict()/load_state_dict() keep torch.optim.AdamW's format (per-parameter `step`, `exp_avg`, `exp_avg_sq`), so
checkpoints written by the reference Trainer load here and vice versa (src/train_utils.py:122-147).

Semantics kept from the reference run: one parameter group, decoupled weight decay on every tensor that has a
gradient (norm scales, biases and `logvar` included), parameters whose .grad is None are skipped entirely
(the never-used downs.2.down / ups.2.up).
"""
import torch

from . import ops

_ALIGN = 16  # elements (64 bytes)


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=None):
        betas = tuple(betas)
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FusedAdamW supports a single parameter group (as the reference's AdamW call does)")
        self.max_grad_norm = max_grad_norm
        self._flatten()

    # ------------------------------------------------------------------------------------------------ layout
    def _flatten(self):
        ps = [p for p in self.param_groups[0]["params"]]
        if not ps:
            raise ValueError("no parameters")
        dev = ps[0].device
        if dev.type != "cuda":
            raise ops._lib.TvaeError("FusedAdamW needs CUDA parameters (there is no CPU fallback)")
        offs, total = [], 0
        for p in ps:
            if p.device != dev or p.dtype != torch.float32:
                raise ValueError("all parameters must be fp32 on one CUDA device")
            offs.append(total)
            total += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self._offsets, self._total = offs, total
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self._step = 0
        self._ever_live = set()
        with torch.no_grad():
            for p, off in zip(ps, offs):
                n = p.numel()
                self.flat_param[off:off + n].copy_(p.data.reshape(-1))
                p.data = self.flat_param[off:off + n].view(p.shape)
                p._tvae_grad = self.flat_grad[off:off + n].view(p.shape)
                p._tvae_flat_range = (off, off + n)
                p.grad = None
        self._bump()

    @staticmethod
    def _bump():
        from .model import ENGINE
        ENGINE.params_changed()

    def param_ranges(self):
        """[(param, start, end)] in flat-buffer order."""
        return [(p,) + p._tvae_flat_range for p in self.param_groups[0]["params"]]

    # ------------------------------------------------------------------------------------------------ stepping
    def zero_grad(self, set_to_none: bool = True):
        for p in self.param_groups[0]["params"]:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm=None, grad_scale=1.0):
        """One AdamW step over every parameter that has a gradient. `max_grad_norm` (default: the constructor's)
        fuses clip_grad_norm_ into the update: the clip coefficient is computed on the device."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        ps = self.param_groups[0]["params"]
        g = self.param_groups[0]
        live = [p for p in ps if p.grad is not None]
        if not live:
            return loss
        for p in live:  # gradients that autograd produced outside the flat buffer are folded in
            if p.grad.data_ptr() != p._tvae_grad.data_ptr():
                p._tvae_grad.copy_(p.grad)
                p.grad = p._tvae_grad
        dead = [p for p in ps if p.grad is None]
        for p in dead:  # skipped tensors stay inert: no decay (ranges below) and nothing stale in the global norm
            if id(p) in self._ever_live:
                p._tvae_grad.zero_()
        self._ever_live.update(id(p) for p in live)
        self._step += 1
        max_norm = self.max_grad_norm if max_grad_norm is None else max_grad_norm
        clip = max_norm is not None and max_norm > 0
        if clip:
            ops.sumsq(self.flat_grad, self._sumsq)
        kw = dict(lr=float(g["lr"]), beta1=float(g["betas"][0]), beta2=float(g["betas"][1]), eps=float(g["eps"]),
                  weight_decay=float(g["weight_decay"]), step=self._step, sumsq_buf=self._sumsq if clip else None,
                  max_norm=float(max_norm) if clip else 0.0, grad_scale=float(grad_scale))
        for s, e in self._live_ranges(dead):
            ops.adamw(self.flat_param[s:e], self.flat_grad[s:e], self.flat_exp_avg[s:e], self.flat_exp_avg_sq[s:e], **kw)
        self._bump()
        return loss

    def _live_ranges(self, dead):
        """Contiguous flat ranges that exclude parameters without gradients (16-element aligned on both ends)."""
        if not dead:
            return [(0, self._total)]
        cuts = sorted((p._tvae_flat_range[0], (p._tvae_flat_range[1] + _ALIGN - 1) // _ALIGN * _ALIGN) for p in dead)
        out, cur = [], 0
        for s, e in cuts:
            if s > cur:
                out.append((cur, s))
            cur = max(cur, e)
        if cur < self._total:
            out.append((cur, self._total))
        return out

    def grad_norm(self):
        """Global L2 norm of the current gradients (device scalar, fp64)."""
        ops.sumsq(self.flat_grad, self._sumsq)
        return self._sumsq.sqrt()

    # ------------------------------------------------------------------------------------------------ checkpoints
    def state_dict(self):
        """torch.optim.AdamW-compatible: state[i] = {step, exp_avg, exp_avg_sq} for parameters that were updated."""
        ps = self.param_groups[0]["params"]
        state = {}
        if self._step > 0:
            for i, p in enumerate(ps):
                if id(p) not in self._ever_live:
                    continue  # never received a gradient: torch's AdamW has no state entry for it either
                s, e = p._tvae_flat_range
                state[i] = {
                    "step": torch.tensor(float(self._step)),
                    "exp_avg": self.flat_exp_avg[s:e].view(p.shape).clone(),
                    "exp_avg_sq": self.flat_exp_avg_sq[s:e].view(p.shape).clone(),
                }
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group.setdefault("amsgrad", False)
        group.setdefault("maximize", False)
        group.setdefault("foreach", None)
        group.setdefault("capturable", False)
        group.setdefault("differentiable", False)
        group.setdefault("fused", None)
        group["params"] = list(range(len(ps)))
        return {"state": state, "param_groups": [group]}

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        ps = self.param_groups[0]["params"]
        groups = state_dict["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(ps):
            raise ValueError("loaded state dict has a different number of parameter groups / parameters")
        for k in ("lr", "betas", "eps", "weight_decay"):
            if k in groups[0]:
                self.param_groups[0][k] = tuple(groups[0][k]) if k == "betas" else groups[0][k]
        self.flat_exp_avg.zero_()
        self.flat_exp_avg_sq.zero_()
        step = 0
        for idx, st in state_dict["state"].items():
            p = ps[int(idx)]
            s, e = p._tvae_flat_range
            self.flat_exp_avg[s:e].copy_(st["exp_avg"].reshape(-1))
            self.flat_exp_avg_sq[s:e].copy_(st["exp_avg_sq"].reshape(-1))
            step = max(step, int(float(st["step"])))
            self._ever_live.add(id(p))
        self._step = step
        self._bump()
