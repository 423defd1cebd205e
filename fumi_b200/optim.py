"""Outer-loop optimizers backed by the fused Adam kernel (fumi_adam_step in the C ABI).

Reference: utils.init_optim (fumi/utils/utils.py:277-299) builds torch.optim.Adam(lr, weight_decay)
by default (L2 term in the gradient), AdamW for --optim adamw*, SGD for --optim SGD.
FusedAdam keeps torch's Optimizer interface and state_dict schema (per-parameter 'step', 'exp_avg',
'exp_avg_sq'), so reference checkpoints load, but parameters, gradients and both moments are views
of four flat fp32 buffers: the whole step is ONE kernel launch and the multi-GPU meta-gradient
all-reduce is ONE call on the flat gradient.
"""
import torch

from . import _lib


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled=decoupled)
        super().__init__(params, defaults)
        self._flat = {}
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.requires_grad]
            if not ps or not all(p.dtype == torch.float32 for p in ps):
                continue
            dev = ps[0].device
            _lib.require_cuda(dev, "FusedAdam (move the model to the device first, as utils.init_model does)")
            n = sum(p.numel() for p in ps)
            fp = torch.empty(n, dtype=torch.float32, device=dev)
            # two extra floats behind the gradients carry [loss, acc] of the meta-batch, so that the multi-GPU outer
            # step needs ONE all-reduce; the Adam kernel only sees the first n
            fg_ext = torch.zeros(n + 2, dtype=torch.float32, device=dev)
            fg = fg_ext[:n]
            fm = torch.zeros_like(fp)
            fv = torch.zeros_like(fp)
            off = 0
            for p in ps:
                k = p.numel()
                fp[off:off + k].copy_(p.data.reshape(-1))
                p.data = fp[off:off + k].view(p.shape)
                p.grad = fg[off:off + k].view(p.shape)
                p._fumi_flat_opt = self
                st = self.state[p]
                st["step"] = torch.zeros((), dtype=torch.float32)
                st["exp_avg"] = fm[off:off + k].view(p.shape)
                st["exp_avg_sq"] = fv[off:off + k].view(p.shape)
                off += k
            self._flat[gi] = dict(p=fp, g=fg, g_ext=fg_ext, m=fm, v=fv, params=ps, step=0)

    def zero_grad(self, set_to_none=False):
        """One memset of the flat gradient buffer.  Gradients that were detached from it (Module.zero_grad(),
        set_to_none, a foreign backward) are re-attached to their flat views, so the single-launch Adam step and
        the single flat all-reduce stay valid."""
        for f in self._flat.values():
            f["g_ext"].zero_()
            off = 0
            for p in f["params"]:
                k = p.numel()
                if p.grad is None or p.grad.data_ptr() != f["g"].data_ptr() + 4 * off:
                    p.grad = f["g"][off:off + k].view(p.shape)
                off += k

    def flat_grad(self, params):
        """The flat gradient buffer (n gradients + 2 scalars for [loss, acc]) if every gradient of `params` still
        lives in it, in order; else None."""
        for f in self._flat.values():
            if len(f["params"]) == len(params) and all(a is b for a, b in zip(f["params"], params)):
                return f["g_ext"] if self._is_flat(f) else None
        return None

    def _is_flat(self, f):
        off = 0
        for p in f["params"]:
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != f["g"].data_ptr() + 4 * off or \
                    p.data.data_ptr() != f["p"].data_ptr() + 4 * off:
                return False
            st = self.state[p]
            if st["exp_avg"].data_ptr() != f["m"].data_ptr() + 4 * off or \
                    st["exp_avg_sq"].data_ptr() != f["v"].data_ptr() + 4 * off:
                return False
            off += k
        return True

    @torch.no_grad()
    def step(self, closure=None):
        L = _lib.lib()
        for gi, group in enumerate(self.param_groups):
            f = self._flat.get(gi)
            if f is None:
                continue
            b1, b2 = group["betas"]
            args = (float(group["lr"]), float(b1), float(b2), float(group["eps"]), float(group["weight_decay"]))
            stream = _lib.stream_ptr(f["p"].device) if f["p"].is_cuda else None
            step = int(self.state[f["params"][0]]["step"].item()) + 1
            if self._is_flat(f):
                _lib.check(L.fumi_adam_step(_lib.ptr(f["p"]), _lib.ptr(f["g"]), _lib.ptr(f["m"]), _lib.ptr(f["v"]),
                                            f["p"].numel(), *args, step, int(group["decoupled"]), stream),
                           "fumi_adam_step")
            else:       # views were replaced (e.g. load_state_dict / zero_grad(set_to_none)): per-tensor launches
                if not getattr(self, "_warned_unflat", False):
                    import warnings
                    warnings.warn("FusedAdam: parameter/gradient/moment views left the flat buffers; falling back to "
                                  "one Adam launch per tensor (call FusedAdam.zero_grad() to re-attach gradients)")
                    self._warned_unflat = True
                for p in f["params"]:
                    if p.grad is None:
                        continue
                    st = self.state[p]
                    g = p.grad.contiguous()
                    st["exp_avg"] = st["exp_avg"].to(p.device).contiguous()
                    st["exp_avg_sq"] = st["exp_avg_sq"].to(p.device).contiguous()
                    _lib.check(L.fumi_adam_step(_lib.ptr(p.data), _lib.ptr(g), _lib.ptr(st["exp_avg"]),
                                                _lib.ptr(st["exp_avg_sq"]), p.numel(), *args, step,
                                                int(group["decoupled"]), stream), "fumi_adam_step")
            for p in f["params"]:
                self.state[p]["step"] = torch.tensor(float(step))
