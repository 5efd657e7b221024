"""FusedAdam — torch.optim.Adam drop-in for a hig_b200 denoiser whose step() is ONE kernel sweep over the FlatParams
buffers (train_engine.py): gradient clipping by the global norm (clip_grad_norm_, codes/trainers/mul_ddpm_trainer.py:253),
the Adam update (:254-255, optimizer built at :291) and the refresh of the bf16 GEMM-operand mirror, without a host
synchronisation.  state_dict() / load_state_dict() use torch.optim.Adam's format (the reference checkpoints store
`opt_encoder.state_dict()`, :270-272), so `--is_continue` works across the two optimizers.
"""
import torch

from . import ops
from .train_engine import flat_params


class FusedAdam(torch.optim.Adam):
    def __init__(self, module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        net = module.module if hasattr(module, "module") and not hasattr(module, "engine") else module
        if not hasattr(net, "engine"):
            raise TypeError("FusedAdam drives a hig_b200 MotionInteractionTransformer (optionally DataParallel-wrapped)")
        self.net = net
        self.fp = flat_params(net)
        named = dict(net.named_parameters())
        # same parameter list and order as `optim.Adam(self.encoder.parameters(), ...)` (:291): state_dict()s are interchangeable
        super().__init__(list(net.parameters()), lr=lr, betas=betas, eps=eps)
        dev = self.fp.param.device
        self.exp_avg = torch.zeros_like(self.fp.param)
        self.exp_avg_sq = torch.zeros_like(self.fp.param)
        self.gnorm2 = torch.zeros((), device=dev, dtype=torch.float64)
        self.step_count = 0
        self._n_den = self.fp.n_den_params
        self._others = self.fp.params[self._n_den:]
        self._other_gviews = [self.fp.gviews[n] for n in self.fp.other_names]
        self._attach()

    def _attach(self):
        """Denoiser gradients are written in place by the backward graphs: expose them as persistent p.grad views."""
        self.fp.direct = True
        for n, p in zip(self.fp.names[:self._n_den], self.fp.params):
            p.grad = self.fp.gviews[n]

    def _check_layout(self):
        fp = flat_params(self.net)
        if fp is not self.fp:
            raise RuntimeError("FusedAdam: the module's parameters were moved or re-created after the optimizer was built "
                               "(build the optimizer after .to(device))")

    def zero_grad(self, set_to_none=True):
        """The backward graphs zero the flat gradient buffer themselves; only parameters torch.autograd differentiates
        (text side) follow the usual protocol."""
        for p in self._others:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()
        first = self.fp.params[0]
        if first.grad is None or first.grad.data_ptr() != self.fp.gviews[self.fp.names[0]].data_ptr():
            self._attach()

    @torch.no_grad()
    def step(self, closure=None, clip_norm=None):
        if closure is not None:
            raise NotImplementedError("FusedAdam.step takes no closure")
        self._check_layout()
        fp = self.fp
        lo, hi = fp.other_bounds
        if self._others:
            got = [(v, p.grad) for v, p in zip(self._other_gviews, self._others) if p.grad is not None]
            fp.grad[lo:hi].zero_()
            if got:
                torch._foreach_copy_([v for v, _ in got], [g for _, g in got])
        g = self.param_groups[0]
        gn = None
        if clip_norm is not None:
            self.gnorm2.zero_()
            ops.sumsq(fp.grad, self.gnorm2)
            gn = self.gnorm2
        self.step_count += 1
        # denoiser region: parameters + bf16 operand mirror; the rest (text side): parameters only
        n_den = fp.n_den
        ops.adam_flat(fp.param[:n_den], fp.grad[:n_den], self.exp_avg[:n_den], self.exp_avg_sq[:n_den], self.step_count,
                      g["lr"], g["betas"], g["eps"], p_bf16=fp.mirror, gnorm2=gn, max_norm=clip_norm or 0.0)
        if hi > n_den:
            ops.adam_flat(fp.param[n_den:hi], fp.grad[n_den:hi], self.exp_avg[n_den:hi], self.exp_avg_sq[n_den:hi],
                          self.step_count, g["lr"], g["betas"], g["eps"], gnorm2=gn, max_norm=clip_norm or 0.0)
        fp.mark_mirror_fresh(self.net)

    def grad_norm(self):
        """Global gradient norm seen by the last clipped step (device scalar, no synchronisation)."""
        return self.gnorm2.sqrt().float()

    # ------------------------------------------------------------------------------------------ torch.optim.Adam format
    def state_dict(self):
        named = dict(self.net.named_parameters())
        for n in self.fp.names:
            p = named[n]
            self.state[p] = {"step": torch.tensor(float(self.step_count)),
                             "exp_avg": self.fp._view(self.exp_avg, n).clone(),
                             "exp_avg_sq": self.fp._view(self.exp_avg_sq, n).clone()}
        sd = super().state_dict()
        self.state.clear()
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        named = dict(self.net.named_parameters())
        steps = []
        with torch.no_grad():
            for n in self.fp.names:
                st = self.state.get(named[n])
                if not st:
                    continue
                self.fp._view(self.exp_avg, n).copy_(st["exp_avg"])
                self.fp._view(self.exp_avg_sq, n).copy_(st["exp_avg_sq"])
                steps.append(int(float(st["step"])))
        self.step_count = max(steps) if steps else 0
        self.state.clear()
