"""Flat-buffer SGD: all parameters (and their gradients) are views into one contiguous fp32 buffer, so the
optimiser step is ONE kernel launch and the DDP gradient all-reduce is one (chunkable) buffer.

Semantics and state_dict layout are those of `torch.optim.SGD(momentum, dampening, weight_decay)` as used at
`pretrain/pointcontrast/lib/ddp_trainer.py:107-111` (no dampening) and `downstream/semseg/lib/solvers.py:50-57` (dampening 0.1);
no nesterov.  `ExponentialLR` (`ddp_trainer.py:113`), `PolyLR` below and the checkpoint's `optimizer` entry
(`ddp_trainer.py:155-161`) work unchanged.
"""
import torch

from ._lib import check, lib, ptr, stream
from .me import bump_weights_epoch


class FlatSGD(torch.optim.Optimizer):
    def __init__(self, params, lr, momentum=0.0, weight_decay=0.0, dampening=0.0):
        params = list(params)
        if any(isinstance(p, dict) for p in params):
            raise NotImplementedError("FlatSGD supports a single parameter group")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay, dampening=dampening, nesterov=False))
        ps = self.param_groups[0]["params"]
        dev = ps[0].device
        if dev.type != "cuda" or any(p.device != dev or p.dtype != torch.float32 for p in ps):
            raise RuntimeError("FlatSGD needs float32 CUDA parameters on one device")
        self._offsets, total = [], 0
        for p in ps:
            self._offsets.append(total)
            total += (p.numel() + 3) // 4 * 4           # keep every view 16-byte aligned
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros_like(self.flat_param)
        self.flat_buf = torch.zeros_like(self.flat_param)
        self._gviews = []
        with torch.no_grad():
            for p, off in zip(ps, self._offsets):
                view = self.flat_param[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
                g = self.flat_grad[off:off + p.numel()].view(p.shape)
                if p.grad is not None:
                    g.copy_(p.grad)
                p.grad = g
                self._gviews.append(g)
        self._first = True
        self.grad_scale = 1.0                            # e.g. 1/world_size after a sum all-reduce
        bump_weights_epoch()

    def _views(self, flat):
        return [flat[off:off + p.numel()].view(p.shape) for p, off in zip(self.param_groups[0]["params"], self._offsets)]

    def _grad_views(self):
        """The per-parameter views into flat_grad, created once (187 tensor views per call would cost ~1 ms of host time per step)."""
        v = self.__dict__.get("_gviews")
        if v is None:
            v = self._gviews = self._views(self.flat_grad)
        return v

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()
        for p, g in zip(self.param_groups[0]["params"], self._grad_views()):
            if p.grad is not g and (p.grad is None or p.grad.data_ptr() != g.data_ptr()):
                p.grad = g

    @torch.no_grad()
    def step(self, closure=None):
        grp = self.param_groups[0]
        ps = grp["params"]
        for p, g in zip(ps, self._grad_views()):              # a grad replaced behind our back is folded in
            if p.grad is not g and p.grad is not None and p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad)
                p.grad = g
        with torch.cuda.device(self.flat_param.device):
            check(lib.pcb_sgd_step(ptr(self.flat_param), ptr(self.flat_grad), ptr(self.flat_buf), self.flat_param.numel(),
                                   float(grp["lr"]), float(grp["momentum"]), float(grp["weight_decay"]), float(self.grad_scale),
                                   1 if self._first else 0, float(grp.get("dampening", 0.0)), stream()))
        bump_weights_epoch()           # the kernel wrote the parameters behind torch's back: drop cached bf16 copies
        if self._first:
            for p, b in zip(ps, self._views(self.flat_buf)):
                self.state[p]["momentum_buffer"] = b
            self._first = False

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        ps = self.param_groups[0]["params"]
        loaded = False
        with torch.no_grad():
            for p, b in zip(ps, self._views(self.flat_buf)):
                mb = self.state.get(p, {}).get("momentum_buffer")
                if mb is not None:
                    b.copy_(mb)
                    self.state[p]["momentum_buffer"] = b
                    loaded = True
        self._first = not loaded


class PolyLR(torch.optim.lr_scheduler.LambdaLR):
    """DeepLab polynomial decay, `downstream/semseg/lib/solvers.py:27-32`: lr = base * (1 - step / (max_iter + 1)) ** power,
    stepped once per optimiser step (`lib/train.py:159-160`)."""

    def __init__(self, optimizer, max_iter, power=0.9, last_step=-1):
        super().__init__(optimizer, lambda s: (1 - s / (max_iter + 1)) ** power, last_step)
