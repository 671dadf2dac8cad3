"""FusedAdam — the optimiser the reference builds at /root/reference/model/modelTrainer.py:44
(``torch.optim.Adam(model.parameters(), lr, weight_decay)``: L2 decay added to the gradient, bias
correction, no amsgrad) with the whole update of a tensor in ONE engine kernel
(``rgcn_adam_step``).  Same constructor defaults and state names
(``step``, ``exp_avg``, ``exp_avg_sq``) as torch's Adam; CUDA fp32 parameters only.

``capturable=True`` keeps the step count in device memory (``rgcn_adam_step_dev``), so the whole
training step can be captured in a CUDA graph (``rgcn_b200.trainer.GraphedTrainStep``).  That counter
is ONE per device, shared by all tensors of the optimiser: every parameter that requires a gradient
must receive one on every step (true for the reference's models), otherwise its bias correction would
run ahead of torch.optim.Adam's per-parameter count — ``step`` raises in that case."""
from __future__ import annotations

import torch

from . import _lib


def _bump_version(p: torch.Tensor) -> None:
    """The kernel writes through data_ptr(): tell autograd the tensor changed in place, so a backward
    that still holds it as a saved tensor raises instead of using the updated values."""
    try:
        torch.autograd.graph.increment_version(p)
    except AttributeError:        # very old torch: fall back to a no-op in-place op
        p.add_(0)


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 capturable: bool = False):
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1 and 0 <= betas[1] < 1):
            raise ValueError('FusedAdam: invalid hyper-parameter')
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.capturable = capturable
        self._step_dev = {}          # device -> int64 step counter shared by all tensors of the optimiser

    def _device_step(self, device) -> torch.Tensor:
        if device not in self._step_dev:
            self._step_dev[device] = torch.zeros((), dtype=torch.int64, device=device)
        return self._step_dev[device]

    def reset_state(self) -> None:
        """Back to step 0 with zero moments, keeping every tensor's address (graph capture warm-up)."""
        for st in self.state.values():
            if st:
                st['step'] = 0
                st['exp_avg'].zero_()
                st['exp_avg_sq'].zero_()
        for t in self._step_dev.values():
            t.zero_()

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        if self.capturable:
            missing = [p for group in self.param_groups for p in group['params'] if p.requires_grad and p.grad is None]
            if missing:
                raise _lib.EngineError('FusedAdam(capturable=True): a parameter that requires grad has none on this step '
                                       '(the device step counter is shared by all tensors)')
            devices = {p.device for group in self.param_groups for p in group['params'] if p.grad is not None}
            for d in devices:
                self._device_step(d).add_(1)
        for group in self.param_groups:
            b1, b2 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise _lib.EngineError('FusedAdam: CUDA fp32 parameters only (no CPU path)')
                if not p.is_contiguous():
                    raise _lib.EngineError('FusedAdam: contiguous parameters only')
                st = self.state[p]
                if not st:
                    st['step'] = 0
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st['step'] += 1     # (host copy; during graph replay only the device counter advances)
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                if self.capturable:
                    with torch.cuda.device(p.device):
                        rc = lib.rgcn_adam_step_dev(p.data_ptr(), g.data_ptr(), st['exp_avg'].data_ptr(),
                                                    st['exp_avg_sq'].data_ptr(), p.numel(), group['lr'], b1, b2,
                                                    group['eps'], group['weight_decay'],
                                                    self._device_step(p.device).data_ptr(),
                                                    torch.cuda.current_stream(p.device).cuda_stream)
                    _lib.check(rc, 'rgcn_adam_step_dev')
                    _bump_version(p)
                    continue
                with torch.cuda.device(p.device):
                    rc = lib.rgcn_adam_step(p.data_ptr(), g.data_ptr(), st['exp_avg'].data_ptr(),
                                            st['exp_avg_sq'].data_ptr(), p.numel(), group['lr'], b1, b2, group['eps'],
                                            group['weight_decay'], st['step'],
                                            torch.cuda.current_stream(p.device).cuda_stream)
                _lib.check(rc, 'rgcn_adam_step')
                _bump_version(p)
        return loss
