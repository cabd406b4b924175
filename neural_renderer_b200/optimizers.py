"""`Adam` (reference optimizers.py:9-39): Adam that (1) leaves every element whose gradient is exactly zero untouched --
parameter AND both moments -- and (2) scales the step of a parameter by its own `param.lr` multiplier when it has one
(`Mesh.set_lr`).  One masked, in-place update per parameter; no snapshots of parameters or state."""
from __future__ import annotations

import math

import torch


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=0.001, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()  # the zero-gradient mask below is taken from the gradients the closure produced
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            for p in group['params']:
                if p.grad is None:
                    continue
                state = self.state[p]
                if not state:
                    state['step'] = 0
                    state['m'] = torch.zeros_like(p)
                    state['v'] = torch.zeros_like(p)
                state['step'] += 1
                t = state['step']
                # chainer's Adam.lr: alpha * sqrt(1 - beta2^t) / (1 - beta1^t), times the parameter's own multiplier
                lr = group['lr'] * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t) * float(getattr(p, 'lr', 1.0))
                if lr == 0:
                    continue  # optimizers.py:18: a zero learning rate skips the parameter, moments included
                g, m, v = p.grad, state['m'], state['v']
                mask = g != 0
                m.add_(torch.where(mask, (1.0 - beta1) * (g - m), torch.zeros_like(g)))
                v.add_(torch.where(mask, (1.0 - beta2) * (g * g - v), torch.zeros_like(g)))
                v.clamp_(min=0)  # optimizers.py:28 (v >= 0 already holds for the untouched elements)
                p.sub_(torch.where(mask, lr * m / (v.sqrt() + group['eps']), torch.zeros_like(g)))
        return loss
