"""`Adam` (optimizers.py:7-39): Adam that leaves elements with exactly-zero gradient untouched."""
from __future__ import annotations

import torch


class Adam(torch.optim.Adam):
    @torch.no_grad()
    def step(self, closure=None):
        # optimizers.py:19-34: moments and parameters are updated only where grad != 0
        before = []
        for group in self.param_groups:
            for p in group['params']:
                if p.grad is None:
                    continue
                state = self.state.get(p, {})
                before.append((p, p.detach().clone(), p.grad != 0,
                               {k: v.detach().clone() for k, v in state.items() if torch.is_tensor(v) and v.shape == p.shape}))
        loss = super().step(closure)
        for p, old, mask, old_state in before:
            p.copy_(torch.where(mask, p, old))
            for k, v in old_state.items():
                self.state[p][k].copy_(torch.where(mask, self.state[p][k], v))
        return loss
