"""Stand-in for ema_pytorch.EMA (only what the reference trainer touches)."""
import copy
import torch
from torch import nn


class EMA(nn.Module):
    def __init__(self, model, beta=0.9999, update_after_step=100, update_every=10, **_):
        super().__init__()
        self.online_model = [model]
        self.ema_model = copy.deepcopy(model).requires_grad_(False)
        self.beta, self.update_after_step, self.update_every = beta, update_after_step, update_every
        self.register_buffer("step", torch.zeros((), dtype=torch.long))
        self.register_buffer("initted", torch.zeros((), dtype=torch.bool))

    @torch.no_grad()
    def update(self):
        step = int(self.step.item())
        self.step += 1
        if step % self.update_every != 0:
            return
        src = self.online_model[0]
        if step <= self.update_after_step or not bool(self.initted.item()):
            self.ema_model.load_state_dict(src.state_dict())
            self.initted.fill_(True)
            return
        for pe, po in zip(self.ema_model.parameters(), src.parameters()):
            pe.lerp_(po.detach(), 1.0 - self.beta)
        for be, bo in zip(self.ema_model.buffers(), src.buffers()):
            be.copy_(bo)

    def forward(self, *a, **k):
        return self.ema_model(*a, **k)
