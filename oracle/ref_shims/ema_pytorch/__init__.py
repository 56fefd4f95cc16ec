"""Stand-in for ema_pytorch.EMA (only what the reference trainer touches): update schedule, decay warm-up
(get_current_decay with inv_gamma=1, power=2/3, min_value=0) and the state_dict schema ('ema_model.*', 'initted',
'step'), restated from the published package (not installed here)."""
import copy
import torch
from torch import nn


class EMA(nn.Module):
    def __init__(self, model, beta=0.9999, update_after_step=100, update_every=10, inv_gamma=1.0, power=2 / 3,
                 min_value=0.0, **_):
        super().__init__()
        self.online_model = [model]
        self.ema_model = copy.deepcopy(model).requires_grad_(False)
        self.beta, self.update_after_step, self.update_every = beta, update_after_step, update_every
        self.inv_gamma, self.power, self.min_value = inv_gamma, power, min_value
        self.register_buffer("initted", torch.tensor(False))
        self.register_buffer("step", torch.tensor(0))

    def get_current_decay(self):
        epoch = max(int(self.step.item()) - self.update_after_step - 1, 0)
        if epoch <= 0:
            return 0.0
        return min(max(1.0 - (1.0 + epoch / self.inv_gamma) ** -self.power, self.min_value), self.beta)

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
        decay = self.get_current_decay()
        for pe, po in zip(self.ema_model.parameters(), src.parameters()):
            pe.lerp_(po.detach(), 1.0 - decay)
        for be, bo in zip(self.ema_model.buffers(), src.buffers()):
            be.copy_(bo)

    def forward(self, *a, **k):
        return self.ema_model(*a, **k)
