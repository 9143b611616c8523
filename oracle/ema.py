"""ema_pytorch.EMA(model, beta, update_every, include_online_model=False) restated from SURVEY
Appendix B2 (ema_pytorch is not installed); call sites d3f/train_deep_fake/lit_module.py:62-70,185,189.
TEST INFRASTRUCTURE ONLY."""
import copy
import torch
import torch.nn as nn


class EMA(nn.Module):
    def __init__(self, model, beta=0.9999, update_after_step=100, update_every=10, inv_gamma=1.0, power=2 / 3,
                 min_value=0.0, include_online_model=True):
        super().__init__()
        self.beta, self.update_after_step, self.update_every = beta, update_after_step, update_every
        self.inv_gamma, self.power, self.min_value = inv_gamma, power, min_value
        if include_online_model:
            self.online_model = model
        else:
            self.online_model = [model]
        self.ema_model = copy.deepcopy(model)
        self.ema_model.requires_grad_(False)
        self.register_buffer("initted", torch.tensor(False))
        self.register_buffer("step", torch.tensor(0))

    @property
    def model(self):
        return self.online_model if isinstance(self.online_model, nn.Module) else self.online_model[0]

    def get_current_decay(self):
        epoch = max(int(self.step.item()) - self.update_after_step - 1, 0)
        value = 1 - (1 + epoch / self.inv_gamma) ** -self.power
        if epoch <= 0:
            return 0.0
        return min(max(value, self.min_value), self.beta)

    @torch.no_grad()
    def copy_params_from_model_to_ema(self):
        for (_, e), (_, m) in zip(self.ema_model.named_parameters(), self.model.named_parameters()):
            e.copy_(m)
        for (_, e), (_, m) in zip(self.ema_model.named_buffers(), self.model.named_buffers()):
            e.copy_(m)

    @torch.no_grad()
    def update(self):
        step = int(self.step.item())
        self.step += 1
        if step % self.update_every != 0:
            return
        if step <= self.update_after_step:
            self.copy_params_from_model_to_ema()
            return
        if not bool(self.initted.item()):
            self.copy_params_from_model_to_ema()
            self.initted.fill_(True)
        decay = self.get_current_decay()
        for (_, e), (_, m) in zip(self.ema_model.named_parameters(), self.model.named_parameters()):
            if e.is_floating_point():
                e.lerp_(m, 1 - decay)
        for (_, e), (_, m) in zip(self.ema_model.named_buffers(), self.model.named_buffers()):
            if e.is_floating_point():
                e.lerp_(m, 1 - decay)

    def forward(self, *a, **k):
        return self.ema_model(*a, **k)
