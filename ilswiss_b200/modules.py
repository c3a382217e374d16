"""Minimal parameter containers with the reference networks' parameter names, order and init,
for users (bench.py, tests, smoke) that do not have the reference checkout on their path.
In a drop-in deployment the reference's own FlattenMlp / policies / MLPDisc objects are passed
to the trainers instead (rlkit/torch/common/networks.py:23-115, policies.py:130-307,
adv_irl/disc_models/simple_disc_models.py:8-48); these classes restate only what the
sampler/eval side needs (forward, get_actions)."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

LOG_SIG_MAX, LOG_SIG_MIN = 2, -20


def _fanin_init(t):
    bound = 1.0 / np.sqrt(t.size(0))   # pytorch_util.py:20-29 uses size[0]
    return t.data.uniform_(-bound, bound)


class Mlp(nn.Module):
    def __init__(self, hidden_sizes, output_size, input_size, init_w=3e-3, b_init_value=0.1, output_activation=None):
        super().__init__()
        assert len(hidden_sizes) == 2 and hidden_sizes[0] == hidden_sizes[1], "fused engine: two equal hidden layers"
        self.fc0 = nn.Linear(input_size, hidden_sizes[0])
        self.fc1 = nn.Linear(hidden_sizes[0], hidden_sizes[1])
        for fc in (self.fc0, self.fc1):
            _fanin_init(fc.weight)
            fc.bias.data.fill_(b_init_value)
        self.last_fc = nn.Linear(hidden_sizes[1], output_size)
        self.last_fc.weight.data.uniform_(-init_w, init_w)
        self.last_fc.bias.data.uniform_(-init_w, init_w)
        self.output_activation = output_activation

    def hidden(self, x):
        return F.relu(self.fc1(F.relu(self.fc0(x))))

    def forward(self, x):
        out = self.last_fc(self.hidden(x))
        return self.output_activation(out) if self.output_activation is not None else out


class FlattenMlp(Mlp):
    def forward(self, *inputs):
        return super().forward(torch.cat(inputs, dim=1))


class TanhGaussianPolicy(Mlp):
    """ReparamTanhMultivariateGaussianPolicy parameter layout (policies.py:191-243)."""

    def __init__(self, hidden_sizes, obs_dim, action_dim, init_w=1e-3):
        super().__init__(hidden_sizes, action_dim, obs_dim, init_w=init_w)
        self.last_fc_log_std = nn.Linear(hidden_sizes[-1], action_dim)
        self.last_fc_log_std.weight.data.uniform_(-init_w, init_w)
        self.last_fc_log_std.bias.data.uniform_(-init_w, init_w)

    def forward(self, obs, deterministic=False):
        h = self.hidden(obs)
        mean = self.last_fc(h)
        log_std = torch.clamp(self.last_fc_log_std(h), LOG_SIG_MIN, LOG_SIG_MAX)
        z = mean if deterministic else mean + torch.randn_like(mean) * torch.exp(log_std)
        return torch.tanh(z), mean, log_std

    @torch.no_grad()
    def get_actions(self, obs_np, deterministic=False):
        dev = self.fc0.weight.device
        return self.forward(torch.as_tensor(obs_np, dtype=torch.float32, device=dev), deterministic)[0].cpu().numpy()


class DeterministicNoisePolicy(Mlp):
    """MlpGaussianNoisePolicy parameter layout + noise attributes (policies.py:130-188)."""

    def __init__(self, hidden_sizes, obs_dim, action_dim, init_w=1e-3, policy_noise=0.1, policy_noise_clip=0.5, max_act=1.0):
        super().__init__(hidden_sizes, action_dim, obs_dim, init_w=init_w, output_activation=torch.tanh)
        self.noise, self.noise_clip, self.max_act = policy_noise, policy_noise_clip, max_act

    def forward(self, obs, deterministic=False):
        pre = self.last_fc(self.hidden(obs))
        a = self.max_act * torch.tanh(pre)
        if not deterministic:
            a = a + torch.clamp(self.noise * torch.randn_like(a), -self.noise_clip, self.noise_clip)
        return a, pre

    @torch.no_grad()
    def get_actions(self, obs_np, deterministic=False):
        dev = self.fc0.weight.device
        return self.forward(torch.as_tensor(obs_np, dtype=torch.float32, device=dev), deterministic)[0].cpu().numpy()


class MLPDisc(nn.Module):
    """simple_disc_models.py:8-48 with num_layer_blocks=2, use_bn=False; hid_act 'tanh' (the shipped yamls) or 'relu'."""

    def __init__(self, input_dim, hid_dim=128, clamp_magnitude=10.0, hid_act="tanh"):
        super().__init__()
        act = {"tanh": nn.Tanh, "relu": nn.ReLU}[hid_act]
        self.clamp_magnitude = clamp_magnitude
        self.mod_list = nn.ModuleList([nn.Linear(input_dim, hid_dim), act(), nn.Linear(hid_dim, hid_dim), act(),
                                       nn.Linear(hid_dim, 1)])
        self.model = nn.Sequential(*self.mod_list)

    def forward(self, x):
        return torch.clamp(self.model(x), -self.clamp_magnitude, self.clamp_magnitude)
