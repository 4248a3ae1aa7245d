"""CPU restatement of the networks on the update path (oracle; TEST INFRASTRUCTURE ONLY).

Follows /root/reference/core/networks.py: ``base_network`` :65-92, ``PointNetFeature`` :182-250,
``QNetwork`` :253-300, ``GaussianPolicy`` :303-377, ``weights_init_`` :100-103.  Parameter names and
construction order are the reference's, so (a) reference checkpoints load and (b) the same
``torch.manual_seed`` gives the same initial weights (pinned by oracle/make_golden.py against the
unmodified reference modules).
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from .pointnet2_ops_cpu import pointnet2_modules as pn2

LOG_SIG_MAX, LOG_SIG_MIN, EPS = 2, -10, 1e-6
ACTION_HIGH = np.array([0.06, 0.06, 0.06, np.pi / 6, np.pi / 6, np.pi / 6])  # utils.py:505-510
ACTION_LOW = -ACTION_HIGH


def make_encoder(in_features, nclusters=32, radius=0.02, scale=1):
    """networks.py:65-92 — three set-abstraction modules, then Linear/BN1d/ReLU x2."""
    sa = nn.ModuleList([
        pn2.PointnetSAModule(npoint=nclusters, radius=radius, nsample=64,
                             mlp=[in_features, 64 * scale, 64 * scale, 128 * scale]),
        pn2.PointnetSAModule(npoint=32, radius=0.04, nsample=128,
                             mlp=[128 * scale, 128 * scale, 128 * scale, 256 * scale]),
        pn2.PointnetSAModule(mlp=[256 * scale, 256 * scale, 256 * scale, 512 * scale]),
    ])
    fc = nn.Sequential(
        nn.Linear(512 * scale, 1024 * scale), nn.BatchNorm1d(1024 * scale), nn.ReLU(True),
        nn.Linear(1024 * scale, 512 * scale), nn.BatchNorm1d(512 * scale), nn.ReLU(True),
    )
    return nn.ModuleList([sa, fc])


def burn_goal_feature_rng():
    """The driver builds a GoalFeature net before the state extractor (model-spec order,
    utils.py:188-201; networks.py:150-167); it is unused by update() but consumes the RNG."""
    make_encoder(3, nclusters=128)
    nn.Linear(512, 4), nn.Linear(512, 3), nn.Linear(512, 1)


class PointFeature(nn.Module):
    """networks.py:182-250."""

    def __init__(self, extra_latent=1, policy_extra_latent=-1, critic_extra_latent=-1, action_concat=True):
        super().__init__()
        self.policy_input_dim = 3 + policy_extra_latent if policy_extra_latent > 0 else 3 + extra_latent
        self.encoder = make_encoder(self.policy_input_dim)
        self.critic_input_dim = 3 + critic_extra_latent if critic_extra_latent > 0 else self.policy_input_dim
        if action_concat:
            self.critic_input_dim = 10  # hard-coded, networks.py:206-207
        self.value_encoder = make_encoder(self.critic_input_dim)

    @staticmethod
    def encode(enc, xyz, feats):
        for sa in enc[0]:
            xyz, feats = sa(xyz, feats)
        return enc[1](feats.squeeze(-1))

    def forward(self, pc, value=False):
        x = pc
        if x.shape[-1] != 1024:  # hand points included -> drop the 6 leading columns (:234-235)
            x = x[..., 6:]
        c = self.critic_input_dim if value else self.policy_input_dim
        x = x[:, :c].contiguous()
        xyz = x.transpose(1, -1)[..., :3].contiguous()
        return self.encode(self.value_encoder if value else self.encoder, xyz, x)


def _xavier(m):
    if isinstance(m, nn.Linear):
        nn.init.xavier_uniform_(m.weight, gain=1)
        nn.init.constant_(m.bias, 0)


def _quat_head(x):
    return torch.cat((F.normalize(x[:, :4], p=2, dim=-1), x[:, 4:]), dim=-1)


class TwinQ(nn.Module):
    """networks.py:253-300 (num_actions = 0: the action enters through the point cloud channels)."""

    def __init__(self, num_inputs, hidden, extra_pred_dim):
        super().__init__()
        self.linear1 = nn.Linear(num_inputs, hidden)
        self.linear2 = nn.Linear(hidden, hidden)
        self.linear3 = nn.Linear(hidden, 1)
        self.extra_pred_dim = extra_pred_dim
        self.linear4 = nn.Linear(num_inputs, hidden)
        self.linear5 = nn.Linear(hidden, hidden)
        self.linear6 = nn.Linear(hidden, 1)
        if extra_pred_dim > 0:
            self.linear7 = nn.Linear(num_inputs, hidden)
            self.linear8 = nn.Linear(hidden, hidden)
            self.extra_pred = nn.Linear(hidden, extra_pred_dim)
        self.apply(_xavier)

    def forward(self, s):
        q1 = self.linear3(F.relu(self.linear2(F.relu(self.linear1(s)))))
        q2 = self.linear6(F.relu(self.linear5(F.relu(self.linear4(s)))))
        aux = None
        if self.extra_pred_dim:
            aux = self.extra_pred(F.relu(self.linear8(F.relu(self.linear7(s)))))
            if self.extra_pred_dim == 7:
                aux = _quat_head(aux)
        return q1, q2, aux


class Policy(nn.Module):
    """networks.py:303-377 (action_space = PandaTaskSpace6D)."""

    def __init__(self, num_inputs, num_actions, hidden, extra_pred_dim):
        super().__init__()
        self.linear1 = nn.Linear(num_inputs, hidden)
        self.linear2 = nn.Linear(hidden, hidden)
        self.extra_pred_dim = extra_pred_dim
        self.mean = nn.Linear(hidden, num_actions)
        self.extra_pred = nn.Linear(hidden, extra_pred_dim)
        self.log_std_linear = nn.Linear(hidden, num_actions)
        self.apply(_xavier)
        self.action_scale = torch.FloatTensor((ACTION_HIGH - ACTION_LOW) / 2.0)
        self.action_bias = torch.FloatTensor((ACTION_HIGH + ACTION_LOW) / 2.0)

    def forward(self, s):
        x = F.relu(self.linear2(F.relu(self.linear1(s))))
        mean = self.mean(x)
        extra = self.extra_pred(x)
        if self.extra_pred_dim == 7:
            extra = _quat_head(extra)
        log_std = torch.clamp(self.log_std_linear(x), min=LOG_SIG_MIN, max=LOG_SIG_MAX)
        return mean, log_std, extra

    def sample(self, s, eps=None):
        """Returns (tanh-mean action, log-prob, sampled action, aux) like :353-371.  ``eps`` replaces the
        N(0,1) draw of rsample (None = draw from the global generator, exactly as the reference does)."""
        mean, log_std, extra = self.forward(s)
        std = log_std.exp()
        if eps is None:
            eps = torch.randn_like(mean)  # Normal.rsample = mean + std * randn (same RNG consumption)
        x_t = mean + std * eps
        y_t = torch.tanh(x_t)
        action = y_t * self.action_scale + self.action_bias
        log_prob = -((x_t - mean) ** 2) / (2 * std ** 2) - log_std - np.log(np.sqrt(2 * np.pi))
        log_prob = log_prob - torch.log(self.action_scale * (1 - y_t.pow(2)) + EPS)
        log_prob = log_prob.sum(1, keepdim=True)
        mean = torch.tanh(mean) * self.action_scale + self.action_bias
        return mean, log_prob, action, extra
