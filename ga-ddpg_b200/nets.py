"""Parameter containers with the reference's module tree, names, shapes and initialisation, stored in flat
fp32 arenas that the CUDA kernels read and write directly.

Mirrors /root/reference/core/networks.py: ``base_network`` :65-92 (three upstream ``PointnetSAModule`` +
FC/BN1d head), ``PointNetFeature`` :182-215, ``QNetwork`` :253-278, ``GaussianPolicy`` :303-337.  The torch
layers built here are never *called*: they exist so that (a) ``state_dict()`` keys/shapes equal the reference's
(``encoder.0.0.mlps.0.0.weight`` ..., SURVEY.md §5 "Checkpoint / resume"), (b) the same ``torch.manual_seed``
reproduces the reference's initial weights (same constructors in the same order) and (c) ``.parameters()``
feeds optimisers.  ``Arena.adopt`` then re-points every parameter's ``.data``/``.grad`` at slices of one
contiguous buffer (params | grads | Adam m | Adam v), so a whole network is one Adam launch and one NCCL
all-reduce range.
"""
import numpy as np
import torch
from torch import nn

ALIGN = 64  # floats; keeps every tensor 256-byte aligned inside an arena

ACTION_HIGH = np.array([0.06, 0.06, 0.06, np.pi / 6, np.pi / 6, np.pi / 6])  # PandaTaskSpace6D, utils.py:505-510
ACTION_LOW = -ACTION_HIGH


def _shared_mlp(spec):
    layers = []
    for i in range(1, len(spec)):
        layers += [nn.Conv2d(spec[i - 1], spec[i], kernel_size=1, bias=False), nn.BatchNorm2d(spec[i]), nn.ReLU(True)]
    return nn.Sequential(*layers)


class SAModuleParams(nn.Module):
    """Same attribute layout as upstream PointnetSAModule (``mlps.0.{0,1,3,4,6,7}``); ``groupers`` hold no state."""

    def __init__(self, mlp, npoint=None, radius=None, nsample=None):
        super().__init__()
        spec = list(mlp)
        spec[0] += 3  # use_xyz
        self.npoint, self.radius, self.nsample = npoint, radius, nsample
        self.groupers = nn.ModuleList([nn.Identity()])
        self.mlps = nn.ModuleList([_shared_mlp(spec)])
        self.widths = spec

    def forward(self, *a, **k):
        raise RuntimeError("SAModuleParams is a parameter container; the fused CUDA encoder runs it")


def make_encoder_params(in_features, nclusters=32, radius=0.02, scale=1):
    sa = nn.ModuleList([
        SAModuleParams([in_features, 64 * scale, 64 * scale, 128 * scale], npoint=nclusters, radius=radius, nsample=64),
        SAModuleParams([128 * scale, 128 * scale, 128 * scale, 256 * scale], npoint=32, radius=0.04, nsample=128),
        SAModuleParams([256 * scale, 256 * scale, 256 * scale, 512 * scale]),
    ])
    fc = nn.Sequential(
        nn.Linear(512 * scale, 1024 * scale), nn.BatchNorm1d(1024 * scale), nn.ReLU(True),
        nn.Linear(1024 * scale, 512 * scale), nn.BatchNorm1d(512 * scale), nn.ReLU(True),
    )
    return nn.ModuleList([sa, fc])


def burn_goal_feature_rng():
    """The reference builds a GoalFeature net before the state extractor (utils.py:188-201, networks.py:150-167);
    it never runs in update(), but constructing it consumes the RNG — do the same for seed parity."""
    make_encoder_params(3, nclusters=128)
    nn.Linear(512, 4), nn.Linear(512, 3), nn.Linear(512, 1)


def _xavier(m):
    if isinstance(m, nn.Linear):
        nn.init.xavier_uniform_(m.weight, gain=1)
        nn.init.constant_(m.bias, 0)


class QNetworkParams(nn.Module):
    """networks.py:253-278 with num_actions = 0."""

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


class GaussianPolicyParams(nn.Module):
    """networks.py:303-337."""

    def __init__(self, num_inputs, num_actions, hidden, extra_pred_dim):
        super().__init__()
        self.linear1 = nn.Linear(num_inputs, hidden)
        self.linear2 = nn.Linear(hidden, hidden)
        self.extra_pred_dim = extra_pred_dim
        self.mean = nn.Linear(hidden, num_actions)
        self.extra_pred = nn.Linear(hidden, extra_pred_dim)
        self.log_std_linear = nn.Linear(hidden, num_actions)
        self.apply(_xavier)


class GradPool:
    """One contiguous fp32 buffer that the gradient regions of SEVERAL arenas are carved from, in construction order, so
    that everything one optimiser phase reduces across GPUs is a single range (value encoder + critic | policy encoder +
    policy): one NCCL all-reduce per phase instead of one per network."""

    def __init__(self, capacity, device):
        self.buf = torch.zeros(capacity, dtype=torch.float32, device=device)
        self.off = 0

    def take(self, n):
        assert self.off + n <= self.buf.numel(), "GradPool too small"
        t = self.buf[self.off: self.off + n]
        self.off += n
        return t

    @property
    def used(self):
        return self.buf[: self.off]

    @staticmethod
    def capacity_for(*modules):
        """Upper bound of the arena sizes of ``modules`` (every tensor padded to ALIGN, plus the arena tail)."""
        return sum(sum((p.numel() + ALIGN) for p in m.parameters()) + ALIGN for m in modules)


class Arena:
    """One contiguous fp32 buffer [params | grads | m | v] (+ optional Polyak target copy elsewhere); with ``grad_pool`` the
    gradient region lives in that shared pool instead ([params | m | v] here).

    ``order`` is a list of (name, tensor) in the order they should be laid out; tensors that must be adjacent
    for stacked GEMMs (e.g. linear1/4/7 of the critic) are simply listed next to each other with ``pack=True``
    rows so no alignment gap is inserted between them.
    """

    def __init__(self, order, device, with_opt=True, grad_pool=None):
        self.offsets, self.sizes = {}, {}
        off = 0
        for name, t, pack in order:
            if not pack:
                off = (off + ALIGN - 1) // ALIGN * ALIGN
            self.offsets[name], self.sizes[name] = off, t.numel()
            off += t.numel()
        self.n = (off + ALIGN - 1) // ALIGN * ALIGN
        pooled = with_opt and grad_pool is not None
        k = (3 if pooled else 4) if with_opt else 1
        self.buf = torch.zeros(k * self.n, dtype=torch.float32, device=device)
        self.p = self.buf[: self.n]
        if pooled:
            self.g = grad_pool.take(self.n)
            self.m, self.v = self.buf[self.n: 2 * self.n], self.buf[2 * self.n:]
        else:
            self.g = self.buf[self.n: 2 * self.n] if with_opt else None
            self.m = self.buf[2 * self.n: 3 * self.n] if with_opt else None
            self.v = self.buf[3 * self.n:] if with_opt else None
        for name, t, _ in order:
            view = self.view(name, t.shape)
            view.copy_(t.detach().to(device))
            if isinstance(t, nn.Parameter):
                t.data = view
                if with_opt:
                    t.grad = self.view(name, t.shape, self.g)

    def view(self, name, shape, base=None):
        base = self.p if base is None else base
        o, n = self.offsets[name], self.sizes[name]
        return base[o: o + n].view(shape)

    def gview(self, name, shape):
        return self.view(name, shape, self.g)
