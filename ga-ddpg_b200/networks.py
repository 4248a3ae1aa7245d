"""``PointNetFeatureB200`` — the drop-in for ``core.networks.PointNetFeature`` (/root/reference/core/networks.py:
182-250): same constructor kwargs (utils.py:190-198), same sub-modules ``.encoder`` / ``.value_encoder`` (used to
build the encoder optimisers, utils.py:221-234, and the log statistics, agent.py:247-250), same ``state_dict``
keys, same ``forward(pc, grasp, concat_option, rotz, feature_2, train) -> (z[B,512], pc)`` contract
(agent.py:72-78) — but the whole encoder runs in the fused CUDA kernels, forward and backward.

``register()`` installs the class into the reference's ``core.networks`` namespace so a model-spec YAML can say
``class: PointNetFeatureB200`` (utils.py:200 resolves the name with getattr).

Autograd mode (this module's ``forward``) is the compatibility path that lets the reference's own ``Agent`` /
``DDPG`` classes run unmodified on top of the CUDA encoder.  The fast path is ``agent.DDPGB200``, which drives
the same engine directly with hand-written backward passes and no autograd graph.
"""
import torch
from torch import nn

from . import engine, nets


class _EncoderFn(torch.autograd.Function):
    """z = encoder(cloud, bc).  Differentiable w.r.t. the encoder parameters and the broadcast channels ``bc``."""

    @staticmethod
    def forward(ctx, owner, which, cloud, skip, Cp, bc, train, *params):
        ef = owner._flat(which)
        dev = cloud.device
        ws = owner._workspace(dev)
        B, _, Np = cloud.shape
        geom = engine.Geometry(B, Np - skip, dev).build(cloud, skip)
        caps = (geom.lv[0].cap, geom.lv[1].cap)
        ectx = engine.EncoderCtx(B, caps, engine.WIDTHS, dev)
        bcc = bc.detach().contiguous() if bc is not None else None
        feat = engine.encoder_forward(ws, ef, geom, cloud, skip, Cp, bcc, ectx, time=None, train=train)
        ctx.owner, ctx.which, ctx.ectx, ctx.caps = owner, which, ectx, caps
        ctx.has_bc = bc is not None
        return feat[:, :512].clone()

    @staticmethod
    def backward(ctx, dz):
        owner, ef = ctx.owner, ctx.owner._flat(ctx.which)
        dev = dz.device
        ws = owner._workspace(dev)
        B = dz.shape[0]
        sc = engine.BwdScratch(B, ctx.caps, engine.WIDTHS, dev)
        dbc = engine.encoder_backward(ws, ef, ctx.ectx, sc, want_dw=True, want_dbc=ctx.has_bc, accumulate=0,
                                      dfeat=dz.contiguous())
        # hand autograd copies: it accumulates them into .grad with its own semantics (None vs +=)
        grads = [ef.arena.gview_of(p).clone() for p in owner._param_list(ctx.which)]
        return (None, None, None, None, None, dbc.clone() if dbc is not None else None, None) + tuple(grads)


class PointNetFeatureB200(nn.Module):
    def __init__(self, input_dim=3, pointnet_nclusters=32, pointnet_radius=0.02, model_scale=1, extra_latent=0,
                 split_feature=False, policy_extra_latent=-1, critic_extra_latent=-1, action_concat=False):
        super().__init__()
        assert model_scale == 1 and pointnet_nclusters == 32, "fused kernels cover the reference configuration"
        self.input_dim = 3 + extra_latent
        self.split_feature = False
        self.action_concat = action_concat
        self.policy_input_dim = 3 + policy_extra_latent if policy_extra_latent > 0 else self.input_dim
        self.encoder = nets.make_encoder_params(self.policy_input_dim, pointnet_nclusters, pointnet_radius)
        self.critic_input_dim = 3 + critic_extra_latent if critic_extra_latent > 0 else self.policy_input_dim
        if action_concat:
            self.critic_input_dim = 10  # networks.py:206-207
        self.value_encoder = nets.make_encoder_params(self.critic_input_dim, pointnet_nclusters, pointnet_radius)
        self._flats, self._ws = {}, {}

    # ---- flat arenas are created lazily, once the module sits on its CUDA device -----------------------
    def _flat(self, which):
        mod = self.value_encoder if which == "value" else self.encoder
        dev = next(mod.parameters()).device
        key = (which, str(dev))
        if key not in self._flats:
            if dev.type != "cuda":
                raise RuntimeError("PointNetFeatureB200 runs on CUDA only (no CPU fallback)")
            ef = engine.EncoderFlat(mod, dev)
            for prm in mod.parameters():
                prm.grad = None  # autograd owns .grad in plug-in mode; the arena's grad region is kernel scratch
            ef.arena.gview_of = lambda p, A=ef.arena: A.g[(p.data_ptr() - A.p.data_ptr()) // 4: (p.data_ptr() - A.p.data_ptr()) // 4 + p.numel()].view(p.shape)
            self._flats[key] = ef
        return self._flats[key]

    def _workspace(self, dev):
        k = str(dev)
        if k not in self._ws:
            self._ws[k] = engine.Workspace(dev)
        return self._ws[k]

    def _param_list(self, which):
        return list((self.value_encoder if which == "value" else self.encoder).parameters())

    def refresh(self, which=None):
        """Call after parameters changed outside the engine (optimizer.step(), load_state_dict)."""
        for (w, _), ef in self._flats.items():
            if which is None or w == which:
                ef.refresh_derived()

    def split_input(self, pc, feature_2):
        """Reference slicing (networks.py:232-242) expressed as (cloud, skip, per-point channels, broadcast channels).
        With ``action_concat`` the last 6 rows of ``pc`` are the action broadcast over the points
        (concat_state_action_channelwise, utils.py:291-297): they are constant per sample, so they are passed as
        a (B, Cb) matrix instead of being gathered per point."""
        skip = 6 if pc.shape[-1] != 1024 else 0
        c_use = self.critic_input_dim if feature_2 else self.policy_input_dim
        n_cloud = pc.shape[1] - 6 if (feature_2 and self.action_concat and pc.shape[1] > 6) else pc.shape[1]
        Cp = min(n_cloud, c_use)
        Cb = c_use - Cp
        bc = pc[:, n_cloud: n_cloud + Cb, skip] if Cb > 0 else None
        return skip, Cp, bc

    def forward(self, pc, grasp=None, concat_option="channel_wise", rotz=True, feature_2=False, train=True):
        if not pc.is_cuda:
            raise RuntimeError("PointNetFeatureB200: CPU tensors are not supported (no CPU fallback)")
        which = "value" if feature_2 else "policy"
        skip, Cp, bc = self.split_input(pc, feature_2)
        cloud = pc.detach()
        if cloud.dtype != torch.float32 or not cloud.is_contiguous():
            cloud = cloud.float().contiguous()
        self._flat(which).refresh_derived()  # parameters may have been stepped by a torch optimiser since last call
        z = _EncoderFn.apply(self, which, cloud, skip, Cp, bc, self.training, *self._param_list(which))
        return z, pc


def register():
    """Make ``class: PointNetFeatureB200`` resolvable by the reference's make_nets_opts_schedulers (utils.py:200)."""
    import sys

    mod = sys.modules.get("core.networks")
    if mod is None:
        import importlib

        mod = importlib.import_module("core.networks")
    mod.PointNetFeatureB200 = PointNetFeatureB200
    return mod
