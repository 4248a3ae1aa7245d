"""Fused B200 agents: ``DDPGB200`` / ``BCB200`` — drop-ins for the reference's ``DDPG`` / ``BC`` classes.

Same public surface as /root/reference/core/agent.py + ddpg.py + bc.py: constructor
``(num_inputs, action_space, CONFIG)``, ``setup_feature_extractor(net_dict)``, ``update_parameters(batch_data,
updates, k) -> dict[str, float]`` with the 11 keys of get_loss_info_dict (utils.py:1008-1020),
``select_action(state, remain_timestep=...)``, ``step_scheduler``, ``get_lr``, ``update_step``,
``get_weight/load_weight`` — but ``update_parameters`` is one fixed sequence of hand-written CUDA kernels:
no autograd graph, no allocation, no host synchronisation except the single read-back of the scalars.

Step structure (ddpg.py:146-185; F = encoder forward, B = backward):
  phase 1   state chain: geometry(state); F1 value(state,a); F4 policy(state); critic(F1)   || on a second stream,
            target chain: geometry(next); F2 policy(next) -> policy_target -> TD3 noise; F3 value(next,a') ->
            critic_target -> y;  then losses; B1 through critic + value encoder (weight gradients on a third stream)
  [all-reduce of the value-encoder + critic gradient range when sample-sharded over several GPUs]
  phase 2   clip_grad_norm(critic), Adam(value encoder), Adam(critic); policy(F4) -> pi;
            even steps: F5 value(state,pi) -> critic -> -mix*mean(minQ), B through critic + value encoder to dpi;
            actor losses; B2 through policy + policy encoder
  [all-reduce of the policy-encoder + policy gradient range]
  phase 3   Adam(policy), Adam(policy encoder), Polyak targets, statistics
Each phase is captured once per step parity into a CUDA graph and replayed.
"""
import contextlib
import math
import os
from types import SimpleNamespace as NS

import numpy as np
import torch
from torch.optim.lr_scheduler import MultiStepLR

from . import checkpoint, engine, nets
from .capi import current_stream, lib
from .config import DEFAULTS, LOSS_KEYS
from .engine import FEAT_LD, QA_AUX, QA_LD, QA_Q2, dp
from .networks import PointNetFeatureB200

RING = 2   # host staging slots (see AgentB200._begin_step)
O_CRITIC, O_CRITIC_AUX, O_NGOAL_C, O_BC, O_POLICY_AUX, O_NGOAL_A, O_AC, O_PPARAM, O_CGRAD, O_CPARAM, O_CLIP, O_GNORM = range(12)


class _Sched:
    """MultiStepLR on a parameter-less optimiser: keeps torch's own schedule arithmetic for the lr."""

    def __init__(self, lr, milestones, gamma):
        self.opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=lr)
        self.sched = MultiStepLR(self.opt, milestones=list(milestones), gamma=gamma)
        self.opt._opt_called = True  # silence the "scheduler before optimizer" warning; our Adam is fused

    def step(self):
        self.sched.step()

    @property
    def lr(self):
        return self.opt.param_groups[0]["lr"]


def _on_device(fn):
    """Run a public entry point with the agent's device current: every launch goes to ``torch.cuda.current_stream()`` of
    the CURRENT device, so an agent built with ``device="cuda:1"`` must not depend on what the caller left selected."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        if torch.cuda.current_device() == self.device.index:
            return fn(self, *a, **k)
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)

    return wrapped


class PendingResult:
    """Handle returned by ``update_parameters(..., defer=True)``: the 11 scalars of a step whose kernels may still be
    running.  ``result()`` waits for that step only — the caller can enqueue the next step first (feed.FeedLoop)."""

    def __init__(self, agent, slot, event):
        self.agent, self.slot, self.event, self._r = agent, slot, event, None

    def result(self):
        if self._r is None:
            self.event.synchronize()
            self._r = self.agent._scalars(self.agent.out_host[self.slot])
        return self._r


def _sched_args(entry, opt_key, sch_key, lr, milestones, gamma, eps=1e-8, wd=0.0):
    """Learning-rate schedule / Adam constants of one optimiser of the reference's net_dict (utils.py:211-234) when the
    caller built them (model-spec ``opt_kwargs`` / ``scheduler_kwargs``), else the configured defaults."""
    opt, sch = (entry or {}).get(opt_key), (entry or {}).get(sch_key)
    if opt is not None and getattr(opt, "param_groups", None):
        g = opt.param_groups[0]
        lr, eps, wd = g.get("initial_lr", g["lr"]), g.get("eps", eps), g.get("weight_decay", wd)
    if sch is not None and hasattr(sch, "milestones"):
        milestones, gamma = sorted(sch.milestones.elements()), sch.gamma
    return float(lr), list(milestones), float(gamma), float(eps), float(wd)


def _cfg_get(cfg, k):
    if cfg is not None:
        try:
            if k in cfg:
                return cfg[k]
        except TypeError:
            pass
        if hasattr(cfg, k):
            return getattr(cfg, k)
    return DEFAULTS[k]


class AgentB200:
    name = "Agent"

    def __init__(self, num_inputs=512, action_space=None, args=None, device=None, seed=None, world=None):
        for k in DEFAULTS:
            setattr(self, k, _cfg_get(args, k))
        if args is not None:
            try:
                for k, v in args.items():
                    setattr(self, k, v)
            except AttributeError:
                pass
        if not torch.cuda.is_available():
            raise RuntimeError("gaddpg_b200 agents need a CUDA device: there is no CPU fallback")
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.has_critic = self.name != "BC"
        self.update_step = 1
        self.init_step = 1
        self.extra_pred_dim = 7 if self.policy_aux else 1
        self.critic_extra_pred_dim = 7 if self.critic_aux else 0
        self.num_inputs = num_inputs + (1 if self.use_time else 0)
        assert self.num_inputs == 513 and self.hidden_size == 256 and self.use_time, "fused kernels cover the reference widths"
        assert self.sa_channel_concat and self.noise_type == "uniform" and self.use_action_limit
        self.world = world  # None or a torch.distributed process group handle (see dist.py)
        if seed is not None:
            torch.manual_seed(seed)
        self._own_extractor = None
        self._built = False
        self._pending_seeded_build = seed is not None
        # policy / target policy are created here, like Agent.__init__ (agent.py:21-48, utils.py:960-981) — but only
        # after the feature extractor exists when we own the seed, to keep the reference's RNG order
        if not self._pending_seeded_build:
            self._make_heads()

    # ---- construction ----------------------------------------------------------------------------------
    def _make_heads(self):
        self.policy = nets.GaussianPolicyParams(self.num_inputs, 6, self.hidden_size, self.extra_pred_dim)
        self.policy_target = nets.GaussianPolicyParams(self.num_inputs, 6, self.hidden_size, self.extra_pred_dim)
        if self.has_critic:
            self.critic = nets.QNetworkParams(self.num_inputs, self.hidden_size, self.critic_extra_pred_dim)
            self.critic_target = nets.QNetworkParams(self.num_inputs, self.hidden_size, self.critic_extra_pred_dim)

    def build_standalone(self):
        """Create the feature extractor the way the reference driver does (model-spec order: goal extractor first,
        utils.py:188-201) and then the heads, so ``seed`` reproduces the reference's initial weights."""
        nets.burn_goal_feature_rng()
        net = PointNetFeatureB200(input_dim=self.channel_num, extra_latent=self.extra_latent,
                                  policy_extra_latent=self.policy_extra_latent, critic_extra_latent=self.critic_extra_latent,
                                  action_concat=self.sa_channel_concat)
        self._make_heads()
        self._pending_seeded_build = False
        self.setup_feature_extractor({"state_feature_extractor": {"net": net}, "goal_feature_extractor": {"net": None}})
        return self

    def setup_feature_extractor(self, net_dict, eval=False):
        """agent.py:149-164.  Accepts the reference's net_dict (net possibly wrapped in nn.DataParallel).  The torch
        optimisers in it are not stepped — Adam runs fused on the arenas — but their hyper-parameters (lr, eps,
        weight_decay) and their schedulers' milestones / gamma are read from it when present (model-spec ``opt_kwargs`` /
        ``scheduler_kwargs``, utils.py:211-234); otherwise the configured defaults apply."""
        if self._pending_seeded_build:
            self._make_heads()
            self._pending_seeded_build = False
        sfe = net_dict["state_feature_extractor"]
        net = sfe["net"]
        self.state_feature_extractor = net
        self._extractor = net.module if hasattr(net, "module") else net
        assert isinstance(self._extractor, PointNetFeatureB200), "model_spec must select class: PointNetFeatureB200"
        self.goal_feature_extractor = net_dict.get("goal_feature_extractor", {}).get("net")
        c = self
        fm = c.overwrite_feat_milestone or c.feat_milestones
        e_lr, e_ms, e_g, e_eps, e_wd = _sched_args(sfe, "encoder_opt", "encoder_scheduler", c.feat_lr, fm, c.feat_gamma)
        v_lr, v_ms, v_g, v_eps, v_wd = _sched_args(sfe, "val_encoder_opt", "val_encoder_scheduler", c.feat_lr, fm, c.feat_gamma)
        self._adam_hp = dict(enc=(e_eps, e_wd), venc=(v_eps, v_wd), policy=(1e-5, 1e-5), critic=(1e-5, 1e-5))  # utils.py:969,993
        self._sch = NS(
            policy=_Sched(c.lr, c.policy_milestones, c.lr_gamma),
            critic=_Sched(c.value_lr, c.value_milestones, c.value_lr_gamma) if self.has_critic else None,
            enc=_Sched(e_lr, e_ms, e_g),
            venc=_Sched(v_lr, v_ms, v_g),  # never stepped (agent.py:179-190)
        )
        dev = self.device
        engine._init_once(dev)
        with torch.cuda.device(dev):
            self._setup_device_state(dev)

    def _setup_device_state(self, dev):
        self.ws = engine.Workspace(dev)
        # stream-level overlap inside a step (identical arithmetic, see _phase1): a second encoder chain and the
        # weight-gradient products run on side streams; ``overlap = False`` issues everything on one stream
        self.overlap = True
        # F4 (policy encoder on the state cloud) beside the critic backward instead of beside the target chain (see _phase1_critic);
        # GADDPG_F4_LATE=0/1 for A/B
        # measured at cfg2 (A/B/A/B in one session): 218.3 / 219.0 steps/s late vs 222.1 / 222.4 early -> off
        self.f4_late = os.environ.get("GADDPG_F4_LATE", "0") == "1"
        # F4 as a THIRD concurrent chain on the (idle in phase 1) weight-gradient stream: 226.1 / 225.7 vs 229.6 / 229.6 steps/s -> off
        self.f4_par = os.environ.get("GADDPG_F4_PAR", "0") == "1"
        # geometry is launched before the small-field staging for device-resident batches (+1 % steps/s: the GPU no longer
        # idles behind ~100 us of host work); for host batches that ordering measured 0.17 ms SLOWER per step (5.59 vs
        # 5.42 ms, A/B in one process), so they keep transfers -> staging -> geometry
        self.early_geometry_host = False
        self.side_enc = engine.SideStream(dev)
        self.side_dw = engine.SideStream(dev)
        # one contiguous gradient range per optimiser phase: [policy encoder | policy] and [value encoder | critic]
        ex = self._extractor
        self.gpool_a = nets.GradPool(nets.GradPool.capacity_for(ex.encoder, self.policy), dev)
        self.gpool_c = nets.GradPool(nets.GradPool.capacity_for(ex.value_encoder, self.critic), dev) if self.has_critic else None
        self.ef_p = engine.EncoderFlat(self._extractor.encoder, dev, grad_pool=self.gpool_a)
        self.ef_v = engine.EncoderFlat(self._extractor.value_encoder, dev, grad_pool=self.gpool_c)
        self._extractor._flats[("policy", str(dev))] = self.ef_p
        self._extractor._flats[("value", str(dev))] = self.ef_v
        # per-pass staging of the BatchNorm running statistics (F1, F3: value encoder | F2, F4: policy encoder)
        self.bnst = {1: engine.BNStage(self.ef_v, dev), 2: engine.BNStage(self.ef_p, dev), 3: engine.BNStage(self.ef_v, dev),
                     4: engine.BNStage(self.ef_p, dev)}
        self.pf = engine.PolicyFlat(self.policy, dev, grad_pool=self.gpool_a)
        self.pft = engine.PolicyFlat(self.policy_target, dev, with_opt=False)
        if self.has_critic:
            self.cf = engine.CriticFlat(self.critic, dev, grad_pool=self.gpool_c)
            self.cft = engine.CriticFlat(self.critic_target, dev, with_opt=False)
            self.tau_soft, self.tau_hard = self.cft.tau_vectors(self.tau)
        self._set_channels(3 + self.extra_latent)
        self.opt_steps = dict(policy=0, critic=0, enc=0, venc=0)
        # host staging is a ring of RING slots (pinned): a slot is rewritten only after the step that read it has finished
        # on the GPU, so ``update_parameters(..., defer=True)`` may run one step ahead of the device (feed.FeedLoop)
        self.dyn = torch.zeros(4, 2, dtype=torch.float32, device=dev)
        self.dyn_host = [torch.zeros(4, 2, dtype=torch.float32).pin_memory() for _ in range(RING)]
        self.out = torch.zeros(16, dtype=torch.float32, device=dev)
        self.out_host = [torch.zeros(16, dtype=torch.float32).pin_memory() for _ in range(RING)]
        self._slot_done = [torch.cuda.Event() for _ in range(RING)]
        self._slot = 0
        self._pending = [None] * RING
        self._optjobs = {}
        self._h2d_stream = None
        self._prep_pending = False
        # pipelined callers (defer=True) that pass dicts of DEVICE tensors: True = those tensors are complete when they are handed
        # over (e.g. pre-generated batches), so their staging copies + geometry may run on the prep streams beside the running step
        self.static_device_batches = False
        self._pending_reduce = []
        self.step_start_events = self.step_end_events = None    # feed.FeedLoop installs lists here to measure the GPU idle gap between steps
        # sharded runs: False = ONE all-reduce per optimiser phase over its contiguous gradient range (NCCL: captured inside the
        # whole-step graph); True = everything but SA1's gradients reduced asynchronously behind the SA1 backward (4 calls per step,
        # eager between graph segments).  Measured on 8 B200s: 1662.6 (single) vs 1647.1 (split) aggregate steps/s
        self.split_reduce = False
        self._shape = None
        self._graphs = {}
        self.use_graph = True
        self._in_outer = False
        # phases 1-3 of an update (and, in a sharded run with one all-reduce per phase, the two NCCL calls between them) as ONE
        # CUDA graph per (step parity, target-copy step, schedule index) instead of five; GADDPG_WHOLE_GRAPH=0/1 for A/B
        self.whole_graph = os.environ.get("GADDPG_WHOLE_GRAPH", "1") == "1"
        self._built = True
        self.policy.sample = self.policy_sample   # reference call style: agent.policy.sample(feat) (networks.py:353-371)

    def _set_channels(self, C):
        """Per-point / broadcast channel split of the two encoders for clouds with ``C`` channels — the same rule as the
        plug-in path (networks.PointNetFeatureB200.split_input, reference networks.py:232-242): the policy encoder takes
        the first policy_input_dim cloud channels, the value encoder min(C, critic_input_dim) cloud channels followed by
        the first critic_input_dim - that many action channels."""
        ex = self._extractor
        if C < ex.policy_input_dim:
            raise ValueError("clouds have %d channels, the policy encoder needs %d" % (C, ex.policy_input_dim))
        self.Cp_policy = ex.policy_input_dim
        self.Cp_value = min(C, ex.critic_input_dim)
        self.Cb_value = ex.critic_input_dim - self.Cp_value
        if self.Cb_value > 6:
            raise ValueError("value encoder would need %d action channels (critic_input_dim %d, cloud channels %d)"
                             % (self.Cb_value, ex.critic_input_dim, C))

    # ---- buffers sized for one (B, C, N) ------------------------------------------------------------------
    def _alloc(self, B, C, Np):
        dev = self.device
        self._set_channels(C)
        skip = 6 if Np != 1024 else 0
        N = Np - skip
        self._shape = (B, C, Np)
        self.B, self.skip, self.N = B, skip, N
        # The step's INPUTS (clouds, small fields, geometry) exist once per staging slot: a pipelined caller prepares step i+1 —
        # transfers / gather AND the xyz-only geometry kernels — on the prep streams while step i still computes from the
        # other slot (prepare_data(prefetch=True)).  ``self.cloud`` etc. are properties of the current slot.
        self._cloud_s = [torch.zeros(B, C, Np, dtype=torch.float32, device=dev) for _ in range(RING)]
        self._next_cloud_s = [torch.zeros(B, C, Np, dtype=torch.float32, device=dev) for _ in range(RING)]
        seg = lambda n: (n + 3) // 4 * 4  # noqa: E731
        names = [("action", 6), ("expert_action", 6), ("goal", 7), ("noise_u", 6), ("reward", 1), ("ret", 1), ("done", 1),
                 ("time", 1), ("expert_flag", 1), ("perturb_flag", 1)]
        total, self._vec_off = 0, {}
        for n, w in names:
            self._vec_off[n] = (total, w)
            total += seg(B * w)
        self._vec_s = [torch.zeros(total, dtype=torch.float32, device=dev) for _ in range(RING)]
        self.vec_host = [torch.zeros(total, dtype=torch.float32).pin_memory() for _ in range(RING)]
        self._cloud_shape = (B, C, Np)
        self._cloud_host = [None] * RING        # pinned cloud staging: allocated on first use (host-fed batches only)
        self._next_cloud_host = [None] * RING
        self._v_s = [NS(**{n: vec[o: o + B * w].view(B, w) if w > 1 else vec[o: o + B] for n, (o, w) in self._vec_off.items()})
                     for vec in self._vec_s]
        self.vh = [NS(**{n: vh[o: o + B * w].view(B, w) if w > 1 else vh[o: o + B] for n, (o, w) in self._vec_off.items()})
                   for vh in self.vec_host]
        self._geom_s_s = [engine.Geometry(B, N, dev) for _ in range(RING)]
        self._geom_n_s = [engine.Geometry(B, N, dev) for _ in range(RING)]
        caps = (self.geom_s.lv[0].cap, self.geom_s.lv[1].cap)
        mk = lambda: engine.EncoderCtx(B, caps, engine.WIDTHS, dev)  # noqa: E731
        self.ctx_p = mk()                      # F4 (policy encoder, state) — kept for B2
        if self.has_critic:
            self.ctx_v1, self.ctx_n, self.ctx_v5 = mk(), mk(), mk()   # F1 | F2,F3 (no grad) | F5
            self.cc1, self.cct, self.cc5 = (engine.critic_ctx(B, self.cf, dev) for _ in range(3))
            self.pct = engine.policy_ctx(B, self.pft, dev)
            self.next_action = torch.zeros(B, 6, dtype=torch.float32, device=dev)
            self.y = torch.zeros(B, dtype=torch.float32, device=dev)
        self.sc = engine.BwdScratch(B, caps, engine.WIDTHS, dev)
        self.pc = engine.policy_ctx(B, self.pf, dev)
        self.dpi_ac = torch.zeros(B, 6, dtype=torch.float32, device=dev)
        self._bc_buf = [torch.zeros(B, max(self.Cb_value, 1), dtype=torch.float32, device=dev) for _ in range(3)]
        self._graphs = {}

    # inputs of the current staging slot
    cloud = property(lambda self: self._cloud_s[self._slot])
    next_cloud = property(lambda self: self._next_cloud_s[self._slot])
    vec = property(lambda self: self._vec_s[self._slot])
    v = property(lambda self: self._v_s[self._slot])
    geom_s = property(lambda self: self._geom_s_s[self._slot])
    geom_n = property(lambda self: self._geom_n_s[self._slot])

    def _prep_streams(self):
        """Two streams that carry the input staging + geometry of the NEXT step of a pipelined caller (state side / next-state
        side), and one event per slot and side that the compute stream waits for before the step's first phase."""
        if self._h2d_stream is None:
            self._h2d_stream = (torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device))
            self._h2d_ev = [(torch.cuda.Event(), torch.cuda.Event()) for _ in range(RING)]
        return self._h2d_stream

    def _wait_prep(self):
        """Compute stream <- the prep streams of this slot (no-op for a synchronous prepare_data)."""
        if self._prep_pending:
            cur = torch.cuda.current_stream()
            for e in self._h2d_ev[self._slot]:
                cur.wait_event(e)
            self._prep_pending = False

    def _bc(self, action, slot):
        """The value encoder sees critic_input_dim - cloud_channels action channels (6 with the reference spec; the
        hard-coded critic_input_dim = 10 truncates the action when the cloud has more channels, networks.py:206-207,238)."""
        if self.Cb_value == action.shape[1]:
            return action
        self._bc_buf[slot].copy_(action[:, : self.Cb_value])
        return self._bc_buf[slot]

    # ---- data staging (agent.py:211-240 prepare_data) ------------------------------------------------------
    @_on_device
    def prepare_data(self, batch, noise_u=None, after_clouds=None, prefetch=False):
        """Stage one replay minibatch (the dict BaseMemory.sample returns, replay_memory.py:166-176) into the device
        input buffers.  Host batches go through pinned buffers (float64 clouds are converted on the host like
        torch.cuda.FloatTensor(ndarray) does in the reference); a lazy ``ReplayBatch`` of the device-resident replay is
        gathered straight into the buffers by one CUDA launch pair.  ``after_clouds`` is called as soon as the cloud
        transfers are issued (the agents launch the xyz-only geometry kernels there, so the GPU is busy while the host
        stages the small per-sample fields).  ``prefetch`` (pipelined callers: update_parameters(defer=True) / feed.FeedLoop):
        host clouds travel on a dedicated copy stream into per-slot device staging buffers — the DMA of minibatch i+1 runs while
        step i computes — and reach the kernels' input buffers by a device-to-device copy once both are done."""
        from .replay_memory import ReplayBatch

        lazy = isinstance(batch, ReplayBatch) and not batch.materialised
        if lazy:
            mem = batch.memory
            B, (C, Np) = len(batch.batch_idx), mem.row
            if self._shape != (B, C, Np):
                self._alloc(B, C, Np)
            pre = prefetch and after_clouds is not None
            with (torch.cuda.stream(self._prep_streams()[0]) if pre else contextlib.nullcontext()):
                mem.gather_into(batch.batch_idx, self.cloud, self.next_cloud if self.has_critic else None, self.vec, self._vec_off)
                mem._lazy.pop(id(batch), None)   # consumed: a later write to the buffer need not snapshot it
                if pre:     # gather + geometry of both clouds behind the running step; the compute stream joins in _wait_prep
                    self._geometry(prep=True)
                elif after_clouds is not None:
                    after_clouds()
                if self.has_critic:
                    self.v.noise_u.copy_(torch.rand(B, 6, device=self.device) if noise_u is None
                                         else torch.as_tensor(noise_u, dtype=torch.float32).view(B, 6), non_blocking=True)
                if pre:
                    for e in self._h2d_ev[self._slot]:
                        e.record()
                    self._prep_pending = True
            return
        cloud = batch["point_state_batch"]
        B, C, Np = cloud.shape
        if self._shape != (B, C, Np):
            self._alloc(B, C, Np)

        slot = self._slot

        def put_cloud(dst, ring, src):
            if torch.is_tensor(src) and src.dtype == torch.float32 and (src.is_cuda or src.is_pinned()):
                dst.copy_(src, non_blocking=True)       # already pinned fp32 (or device resident): straight DMA
            else:
                if ring[slot] is None:
                    ring[slot] = torch.zeros(self._cloud_shape, dtype=torch.float32).pin_memory()
                ring[slot].copy_(torch.as_tensor(src))   # float64 ndarray of the reference's replay buffer: convert on host
                dst.copy_(ring[slot], non_blocking=True)

        # device-resident sources take the prep streams only when the caller vouches that nothing is still writing them
        # (static_device_batches): the prep streams do not wait for the compute stream — that is their point
        dev_src = torch.is_tensor(cloud) and cloud.is_cuda
        pre = prefetch and after_clouds is not None and (not dev_src or self.static_device_batches)
        if pre:
            # pipelined caller: this slot's buffers are free (the step that used them has finished, _begin_step), the running
            # step reads the OTHER slot -> transfers, small fields and both geometries go to the two prep streams now
            p_state, p_next = self._prep_streams()
            ev = self._h2d_ev[slot]
            if self.has_critic:
                with torch.cuda.stream(p_next):
                    put_cloud(self.next_cloud, self._next_cloud_host, batch["next_point_state_batch"])
                    self._run(("gn",), lambda: self.geom_n.build(self.next_cloud, self.skip))
                    ev[1].record()
            else:
                ev[1].record(p_next)
            self._prep_pending = True
        with (torch.cuda.stream(self._prep_streams()[0]) if pre else contextlib.nullcontext()):
            self._prepare_state_side(batch, cloud, noise_u, after_clouds, put_cloud, slot, pre)
            if pre:
                self._h2d_ev[slot][0].record()

    def _prepare_state_side(self, batch, cloud, noise_u, after_clouds, put_cloud, slot, pre):
        """State cloud, (synchronous callers: next-state cloud,) geometry, the small per-sample fields."""
        B = cloud.shape[0]
        put_cloud(self.cloud, self._cloud_host, cloud)
        if pre:
            self._run(("gs",), lambda: self.geom_s.build(self.cloud, self.skip))
            after_clouds = None
        if self.has_critic and not pre:
            nxt = batch["next_point_state_batch"]
            # only the target chain (side stream) reads it: its H2D copy overlaps the state chain.  A device-resident
            # batch was produced on the current stream, so its D2D copy stays there
            if self.overlap and not (torch.is_tensor(nxt) and nxt.is_cuda):
                with torch.cuda.stream(self.side_enc.stream):
                    put_cloud(self.next_cloud, self._next_cloud_host, nxt)
            else:
                put_cloud(self.next_cloud, self._next_cloud_host, nxt)
        on_device = torch.is_tensor(cloud) and cloud.is_cuda
        early = after_clouds is not None and (on_device or self.early_geometry_host)
        if early:
            after_clouds()
        tgt = self.v if on_device else self.vh[slot]   # device-resident batch: D2D straight into the kernel inputs

        def put(dst, key_or_val):
            src = batch[key_or_val] if isinstance(key_or_val, str) else key_or_val
            if not torch.is_tensor(src):
                src = torch.as_tensor(np.asarray(src, dtype=np.float32))
            dst.copy_(src.view(dst.shape), non_blocking=True)

        put(tgt.action, "action_batch"), put(tgt.expert_action, "expert_action_batch"), put(tgt.goal, "goal_batch")
        put(tgt.reward, "reward_batch"), put(tgt.ret, "return_batch"), put(tgt.done, "mask_batch"), put(tgt.time, "time_batch")
        put(tgt.expert_flag, "expert_flag_batch"), put(tgt.perturb_flag, "perturb_flag_batch")
        if self.has_critic:
            # torch.rand_like in get_noise_delta (utils.py:575) unless the caller injects the uniform draw
            put(tgt.noise_u, torch.rand(B, 6, device=self.device if on_device else "cpu") if noise_u is None else noise_u)
        if not on_device:
            self.vec.copy_(self.vec_host[slot], non_blocking=True)
        if after_clouds is not None and not early:
            after_clouds()

    # ---- geometry (FPS, ball query, row tables) of the minibatch's clouds: needs xyz only ------------------------
    def _geometry(self, prep=False):
        """Launched right after the cloud transfers are issued.  State cloud on the main stream, next-state cloud on the
        target chain's stream (which already carries its H2D copy, or is forked off the main stream for device batches).
        ``prep``: both on the current (prep) stream — they run behind the previous step, off its critical path."""
        if self.has_critic:
            if self.overlap and not prep:
                side = self.side_enc
                side.fork()
                with torch.cuda.stream(side.stream):
                    self._run(("gn",), lambda: self.geom_n.build(self.next_cloud, self.skip))
            else:
                self._run(("gn",), lambda: self.geom_n.build(self.next_cloud, self.skip))
        self._run(("gs",), lambda: self.geom_s.build(self.cloud, self.skip))

    def h2d_bytes(self):
        B, C, Np = self._cloud_shape
        n = B * C * Np * (2 if self.has_critic else 1) + self.vec_host[0].numel()
        return 4 * n

    # ---- schedules ---------------------------------------------------------------------------------------
    def _mix_idx(self):
        return int((self.update_step > np.array(self.mix_milestones)).sum())

    def get_mix_ratio(self, update_step=None):
        idx = self._mix_idx()
        pick = lambda lst: lst[min(len(lst) - 1, idx)]  # noqa: E731
        return (min(pick(self.mix_value_ratio_list), self.ddpg_coefficients[3]),
                min(pick(self.mix_policy_ratio_list), self.ddpg_coefficients[4]))

    def _noise_scale(self):
        idx = self._mix_idx()
        return self.action_noise * self.noise_ratio_list[min(len(self.noise_ratio_list) - 1, idx)]

    def step_scheduler(self, step=None):
        """agent.py:179-190: critic, policy, whole-extractor and policy-encoder schedulers; the value-encoder
        scheduler is never stepped in the reference, so its lr stays at the initial value."""
        if self.has_critic:
            self._sch.critic.step()
        self._sch.policy.step()
        if self.train_feature or self.train_value_feature:
            self._sch.enc.step()

    def get_lr(self):
        return {"policy_lr": self._sch.policy.lr, "feature_lr": self._sch.enc.lr,
                "value_lr": self._sch.critic.lr if self.has_critic else 0}

    def _set_dyn(self, names):
        lrs = dict(policy=self._sch.policy.lr, critic=self._sch.critic.lr if self.has_critic else 0.0, enc=self._sch.enc.lr,
                   venc=self._sch.venc.lr)
        for i, n in enumerate(("policy", "critic", "enc", "venc")):
            if n in names:
                self.opt_steps[n] += 1
            t = max(self.opt_steps[n], 1)
            self.dyn_host[self._slot][i, 0] = lrs[n] / (1.0 - 0.9 ** t)
            self.dyn_host[self._slot][i, 1] = math.sqrt(1.0 - 0.999 ** t)
        self.dyn.copy_(self.dyn_host[self._slot], non_blocking=True)

    # ---- fused optimiser helpers ---------------------------------------------------------------------------
    def _adam(self, arena, off, n, which, clip=None, write_back=0):
        i = ("policy", "critic", "enc", "venc").index(which)
        eps, wd = self._adam_hp[which]
        gs = 1.0  # sharded runs average the gradients right after the all-reduce (see _allreduce)
        lib.gaddpg_adam_step(arena.p.data_ptr() + 4 * off, arena.g.data_ptr() + 4 * off, arena.m.data_ptr() + 4 * off,
                             arena.v.data_ptr() + 4 * off, n, 0.0, 0.9, 0.999, eps, wd, 0, self.dyn.data_ptr() + 8 * i, gs,
                             clip, write_back, None, 0.0, current_stream())

    def _opt_jobs(self, which):
        """Job tables of gaddpg_optim_multi, built once (every pointer in them is static)."""
        t = self._optjobs.get(which)
        if t is not None:
            return t
        t = engine.OptJobs(self.device)
        dyn = lambda name: self.dyn.data_ptr() + 8 * ("policy", "critic", "enc", "venc").index(name)  # noqa: E731
        out = lambda slot: self.out.data_ptr() + 4 * slot  # noqa: E731
        if which == "p2":
            A = self.ef_v.arena
            t.adam(A, 0, A.n, dyn("venc"), *self._adam_hp["venc"])
            A = self.cf.arena
            t.adam(A, 0, A.n, dyn("critic"), *self._adam_hp["critic"], clip=out(O_CLIP), write_back=1)
        else:   # "p3", "p3h" (DDPG) and "bc"
            A, T = self.pf.arena, self.pft.arena
            ranges, _ = self.pf.adam_ranges(self.policy_aux)
            pos = 0
            for off, n in ranges:       # Adam ranges start on 16-byte aligned arena segments; the gaps in between only see Polyak
                if off > pos:
                    t.polyak(A.p, T.p, pos, off - pos, tau=float(self.tau), absmax_p=out(O_PPARAM))
                t.adam(A, off, n, dyn("policy"), *self._adam_hp["policy"], target=T.p, tau=float(self.tau), absmax_p=out(O_PPARAM))
                pos = off + n
            if pos < A.n:
                t.polyak(A.p, T.p, pos, A.n - pos, tau=float(self.tau), absmax_p=out(O_PPARAM))
            if self.train_feature:
                E = self.ef_p.arena
                t.adam(E, 0, E.n, dyn("enc"), *self._adam_hp["enc"])
            if self.has_critic:
                C = self.cf.arena
                tv = (self.tau_soft + self.tau_hard) if which == "p3h" else self.tau_soft   # disjoint supports (Q1 | Q2)
                self._tau_keep = getattr(self, "_tau_keep", []) + [tv]
                t.polyak(C.p, self.cft.arena.p, 0, C.n, tau_vec=tv, grads=C.g, absmax_p=out(O_CPARAM), absmax_g=out(O_CGRAD))
        self._optjobs[which] = t
        return t

    # ---- gradient all-reduce (sharded runs): ONE contiguous range per optimiser phase (nets.GradPool) ----------------
    # DDP semantics: per-rank mean losses averaged over ranks.  The pool is laid out [encoder: SA1 | SA2 SA3 FC | heads];
    # with ``split_reduce`` everything behind the SA1 block — 99 % of the bytes, complete once the SA2 backward is done —
    # is reduced asynchronously on the collective's own stream while the SA1 backward (the longest part of a backward
    # pass) still runs, and only SA1's ~13 k gradients are reduced after it; otherwise one blocking call per phase.
    def _sharded(self):
        return self.world is not None and self.world.size > 1

    def _reduce_early(self, pool, ef):
        if self._sharded() and self.split_reduce:
            self._pending_reduce = [self.world.all_reduce_mean(pool.used[ef.sa1_end:], async_op=True)]

    def _reduce_late(self, pool, ef):
        if not self._sharded():
            return
        if self.split_reduce:
            self._pending_reduce.append(self.world.all_reduce_mean(pool.used[: ef.sa1_end], async_op=True))
            for h in self._pending_reduce:
                self.world.wait(h)      # stream-side wait: the host does not block
            self._pending_reduce = []
        else:
            self.world.all_reduce_mean(pool.used)

    def _begin_step(self):
        """Pick the host staging slot of this step; wait (host side) until the step that last used it has finished on the
        GPU, so its pinned buffers may be overwritten."""
        self._slot = (self._slot + 1) % RING
        old = self._pending[self._slot]
        if old is not None:
            old.result()                 # a deferred result nobody read yet: keep its scalars before the slot is reused
            self._pending[self._slot] = None
        self._slot_done[self._slot].synchronize()
        if self.step_start_events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.step_start_events.append(e)

    def _run(self, key, fn, outer=False):
        """Run ``fn`` eagerly the first time for a key, then capture and replay it as a CUDA graph.  ``outer``: ``fn`` is a
        whole sequence whose pieces call ``_run`` themselves — they run inline, so the sequence becomes ONE graph."""
        if not self.use_graph or self._in_outer:
            return fn()
        key = key + (self._slot,)   # the step's inputs (clouds, fields, geometry) are per staging slot: one capture per slot
        if outer:
            inner = fn

            def fn():
                self._in_outer = True
                try:
                    inner()
                finally:
                    self._in_outer = False
        g = self._graphs.get(key)
        if g is None:
            fn()  # eager warm-up doubles as this call's execution
            self._graphs[key] = "warm"
            return
        if g == "warm":
            # capture for the NEXT calls: state-changing kernels must not run twice, so capture happens on a side
            # stream without executing (torch.cuda.graph records, it does not run)
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            # thread_local: the NCCL watchdog thread of a sharded run keeps polling CUDA events while we capture
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                fn()
            self._graphs[key] = graph
            g = graph
        g.replay()

    # ---- statistics / result ----------------------------------------------------------------------------
    def _finish(self, defer=False):
        slot = self._slot
        self.out_host[slot].copy_(self.out, non_blocking=True)
        self._slot_done[slot].record()
        if self.step_end_events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.step_end_events.append(e)
        if defer:
            self._pending[slot] = PendingResult(self, slot, self._slot_done[slot])
            return self._pending[slot]
        self._slot_done[slot].synchronize()
        return self._scalars(self.out_host[slot])

    def _scalars(self, o):
        r = {k: 0.0 for k in LOSS_KEYS}
        r.update(bc_loss=float(o[O_BC]), policy_grasp_aux_loss=float(o[O_POLICY_AUX]), policy_param=float(o[O_PPARAM]))
        if self.has_critic:
            r.update(critic_grasp_aux_loss=float(o[O_CRITIC_AUX]), critic_loss=float(o[O_CRITIC]), actor_critic_loss=float(o[O_AC]),
                     reward_mask_num=float(o[O_NGOAL_C]), critic_grad=float(o[O_CGRAD]), critic_param=float(o[O_CPARAM]))
        else:
            r.update(reward_mask_num=float(o[O_NGOAL_A]))
        return r

    # ---- inference (agent.py:82-125) ----------------------------------------------------------------------
    @torch.no_grad()
    @_on_device
    def select_action(self, state, actions=None, goal_state=None, vis=False, remain_timestep=0, grasp_set=None,
                      gt_goal_rollout=False, repeat=False, eps=None):
        cloud = torch.as_tensor(np.asarray(state[0][0], dtype=np.float32))[None].to(self.device).contiguous()
        pi, logp, act, aux = self._act(cloud, float(remain_timestep), eps)
        return pi[0], logp[0], act[0], aux[0]

    @torch.no_grad()
    @_on_device
    def select_action_batch(self, point_state_batch, remain_timestep=0, eps=None):
        """Batched inference — what the reference's deployment script does by hand for a set of simulated view points
        (test_realworld_ros_final.py:1257-1261: ``extract_feature(img, point_state_batch, value=False, time_batch=time)``
        followed by ``policy.sample``): eval-mode policy path for a (B, C, N[+6]) batch of clouds and per-sample (or one
        scalar) remaining time steps.  Returns (tanh-mean action (B,6), log-prob (B,), sampled action (B,6), aux (B,E))."""
        cloud = point_state_batch if torch.is_tensor(point_state_batch) else torch.as_tensor(np.asarray(point_state_batch, dtype=np.float32))
        cloud = cloud.to(self.device, dtype=torch.float32).contiguous()
        assert cloud.dim() == 3, "point_state_batch must be (B, C, N)"
        return self._act(cloud, remain_timestep, eps)

    @torch.no_grad()
    @_on_device
    def extract_feature(self, image_batch, point_state_batch, action_batch=None, goal_batch=None, traj_point_state=None,
                        time_batch=None, vis=False, value=False, repeat=False, train=False):
        """ddpg.py:36-59 / bc.py:27-38 as an inference entry point (the update itself never calls this): eval-mode
        BatchNorm unless ``train``; ``value=True`` runs the value encoder on cloud (+) action.  Returns a NEW device tensor
        (B, 513) = [z | time] like the reference (use_time)."""
        cloud = point_state_batch if torch.is_tensor(point_state_batch) else torch.as_tensor(np.asarray(point_state_batch, dtype=np.float32))
        cloud = cloud.to(self.device, dtype=torch.float32).contiguous()
        B, C, Np = cloud.shape
        skip = 6 if Np != 1024 else 0
        st = self._act_state(B, C, Np)
        t = torch.zeros(B, device=self.device) if time_batch is None else torch.as_tensor(time_batch, dtype=torch.float32).to(self.device).view(B)
        st.geom.build(cloud, skip)
        if value:
            assert action_batch is not None, "value features need the action (channel-wise concat, utils.py:291-297)"
            Cp = min(C, self._extractor.critic_input_dim)
            a = torch.as_tensor(action_batch, dtype=torch.float32).to(self.device).view(B, -1)[:, : self._extractor.critic_input_dim - Cp].contiguous()
            feat = engine.encoder_forward(self.ws, self.ef_v, st.geom, cloud, skip, Cp, a if a.shape[1] else None, st.ctx,
                                          time=t, train=train, keep=False)
        else:
            feat = engine.encoder_forward(self.ws, self.ef_p, st.geom, cloud, skip, self.Cp_policy, None, st.ctx, time=t, train=train,
                                          keep=False)
        return feat[:, :513].clone()

    @torch.no_grad()
    @_on_device
    def policy_sample(self, feat, eps=None):
        """GaussianPolicy.sample (networks.py:353-371) on a (B, 513) device feature tensor through the CUDA heads:
        (tanh-mean action, log-prob, sampled action, aux) as device tensors.  Also bound as ``agent.policy.sample``."""
        B = feat.shape[0]
        f = torch.zeros(B, FEAT_LD, device=self.device)
        f[:, :513] = feat
        pc = engine.policy_ctx(B, self.pf, self.device)
        raw = engine.policy_forward(self.pf, f, pc, B)
        s = current_stream()
        act, logp, aux = torch.zeros(B, 6, device=self.device), torch.zeros(B, device=self.device), torch.zeros(B, 7, device=self.device)
        e = torch.randn(B, 6, device=self.device) if eps is None else torch.as_tensor(eps, dtype=torch.float32).to(self.device).view(B, 6)
        lib.gaddpg_policy_head_fwd(dp(raw), self.pf.NHp, B, dp(pc.pi), s)
        lib.gaddpg_policy_sample(dp(raw), self.pf.NHp, 6 + self.pf.E, dp(e.contiguous()), B, dp(act), dp(logp), s)
        if self.policy_aux:
            lib.gaddpg_quat_head(raw.data_ptr() + 4 * 6, self.pf.NHp, B, dp(aux), s)
        else:
            aux[:, : self.pf.E].copy_(raw[:, 6:6 + self.pf.E])
        return pc.pi, logp.view(B, 1), act, aux[:, : (7 if self.policy_aux else self.pf.E)]

    def _act_state(self, B, C, Np):
        """Static buffers of the inference path for one input shape (a captured graph is bound to them)."""
        dev = self.device
        skip = 6 if Np != 1024 else 0
        key = ("act", B, C, Np)
        if not hasattr(self, "_act_states"):
            self._act_states = {}
        st = self._act_states.get(key)
        if st is None:
            geom = engine.Geometry(B, Np - skip, dev)
            caps = (geom.lv[0].cap, geom.lv[1].cap)
            st = NS(key=key, geom=geom, ctx=engine.EncoderCtx(B, caps, engine.WIDTHS, dev), pc=engine.policy_ctx(B, self.pf, dev),
                    cloud=torch.zeros(B, C, Np, device=dev), time=torch.zeros(B, device=dev), eps=torch.zeros(B, 6, device=dev),
                    # [pi(6) | sampled action(6) | aux(7) | log-prob(1)] per sample, one D2H
                    res=torch.zeros(B, 20, device=dev), res_host=torch.zeros(B, 20).pin_memory(),
                    in_host=torch.zeros(B, 7).pin_memory(), in_dev=torch.zeros(B, 7, device=dev),
                    ctx_act=torch.zeros(B, 6, device=dev), ctx_logp=torch.zeros(B, device=dev), ctx_aux=torch.zeros(B, 7, device=dev))
            self._act_states[key] = st
        return st

    def _act(self, cloud, remain, eps):
        """Eval-mode policy path for a (B, C, N+6) cloud batch: geometry, policy encoder (running BatchNorm statistics),
        policy heads, tanh-mean / sampled action / log-prob / aux.  ~70 launches with fixed shapes and static buffers:
        the first call of a shape runs eagerly, the second captures a CUDA graph, later calls replay it (one graph launch
        + one 128-byte read-back per action; rollouts in train_test_offline.test are launch-latency bound otherwise)."""
        B, C, Np = cloud.shape
        skip = 6 if Np != 1024 else 0
        st = self._act_state(B, C, Np)
        key = st.key
        st.cloud.copy_(cloud, non_blocking=True)
        st.in_host[:, :6] = torch.randn(B, 6) if eps is None else torch.as_tensor(eps, dtype=torch.float32).view(B, 6)
        st.in_host[:, 6] = torch.as_tensor(remain, dtype=torch.float32)   # scalar or (B,)
        st.in_dev.copy_(st.in_host, non_blocking=True)

        def fn():
            s = current_stream()
            st.eps.copy_(st.in_dev[:, :6])
            st.time.copy_(st.in_dev[:, 6])
            st.geom.build(st.cloud, skip)
            feat = engine.encoder_forward(self.ws, self.ef_p, st.geom, st.cloud, skip, self.Cp_policy, None, st.ctx, time=st.time,
                                          train=False, keep=False)
            raw = engine.policy_forward(self.pf, feat, st.pc, B)
            pi, act, aux, logp = st.res[:, 0:6], st.res[:, 6:12], st.res[:, 12:19], st.res[:, 19]
            lib.gaddpg_policy_head_fwd(dp(raw), self.pf.NHp, B, dp(st.pc.pi), s)
            lib.gaddpg_policy_sample(dp(raw), self.pf.NHp, 6 + self.pf.E, dp(st.eps), B, dp(st.ctx_act), dp(st.ctx_logp), s)
            if self.policy_aux:
                lib.gaddpg_quat_head(raw.data_ptr() + 4 * 6, self.pf.NHp, B, dp(st.ctx_aux), s)
            else:
                st.ctx_aux.zero_()
                st.ctx_aux[:, : self.pf.E].copy_(raw[:, 6:6 + self.pf.E])
            pi.copy_(st.pc.pi), act.copy_(st.ctx_act), aux.copy_(st.ctx_aux), logp.copy_(st.ctx_logp)

        self._run(key, fn)
        st.res_host.copy_(st.res, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        r = st.res_host.numpy()
        E = 7 if self.policy_aux else self.pf.E
        return r[:, 0:6].copy(), r[:, 19].copy(), r[:, 6:12].copy(), r[:, 12:12 + E].copy()

    # ---- weights (ddpg.py:22-34, bc.py:15-25) -------------------------------------------------------------
    def state_dicts(self):
        d = {"policy": self.policy.state_dict(), "policy_target": self.policy_target.state_dict(),
             "state_feat": {"module." + k: v for k, v in self._extractor.state_dict().items()}}
        if self.has_critic:
            d["critic"], d["critic_target"] = self.critic.state_dict(), self.critic_target.state_dict()
        return d

    @_on_device
    def load_state_dicts(self, d):
        self.policy.load_state_dict(d["policy"])
        self.policy_target.load_state_dict(d["policy_target"])
        self._extractor.load_state_dict({k[len("module."):] if k.startswith("module.") else k: v for k, v in d["state_feat"].items()})
        if self.has_critic:
            self.critic.load_state_dict(d["critic"])
            self.critic_target.load_state_dict(d["critic_target"])
        self.refresh_all()

    @_on_device
    def refresh_all(self):
        for f in (self.ef_p, self.ef_v, self.pf, self.pft) + ((self.cf, self.cft) if self.has_critic else ()):
            f.refresh_derived()

    # ---- checkpoints in the reference's file / dict layout (agent.py:282-431) -------------------------------
    def _opt_table(self):
        """(file, key of the optimiser dict, key of the scheduler dict, parameters in torch order, moments callback,
        Adam-step counter name, scheduler, eps, weight decay) for every optimiser the reference checkpoints."""
        ex = self._extractor
        pr, _ = self.pf.adam_ranges(self.policy_aux)
        rows = [("actor", "opt", "sch", list(self.policy.parameters()), checkpoint.arena_moments(self.pf.arena, pr), "policy",
                 self._sch.policy, 1e-5, 1e-5)]
        if self.has_critic:
            rows.append(("critic", "opt", "sch", list(self.critic.parameters()), checkpoint.arena_moments(self.cf.arena), "critic",
                         self._sch.critic, 1e-5, 1e-5))
        rows += [
            # the whole-extractor optimiser is never stepped (only its scheduler is, agent.py:187-189)
            ("state_feat", "opt", "sch", list(ex.parameters()), lambda i, p: None, None, self._sch.enc, 1e-8, 0.0),
            ("state_feat", "encoder_opt", "encoder_sch", list(ex.encoder.parameters()), checkpoint.arena_moments(self.ef_p.arena),
             "enc", self._sch.enc) + self._adam_hp["enc"],
            ("state_feat", "val_encoder_opt", "val_encoder_sch", list(ex.value_encoder.parameters()),
             checkpoint.arena_moments(self.ef_v.arena), "venc", self._sch.venc) + self._adam_hp["venc"],
        ]
        return rows

    @_on_device
    def save_model(self, step, output_dir="", surfix="latest", actor_path=None, critic_path=None, goal_feat_path=None,
                   state_feat_path=None):
        """agent.py:282-352: same three files, same dict keys, torch-format ``Adam`` / ``MultiStepLR`` state dicts
        assembled from the fused arenas, so the reference's ``load_model`` reads them unchanged."""
        torch.cuda.synchronize(self.device)
        p = checkpoint.paths(output_dir, self.name, self.env_name, surfix)
        files = {"actor": actor_path or p["actor"], "critic": critic_path or p["critic"],
                 "state_feat": state_feat_path or p["state_feat"]}
        cpu = lambda sd: {k: v.detach().cpu().clone() for k, v in sd.items()}  # noqa: E731
        out = {"actor": {"net": cpu(self.policy.state_dict())},
               "state_feat": {"net": cpu(self.state_dicts()["state_feat"]), "step": step}}  # "module."-prefixed (DataParallel)
        if self.has_critic:
            out["critic"] = {"net": cpu(self.critic.state_dict())}
        for f, ok, sk, params, moments, which, sch, eps, wd in self._opt_table():
            cpu_m = (lambda mo: (lambda i, q: None if mo(i, q) is None else tuple(t.cpu() for t in mo(i, q))))(moments)
            out[f][ok] = checkpoint.adam_state_dict(params, cpu_m, self.opt_steps[which] if which else 0, sch.lr, eps, wd,
                                                    initial_lr=sch.sched.base_lrs[0])
            out[f][sk] = sch.sched.state_dict()
        for f, d in out.items():
            checkpoint.save(d, files[f])
        return files

    @_on_device
    def load_model(self, output_dir, surfix="latest", set_init_step=False, reinit_value_feat=False):
        """agent.py:354-431.  Returns the restored ``update_step`` (0 when there is no feature-extractor file)."""
        import os
        p = checkpoint.paths(output_dir, self.name, self.env_name, surfix)
        table = {}
        for row in self._opt_table():
            table.setdefault(row[0], []).append(row)
        strip = lambda sd: {k[len("module."):] if k.startswith("module.") else k: v for k, v in sd.items()}  # noqa: E731

        def restore(d, rows, tolerant):
            for _, ok, sk, params, moments, which, sch, eps, wd in rows:
                try:
                    step, lr = checkpoint.load_adam_state_dict(d[ok], params, moments)
                    if sk in d:
                        sch.sched.load_state_dict(d[sk])
                    sch.opt.param_groups[0]["lr"] = lr
                    if which:
                        self.opt_steps[which] = step
                except (KeyError, ValueError, RuntimeError):
                    if not tolerant:
                        raise
                    print("loading feature optim has mismatches")  # agent.py:418-419

        def reinit(sch, milestones):
            # agent.py:380-389 / 401-409: restart at reinit_lr with a fresh gamma=0.5 schedule
            fresh = _Sched(self.reinit_lr, milestones, 0.5)
            sch.opt, sch.sched = fresh.opt, fresh.sched

        if os.path.exists(p["actor"]):
            d = checkpoint.load(p["actor"])
            self.policy.load_state_dict(d["net"])
            restore(d, table["actor"], tolerant=False)
            if getattr(self, "reinit_optim", False) and set_init_step:
                reinit(self._sch.policy, self.policy_milestones)
            self.policy_target.load_state_dict(self.policy.state_dict())       # hard_update (utils.py:761-763)
        if self.has_critic and os.path.exists(p["critic"]):
            d = checkpoint.load(p["critic"])
            self.critic.load_state_dict(d["net"])
            restore(d, table["critic"], tolerant=False)
            if getattr(self, "reinit_optim", False) and set_init_step:
                reinit(self._sch.critic, self.value_milestones)
            self.critic_target.load_state_dict(self.critic.state_dict())
        step = 0
        if os.path.exists(p["state_feat"]):
            d = checkpoint.load(p["state_feat"])
            self._extractor.load_state_dict(strip(d["net"]))
            restore(d, table["state_feat"], tolerant=True)
            self.update_step = step = d["step"]
            if set_init_step:
                self.init_step = self.update_step
        self.refresh_all()
        self._graphs = {}   # schedule index / learning-rate state may have moved: re-capture lazily
        torch.cuda.synchronize(self.device)
        return step


class DDPGB200(AgentB200):
    name = "DDPG"

    def get_weight(self):
        return [self.policy.state_dict(), self.critic.state_dict(), {}, self.state_feature_extractor.state_dict()]

    def load_weight(self, weights):
        self.policy.load_state_dict(weights[0])
        self.critic.load_state_dict(weights[1])
        self.state_feature_extractor.load_state_dict(weights[3])
        self.refresh_all()

    # -- phase 1: critic side ------------------------------------------------------------------------------
    # Two independent encoder chains run side by side on two streams:
    #   state chain  (main): geometry(state), F1 value(state, a), F4 policy(state), critic(F1)
    #   target chain (side): geometry(next),  F2 policy(next) -> a', F3 value(next, a') -> critic_target -> y
    # F4 belongs to the actor branch (ddpg.py:164-166) but depends only on the state cloud and the policy-encoder
    # weights, which nothing touches before phase 3, so it is issued here where it overlaps the target chain.  Each
    # pass stages its BatchNorm batch statistics; _phase1_critic then updates the running statistics in the
    # reference's order (value encoder: F1, F3; policy encoder: F2, F4).  The target chain only needs the next-state
    # cloud, whose H2D copy is issued on the side stream, so it also overlaps the state chain's first kernels.
    def _phase1_state(self):
        B, ws, v = self.B, self.ws, self.v
        self.out.zero_()
        if self.overlap and self.f4_par and not self.f4_late:
            # F4 as a third concurrent chain (the weight-gradient stream and its scratch are idle during phase 1)
            self.side_dw.run(lambda w: self._f4(w))
        f1 = engine.encoder_forward(ws, self.ef_v, self.geom_s, self.cloud, self.skip, self.Cp_value, self._bc(v.action, 0), self.ctx_v1,
                                    time=v.time, time_offset=0.0, train=True, bn_stage=self.bnst[1])              # F1
        if self.overlap and self.f4_par and not self.f4_late:
            self.side_dw.join()
        elif not (self.overlap and self.f4_late):
            self._f4(ws)
        engine.critic_forward(self.cf, f1, self.cc1, B)

    def _f4(self, ws):
        v = self.v
        engine.encoder_forward(ws, self.ef_p, self.geom_s, self.cloud, self.skip, self.Cp_policy, None, self.ctx_p,
                               time=v.time, time_offset=0.0, train=True, bn_stage=self.bnst[4])                   # F4

    def _phase1_target(self, ws):
        B, v, s = self.B, self.v, current_stream()
        f2 = engine.encoder_forward(ws, self.ef_p, self.geom_n, self.next_cloud, self.skip, self.Cp_policy, None, self.ctx_n,
                                    time=v.time, time_offset=-1.0, train=True, bn_stage=self.bnst[2], keep=False)  # F2
        rawt = engine.policy_forward(self.pft, f2, self.pct, B)
        lib.gaddpg_td3_next_action(dp(rawt), self.pft.NHp, dp(v.noise_u), float(self._noise_scale()), B, dp(self.next_action), s)
        f3 = engine.encoder_forward(ws, self.ef_v, self.geom_n, self.next_cloud, self.skip, self.Cp_value,
                                    self._bc(self.next_action, 1), self.ctx_n, time=v.time, time_offset=-1.0, train=True,
                                    bn_stage=self.bnst[3], keep=False)                                             # F3
        qat = engine.critic_forward(self.cft, f3, self.cct, B, nb=2)
        lib.gaddpg_td3_target(dp(qat), QA_LD, QA_Q2, dp(v.reward), dp(v.done), float(self.gamma), B, dp(self.y), s)

    def _phase1_critic(self, part="all"):
        """``part``: "all", or "a" (losses, critic backward, value-encoder backward down to SA2) / "b" (SA1 backward) when
        the sharded run reduces the upper gradients in between (_reduce_early)."""
        B, ws, v, s = self.B, self.ws, self.v, current_stream()
        qa, f1 = self.cc1.qa, self.ctx_v1.feat
        dw = self.side_dw if self.overlap else None
        late = self.overlap and self.f4_late
        if part != "b":
            if late:
                # F4 (policy encoder on the state cloud; consumed in phase 2) beside the critic backward: the upper half of
                # B1 is a chain of small launches that leaves most SMs idle, while phase 1 already has two chains in flight
                side = self.side_enc
                side.fork()
                with torch.cuda.stream(side.stream):
                    self._f4(side.ws)
            for k in ((1, 3) if late else (1, 3, 2, 4)):
                self.bnst[k].apply()
            lib.gaddpg_critic_loss(dp(qa), QA_LD, QA_Q2, QA_AUX, dp(self.y), dp(v.perturb_flag), dp(v.ret), dp(v.goal),
                                   1 if self.critic_aux else 0, B, 1.0, dp(self.cc1.dqa), self.out.data_ptr() + 4 * O_CRITIC, s)
        if part == "all":
            with engine.side_dw(dw):
                et = engine.entry_tail(ws, self.ef_v, self.ctx_v1, self.sc, want_dw=True, accumulate=0)
                engine.critic_backward(ws, self.cf, f1, self.cc1, B, self.cf.nb, self.ctx_v1, self.sc, accumulate=0, enc_tail=et)  # B1
                engine.encoder_backward(ws, self.ef_v, self.ctx_v1, self.sc, want_dw=True, want_dbc=False, accumulate=0,
                                        entry_done=et is not None)
        elif part == "a":
            with engine.side_dw(dw):
                et = engine.entry_tail(ws, self.ef_v, self.ctx_v1, self.sc, want_dw=True, accumulate=0)
                engine.critic_backward(ws, self.cf, f1, self.cc1, B, self.cf.nb, self.ctx_v1, self.sc, accumulate=0, enc_tail=et)
                engine.encoder_backward(ws, self.ef_v, self.ctx_v1, self.sc, want_dw=True, want_dbc=False, accumulate=0, part="upper",
                                        entry_done=et is not None)
        else:
            with engine.side_dw(dw):
                engine.encoder_backward(ws, self.ef_v, self.ctx_v1, self.sc, want_dw=True, want_dbc=False, accumulate=0, part="sa1")
        if late and part != "b":
            self.side_enc.join()
            for k in (2, 4):   # policy encoder's running statistics in the reference's order (F2, F4)
                self.bnst[k].apply()

    def _phase1(self, sig):
        if self.overlap:
            side = self.side_enc
            side.fork()  # the side stream sees the vector / state-cloud copies issued on the main stream
            with torch.cuda.stream(side.stream):
                self._run(("p1t",) + sig, lambda: self._phase1_target(side.ws))
            self._run(("p1s",) + sig, self._phase1_state)
            side.join()
        else:
            self._run(("p1s",) + sig, self._phase1_state)
            self._run(("p1t",) + sig, lambda: self._phase1_target(self.ws))
        if self._sharded() and self.split_reduce:
            self._run(("p1ca",) + sig, lambda: self._phase1_critic("a"))
            self._reduce_early(self.gpool_c, self.ef_v)
            self._run(("p1cb",) + sig, lambda: self._phase1_critic("b"))
        else:
            self._run(("p1c",) + sig, self._phase1_critic)

    # -- phase 2: critic/value-encoder step, then the actor side -----------------------------------------------
    def _phase2(self, even, part="all"):
        B, ws, v, s = self.B, self.ws, self.v, current_stream()
        dw = self.side_dw if self.overlap else None
        if part == "b":
            with engine.side_dw(dw):
                engine.encoder_backward(ws, self.ef_p, self.ctx_p, self.sc, want_dw=True, want_dbc=False, accumulate=0, part="sa1")
            return
        cA = self.cf.arena
        lib.gaddpg_clip_coef(dp(cA.g), cA.n, float(self.clip_grad), self.out.data_ptr() + 4 * O_CLIP,
                             self.out.data_ptr() + 4 * O_GNORM, dp(ws.red), s)
        self._opt_jobs("p2").launch()               # Adam(value encoder) + Adam(critic, clipped, gradient written back): one launch
        engine.refresh_many((self.ef_v, self.cf))   # derived weight layouts of both: one launch
        f4 = self.ctx_p.feat                                                                                     # F4: phase 1
        raw = engine.policy_forward(self.pf, f4, self.pc, B)
        lib.gaddpg_policy_head_fwd(dp(raw), self.pf.NHp, B, dp(self.pc.pi), s)
        mix = self.get_mix_ratio()[1]
        dpi = None
        if even:
            f5 = engine.encoder_forward(ws, self.ef_v, self.geom_s, self.cloud, self.skip, self.Cp_value, self._bc(self.pc.pi, 2),
                                        self.ctx_v5, time=v.time, time_offset=0.0, train=True)                   # F5
            qa5 = engine.critic_forward(self.cf, f5, self.cc5, B, nb=2)
            lib.gaddpg_actor_critic_loss(dp(qa5), QA_LD, QA_Q2, dp(v.ret), dp(v.expert_flag), float(mix), B, 1.0, QA_LD,
                                         dp(self.cc5.dqa), self.out.data_ptr() + 4 * O_AC, s)
            with engine.side_dw(dw):
                et = engine.entry_tail(ws, self.ef_v, self.ctx_v5, self.sc, want_dw=False, accumulate=0)
                engine.critic_backward(ws, self.cf, f5, self.cc5, B, 2, self.ctx_v5, self.sc, accumulate=1, enc_tail=et)
                dpi = engine.encoder_backward(ws, self.ef_v, self.ctx_v5, self.sc, want_dw=False, want_dbc=True, entry_done=et is not None)
            self.dpi_ac.zero_()
            self.dpi_ac[:, : self.Cb_value].copy_(dpi)
        ranges, n_grad = self.pf.adam_ranges(self.policy_aux)
        lib.gaddpg_actor_loss(dp(raw), self.pf.NHp, dp(self.pc.pi), dp(v.expert_action), dp(v.expert_flag), dp(v.ret), dp(v.goal),
                              1 if self.policy_aux else 0, float(1.0 - mix), dp(self.dpi_ac) if even else None, B, 1.0,
                              dp(self.pc.draw), self.pf.NHp, self.out.data_ptr() + 4 * O_BC, s)
        with engine.side_dw(dw):
            et = engine.entry_tail(ws, self.ef_p, self.ctx_p, self.sc, want_dw=True, accumulate=0)
            engine.policy_backward(ws, self.pf, f4, self.pc, B, n_grad, self.ctx_p, self.sc, accumulate=0, enc_tail=et)   # B2
            engine.encoder_backward(ws, self.ef_p, self.ctx_p, self.sc, want_dw=True, want_dbc=False, accumulate=0,
                                    part="upper" if part == "a" else "all", entry_done=et is not None)

    # -- phase 3: actor step, targets, statistics --------------------------------------------------------------
    def _phase3(self, hard):
        """Adam(policy) [+ Polyak of the policy target + max |param|], Adam(policy encoder), half-soft / half-hard update of the
        critic target + max |critic param| + max |critic grad|: ONE launch; the derived weight layouts: one more."""
        self._opt_jobs("p3h" if hard else "p3").launch()
        engine.refresh_many((self.pf, self.ef_p, self.pft, self.cft))

    @_on_device
    def update_parameters(self, batch_data, updates=None, k=None, test=False, noise_u=None, staged=False, defer=False):
        """ddpg.py:146-185.  ``staged=True``: the batch is already in the device buffers (bench/value path).
        ``defer=True``: return a ``PendingResult`` instead of waiting for the step's scalars."""
        self._begin_step()
        if not staged:
            self.prepare_data(batch_data, noise_u, after_clouds=self._geometry, prefetch=defer)
        else:
            self._geometry()
        self._wait_prep()
        even = (self.update_step % self.policy_update_gap) == 0
        hard = (self.update_step % self.target_update_interval) == 0
        sig = (self._mix_idx(),)
        self._set_dyn(("critic", "venc", "policy") + (("enc",) if self.train_feature else ()))
        if self.whole_graph and (not self._sharded() or (self.world.backend == "nccl" and not self.split_reduce)):
            self._run(("step", even, hard) + sig, lambda: self._step_body(even, hard, sig), outer=True)
        else:
            self._step_body(even, hard, sig)
        self.update_step += 1
        return self._finish(defer)

    def _step_body(self, even, hard, sig):
        self._phase1(sig)
        self._reduce_late(self.gpool_c, self.ef_v)
        if self._sharded() and self.split_reduce:
            self._run(("p2a", even) + sig, lambda: self._phase2(even, "a"))
            self._reduce_early(self.gpool_a, self.ef_p)
            self._run(("p2b", even) + sig, lambda: self._phase2(even, "b"))
        else:
            self._run(("p2", even) + sig, lambda: self._phase2(even))
        self._reduce_late(self.gpool_a, self.ef_p)
        self._run(("p3", hard) + sig, lambda: self._phase3(hard))


class BCB200(AgentB200):
    name = "BC"

    def get_weight(self):
        return [self.policy.state_dict(), {}, self.state_feature_extractor.state_dict()]

    def load_weight(self, weights):
        self.policy.load_state_dict(weights[0])
        self.state_feature_extractor.load_state_dict(weights[2])
        self.refresh_all()

    def _phase(self):
        B, ws, v, s = self.B, self.ws, self.v, current_stream()
        self.out.zero_()
        f = engine.encoder_forward(ws, self.ef_p, self.geom_s, self.cloud, self.skip, self.Cp_policy, None, self.ctx_p,
                                   time=v.time, time_offset=0.0, train=True)
        raw = engine.policy_forward(self.pf, f, self.pc, B)
        lib.gaddpg_policy_head_fwd(dp(raw), self.pf.NHp, B, dp(self.pc.pi), s)
        ranges, n_grad = self.pf.adam_ranges(self.policy_aux)
        lib.gaddpg_actor_loss(dp(raw), self.pf.NHp, dp(self.pc.pi), dp(v.expert_action), dp(v.expert_flag), dp(v.ret), dp(v.goal),
                              1 if self.policy_aux else 0, 1.0, None, B, 1.0, dp(self.pc.draw), self.pf.NHp,
                              self.out.data_ptr() + 4 * O_BC, s)
        with engine.side_dw(self.side_dw if self.overlap else None):
            et = engine.entry_tail(ws, self.ef_p, self.ctx_p, self.sc, want_dw=True, accumulate=0)
            engine.policy_backward(ws, self.pf, f, self.pc, B, n_grad, self.ctx_p, self.sc, accumulate=0, enc_tail=et)
            engine.encoder_backward(ws, self.ef_p, self.ctx_p, self.sc, want_dw=True, want_dbc=False, accumulate=0,
                                    entry_done=et is not None)

    def _phase_opt(self):
        self._opt_jobs("bc").launch()
        engine.refresh_many((self.pf, self.ef_p, self.pft))

    @_on_device
    def update_parameters(self, batch_data, updates=None, k=None, noise_u=None, staged=False, defer=False):
        """bc.py:40-56."""
        self._begin_step()
        if not staged:
            self.prepare_data(batch_data, after_clouds=self._geometry, prefetch=defer)
        else:
            self._geometry()
        self._wait_prep()
        self._set_dyn(("policy",) + (("enc",) if self.train_feature else ()))
        self._run(("bc",), self._phase)
        self._reduce_late(self.gpool_a, self.ef_p)   # BC is one short backward: a single blocking call
        self._run(("bcopt",), self._phase_opt)
        self.update_step += 1
        return self._finish(defer)


def make_agent(policy="DDPG", seed=123456, device=None, world=None, **overrides):
    """Stand-alone construction (no reference code needed): same seed -> same initial weights as the reference
    driver (core/train_test_offline.py:305-349), checked against the oracle in tests."""
    cls = DDPGB200 if policy == "DDPG" else BCB200
    return cls(512, None, dict(DEFAULTS, **overrides), device=device, seed=seed, world=world).build_standalone()
