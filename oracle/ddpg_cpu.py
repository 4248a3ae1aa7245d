"""CPU restatement of the offline actor-critic update (oracle; TEST INFRASTRUCTURE ONLY).

One class, ``OracleAgent``, restates what the reference spreads over
/root/reference/core/agent.py (Agent: ctor :21-48, unpack_batch :51-80, select_action :82-125,
compute_loss :127-139, step_scheduler :179-190, optimize :192-209, prepare_data :211-240, log_stat
:242-259, set_mode :261-280), core/ddpg.py (extract_feature :36-59, target_value :61-88,
state_action_value :91-106, get_mix_ratio :108-117, compute_critic_loss :119-130, critic_optimize
:132-143, update_parameters :146-185) and core/bc.py :40-56, with the hyper-parameters of
experiments/config.py:67-131 and the optimiser factories utils.py:183-237,960-1006.

It is pinned: oracle/make_golden.py runs the UNMODIFIED reference classes on the same seeds/batches
(through oracle/refstack.py) and requires equality of every returned scalar and parameter before it
writes tests/golden/.  The set-abstraction ops underneath are the un-pinned part (oracle/__init__.py).

Noise: the reference draws, per DDPG step and in this order, randn(B,6) [policy_target.sample, unused],
rand(B,6) [TD3 target noise, utils.py:575], randn(B,6) [policy.sample, unused].  ``update_parameters``
draws the same three from torch's global generator unless ``noise_u`` (the rand(B,6) in [0,1)) is given,
which is how the GPU parity tests feed identical noise to both sides.
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch.optim import Adam
from torch.optim.lr_scheduler import MultiStepLR

from . import losses_cpu as L
from . import nets_cpu as nets

# experiments/config.py:67-131 (RL_TRAIN subset read by the update path) + model-spec YAML :18-24
DEFAULTS = dict(
    clip_grad=0.5, gamma=0.95, hidden_size=256, tau=0.0001, lr=3e-4, value_lr=3e-4, lr_gamma=0.5,
    value_lr_gamma=0.5, feature_input_dim=512, ddpg_coefficients=[0.0, 0.0, 1.0, 1.0, 0.2],
    value_milestones=[20000, 40000, 60000, 80000], policy_milestones=[20000, 40000, 60000, 80000],
    mix_milestones=[4000, 8000, 20000, 40000, 60000, 80000, 100000, 140000, 180000],
    mix_policy_ratio_list=[0.1, 0.2], mix_value_ratio_list=[1.0], policy_extra_latent=-1,
    critic_extra_latent=-1, train_feature=True, train_value_feature=True, sa_channel_concat=True,
    use_time=True, policy_update_gap=2, policy_aux=True, critic_aux=True, action_noise=0.01,
    noise_ratio_list=[3.0, 2.5, 2.0, 1.5, 1, 0.5], noise_type="uniform", target_update_interval=3000,
    extra_latent=1, feat_lr=1e-3, feat_milestones=[8000, 16000, 30000, 50000, 70000, 90000], feat_gamma=0.3,
)
LOSS_KEYS = ["bc_loss", "policy_grasp_aux_loss", "critic_grasp_aux_loss", "critic_loss", "actor_critic_loss",
             "reward_mask_num", "expert_mask_num", "policy_param", "critic_grad", "critic_param",
             "train_batch_size"]  # utils.py:1008-1020


def _pick(lst, i):
    return lst[min(len(lst) - 1, i)]


def _absmax_param(mod):
    return float(max(float(p.data.abs().max()) for p in mod.parameters()))


def _absmax_grad(mod):
    return float(max((float(p.grad.abs().max()) if p.grad is not None else 0.0) for p in mod.parameters()))


class OracleAgent:
    def __init__(self, policy="DDPG", seed=None, **overrides):
        self.cfg = dict(DEFAULTS, **overrides)
        c = self.cfg
        self.name = policy
        self.has_critic = policy != "BC"
        if seed is not None:
            torch.manual_seed(seed)
        nets.burn_goal_feature_rng()
        self.feat = nets.PointFeature(c["extra_latent"], c["policy_extra_latent"], c["critic_extra_latent"],
                                      action_concat=c["sa_channel_concat"])
        fo = dict(lr=c["feat_lr"])
        fs = dict(milestones=c["feat_milestones"], gamma=c["feat_gamma"])
        self.feat_opt = Adam(self.feat.parameters(), **fo)          # stepped by nobody; scheduler only
        self.feat_sched = MultiStepLR(self.feat_opt, **fs)
        self.enc_opt = Adam(self.feat.encoder.parameters(), **fo)
        self.enc_sched = MultiStepLR(self.enc_opt, **fs)
        self.venc_opt = Adam(self.feat.value_encoder.parameters(), **fo)
        self.venc_sched = MultiStepLR(self.venc_opt, **fs)          # never stepped (agent.py:179-190)
        num_inputs = c["feature_input_dim"] + (1 if c["use_time"] else 0)
        pdim = 7 if c["policy_aux"] else 1
        self.policy = nets.Policy(num_inputs, 6, c["hidden_size"], pdim)
        self.policy_opt = Adam(self.policy.parameters(), lr=c["lr"], eps=1e-5, weight_decay=1e-5)
        self.policy_sched = MultiStepLR(self.policy_opt, milestones=c["policy_milestones"], gamma=c["lr_gamma"])
        self.policy_target = nets.Policy(num_inputs, 6, c["hidden_size"], pdim)
        if self.has_critic:
            cdim = 7 if c["critic_aux"] else 0
            self.critic = nets.TwinQ(num_inputs, c["hidden_size"], cdim)
            self.critic_opt = Adam(self.critic.parameters(), lr=c["value_lr"], eps=1e-5, weight_decay=1e-5)
            self.critic_sched = MultiStepLR(self.critic_opt, milestones=c["value_milestones"],
                                            gamma=c["value_lr_gamma"])
            self.critic_target = nets.TwinQ(num_inputs, c["hidden_size"], cdim)
        self.update_step = 1

    # ---- weights in the reference's checkpoint layout -------------------------------------------
    def state_dicts(self):
        d = {"policy": self.policy.state_dict(), "policy_target": self.policy_target.state_dict(),
             "state_feat": {"module." + k: v for k, v in self.feat.state_dict().items()}}
        if self.has_critic:
            d["critic"] = self.critic.state_dict()
            d["critic_target"] = self.critic_target.state_dict()
        return d

    def load_state_dicts(self, d):
        self.policy.load_state_dict(d["policy"])
        self.policy_target.load_state_dict(d["policy_target"])
        self.feat.load_state_dict({k[len("module."):] if k.startswith("module.") else k: v
                                   for k, v in d["state_feat"].items()})
        if self.has_critic:
            self.critic.load_state_dict(d["critic"])
            self.critic_target.load_state_dict(d["critic_target"])

    # ---- feature extraction (ddpg.py:36-59 / bc.py:27-38) -----------------------------------------
    def features(self, cloud, time, action=None, value=False):
        pc = cloud
        if self.has_critic and self.cfg["sa_channel_concat"] and value:
            pc = torch.cat((pc, action.unsqueeze(2).expand(-1, -1, pc.shape[2])), 1)  # utils.py:291-297
        z = self.feat(pc, value=value)
        if self.cfg["use_time"]:
            z = torch.cat((z, time[:, None]), dim=1)
        return z

    def _mix_policy_ratio(self):
        idx = int((self.update_step > np.array(self.cfg["mix_milestones"])).sum())
        return min(_pick(self.cfg["mix_policy_ratio_list"], idx), self.cfg["ddpg_coefficients"][4])

    def _load(self, batch):
        t = {k: torch.FloatTensor(np.asarray(v)) for k, v in batch.items()}
        d = dict(cloud=t["point_state_batch"], next_cloud=t["next_point_state_batch"], action=t["action_batch"],
                 expert_action=t["expert_action_batch"], reward=t["reward_batch"], ret=t["return_batch"],
                 done=t["mask_batch"], time=t["time_batch"], goal=t["goal_batch"])
        d["reward_mask"] = (t["return_batch"] > 0).view(-1)
        d["expert_mask"] = (t["expert_flag_batch"] >= 1).view(-1)
        d["expert_reward_mask"] = d["reward_mask"] * d["expert_mask"]
        d["train_rows"] = (t["perturb_flag_batch"] < 1).bool()  # agent.py:229
        return d

    # ---- the DDPG step (ddpg.py:146-185) -----------------------------------------------------------
    def update_parameters(self, batch, noise_u=None):
        if not self.has_critic:
            return self._bc_step(batch)
        c = self.cfg
        out = {k: 0.0 for k in LOSS_KEYS}
        mix = self._mix_policy_ratio()
        self.feat.train(), self.policy.train(), self.critic.train()
        self.critic_opt.zero_grad()
        self.venc_opt.zero_grad()
        d = self._load(batch)

        value_feat = self.features(d["cloud"], d["time"], d["action"], value=True)           # F1
        with torch.no_grad():                                                                # target_value
            nt = d["time"] - 1
            nf = self.features(d["next_cloud"], nt, value=False)                             # F2 (train-mode BN)
            na, _, _, _ = self.policy_target.sample(nf)
            idx = int((self.update_step > np.array(c["mix_milestones"])).sum())
            scale = c["action_noise"] * _pick(c["noise_ratio_list"], idx)
            u = torch.rand_like(na) if noise_u is None else torch.as_tensor(noise_u, dtype=torch.float32)
            if c["noise_type"] == "uniform":
                delta = (u * 3 - 6) * scale                                                  # utils.py:575 (sic)
            else:
                raise NotImplementedError("oracle covers the default uniform TD3 noise only")
            delta[:, 3:] *= 5
            delta[:, :3] = torch.clamp(delta[:, :3], -0.01, 0.01)
            na = na + delta
            ntf = self.features(d["next_cloud"], nt, na, value=True)                         # F3
            q1t, q2t, _ = self.critic_target(ntf)
            y = d["reward"] + (1 - d["done"]) * c["gamma"] * torch.min(q1t, q2t).squeeze()
        q1, q2, caux = self.critic(value_feat)
        rows = d["train_rows"]
        critic_loss = F.smooth_l1_loss(q1.view(-1)[rows], y[rows]) + F.smooth_l1_loss(q2.view(-1)[rows], y[rows])
        critic_aux_loss = torch.zeros(1)
        if c["critic_aux"]:
            gm = d["reward_mask"]
            critic_aux_loss = critic_aux_loss + L.goal_pred_loss(caux[gm, :7], d["goal"][gm])
        self.critic_opt.zero_grad()
        self.venc_opt.zero_grad()
        (critic_aux_loss + critic_loss).backward()                                           # B1
        torch.nn.utils.clip_grad_norm_(self.critic.parameters(), c["clip_grad"])
        self.venc_opt.step()
        self.critic_opt.step()

        policy_feat = self.features(d["cloud"], d["time"], value=False)                      # F4
        pi, _, _, aux = self.policy.sample(policy_feat)
        actor_critic_loss = torch.zeros(1)
        if self.update_step % c["policy_update_gap"] == 0:
            vpf = self.features(d["cloud"], d["time"], pi, value=True)                       # F5
            q1p, q2p, _ = self.critic(vpf)
            sel = ~d["expert_reward_mask"]
            actor_critic_loss = -mix * torch.min(q1p.squeeze()[sel], q2p.squeeze()[sel]).mean()
        policy_aux_loss = torch.zeros(1)
        if c["policy_aux"]:
            gm = d["reward_mask"]
            policy_aux_loss = L.goal_pred_loss(aux[gm, :7], d["goal"][gm, :7])
        em = d["expert_mask"]
        bc_loss = L.pose_bc_loss(pi[em], d["expert_action"][em]) * (1 - mix)
        loss = bc_loss + policy_aux_loss + actor_critic_loss
        self.enc_opt.zero_grad()
        self.policy_opt.zero_grad()
        loss.backward()                                                                      # B2
        self.policy_opt.step()
        if c["train_feature"]:
            self.enc_opt.step()
        self._target_updates()
        self.update_step += 1
        out.update(bc_loss=float(bc_loss), policy_grasp_aux_loss=float(policy_aux_loss),
                   critic_grasp_aux_loss=float(critic_aux_loss), critic_loss=float(critic_loss),
                   actor_critic_loss=float(actor_critic_loss), reward_mask_num=float(d["reward_mask"].sum()),
                   policy_param=_absmax_param(self.policy), critic_grad=_absmax_grad(self.critic),
                   critic_param=_absmax_param(self.critic))
        self.last = dict(y=y.detach(), q1=q1.detach(), q2=q2.detach(), pi=pi.detach(), value_feat=value_feat.detach(),
                         policy_feat=policy_feat.detach())
        return out

    def _target_updates(self):
        """agent.py:204-209 + utils.py:750-770: Polyak on the whole policy target; on the critic target only
        linear1-3 (Q1) are Polyak-averaged, linear4-6 (Q2) are hard-copied every target_update_interval,
        and the aux branch (linear7/8/extra_pred) is never copied."""
        tau = self.cfg["tau"]
        with torch.no_grad():
            for tp, p in zip(self.policy_target.parameters(), self.policy.parameters()):
                tp.copy_(tp * (1.0 - tau) + p * tau)
            if self.has_critic:
                for (n, tp), (_, p) in zip(self.critic_target.named_parameters(), self.critic.named_parameters()):
                    if n[:7] in ("linear1", "linear2", "linear3"):
                        tp.copy_(tp * (1.0 - tau) + p * tau)
                if self.update_step % self.cfg["target_update_interval"] == 0:
                    for (n, tp), (_, p) in zip(self.critic_target.named_parameters(), self.critic.named_parameters()):
                        if n[:7] in ("linear4", "linear5", "linear6"):
                            tp.copy_(p)

    # ---- the BC step (bc.py:40-56) ------------------------------------------------------------------
    def _bc_step(self, batch):
        out = {k: 0.0 for k in LOSS_KEYS}
        self.feat.train(), self.policy.train()
        d = self._load(batch)
        f = self.features(d["cloud"], d["time"], value=False)
        pi, _, _, aux = self.policy.sample(f)
        policy_aux_loss = torch.zeros(1)
        if self.cfg["policy_aux"]:
            gm = d["reward_mask"]
            policy_aux_loss = L.goal_pred_loss(aux[gm, :7], d["goal"][gm, :7])
        em = d["expert_mask"]
        bc_loss = L.pose_bc_loss(pi[em], d["expert_action"][em])
        loss = bc_loss + policy_aux_loss
        self.enc_opt.zero_grad()
        self.policy_opt.zero_grad()
        loss.backward()
        self.policy_opt.step()
        if self.cfg["train_feature"]:
            self.enc_opt.step()
        self._target_updates()
        self.update_step += 1
        out.update(bc_loss=float(bc_loss), policy_grasp_aux_loss=float(policy_aux_loss),
                   reward_mask_num=float(d["reward_mask"].float().sum()), policy_param=_absmax_param(self.policy))
        self.last = dict(pi=pi.detach(), policy_feat=f.detach())
        return out

    # ---- checkpoints (agent.py:282-431) -----------------------------------------------------------
    def _paths(self, output_dir, surfix):
        env = self.cfg.get("env_name", "PandaYCBEnv")
        mk = lambda part: "{}/{}_{}_{}_{}".format(output_dir, self.name, part, env, surfix)  # noqa: E731
        return mk("actor"), mk("critic"), mk("state_feat")

    def save_model(self, step, output_dir="", surfix="latest"):
        """agent.py:282-352: three torch.save files with the reference's dict keys."""
        import os
        os.makedirs(output_dir, exist_ok=True)
        actor, critic, feat = self._paths(output_dir, surfix)
        torch.save({"net": self.policy.state_dict(), "opt": self.policy_opt.state_dict(),
                    "sch": self.policy_sched.state_dict()}, actor)
        if self.has_critic:
            torch.save({"net": self.critic.state_dict(), "opt": self.critic_opt.state_dict(),
                        "sch": self.critic_sched.state_dict()}, critic)
        torch.save({"net": self.state_dicts()["state_feat"], "opt": self.feat_opt.state_dict(),
                    "encoder_opt": self.enc_opt.state_dict(), "sch": self.feat_sched.state_dict(),
                    "encoder_sch": self.enc_sched.state_dict(), "val_encoder_opt": self.venc_opt.state_dict(),
                    "val_encoder_sch": self.venc_sched.state_dict(), "step": step}, feat)

    def load_model(self, output_dir, surfix="latest"):
        """agent.py:354-431 (default flags: no optimiser re-initialisation); targets are hard-updated from the
        loaded online networks (utils.py:761-763)."""
        import os
        actor, critic, feat = self._paths(output_dir, surfix)
        ld = lambda f: torch.load(f, weights_only=False)  # noqa: E731
        if os.path.exists(actor):
            d = ld(actor)
            self.policy.load_state_dict(d["net"])
            self.policy_opt.load_state_dict(d["opt"])
            self.policy_sched.load_state_dict(d["sch"])
            self.policy_target.load_state_dict(self.policy.state_dict())
        if self.has_critic and os.path.exists(critic):
            d = ld(critic)
            self.critic.load_state_dict(d["net"])
            self.critic_opt.load_state_dict(d["opt"])
            self.critic_sched.load_state_dict(d["sch"])
            self.critic_target.load_state_dict(self.critic.state_dict())
        if os.path.exists(feat):
            d = ld(feat)
            self.feat.load_state_dict({k[len("module."):] if k.startswith("module.") else k: v for k, v in d["net"].items()})
            self.feat_opt.load_state_dict(d["opt"])
            self.feat_sched.load_state_dict(d["sch"])
            self.enc_opt.load_state_dict(d["encoder_opt"])
            self.enc_sched.load_state_dict(d["encoder_sch"])
            self.venc_opt.load_state_dict(d["val_encoder_opt"])
            self.venc_sched.load_state_dict(d["val_encoder_sch"])
            self.update_step = d["step"]
            return self.update_step
        return 0

    def step_scheduler(self):
        """agent.py:179-190 — the value-encoder scheduler is never stepped."""
        if self.has_critic:
            self.critic_sched.step()
        self.policy_sched.step()
        if self.cfg["train_feature"] or self.cfg["train_value_feature"]:
            self.feat_sched.step()
            self.enc_sched.step()

    @torch.no_grad()
    def select_action(self, cloud, remain_timestep, eps=None):
        """agent.py:82-125: eval-mode, batch of one; cloud is (C, N+6).  Returns the reference's 4-tuple
        (tanh-mean action[6], log-prob scalar, sampled action[6], aux[7])."""
        self.feat.eval(), self.policy.eval()
        pc = torch.FloatTensor(np.asarray(cloud))[None]
        t = torch.Tensor([remain_timestep]).float()
        f = self.features(pc, t, value=False)
        mean, logp, act, aux = self.policy.sample(f, eps=eps)
        return mean.numpy()[0], logp.numpy()[0][0], act.numpy()[0], aux.numpy()[0]
