"""Float64 referee for the parity tests (test infrastructure; CPU only).

Two fp32 implementations of the same step (the CPU oracle and the CUDA kernels) legitimately differ by more than 1e-4 on
a few quantities: gradients that cross a ReLU / max-pool kink differently, and everything computed after the in-step Adam
update (Adam's first steps move every weight by ~lr * g / (|g| + eps), so weights whose gradient is rounding noise land
up to 2 lr apart between ANY two fp32 evaluations).  Instead of widening tolerances on faith, the tests hold the CUDA
result against a float64 evaluation of the same step from the same state:

        |cuda - f64|  <=  k * |oracle32 - f64|  +  floor

i.e. the kernels are as close to the exact answer as the fp32 reference implementation itself is.

``f64_ops()`` swaps the oracle's ctypes-backed set-abstraction ops (float32 only) for dtype-agnostic torch versions; the
FPS / ball-query *indices* still come from the float32 C oracle on float32-rounded coordinates, so all three sides group
the same points.  ``f64_twin(agent)`` deep-copies an ``OracleAgent`` (weights, BatchNorm buffers, Adam moments, step
counters) into float64.
"""
import contextlib
import copy

import numpy as np
import torch

import oracle.losses_cpu as Lc
import oracle.pointnet2_ops_cpu.pointnet2_utils as U
from oracle.ddpg_cpu import OracleAgent


def _grouping(features, idx):
    B, C, _ = features.shape
    ii = idx.long().view(B, 1, -1).expand(-1, C, -1)
    return torch.gather(features, 2, ii).view(B, C, idx.shape[1], idx.shape[2])


def _gather(features, idx):
    return torch.gather(features, 2, idx.long().unsqueeze(1).expand(-1, features.shape[1], -1))


@contextlib.contextmanager
def f64_ops():
    saved = (U.grouping_operation, U.gather_operation, U.furthest_point_sample, U.ball_query)
    fps, bq, cpt = U.fps_raw, U.ball_query_raw, Lc.control_points
    Lc.control_points = lambda rotz: cpt(rotz).double()   # the float32 control-point table, exactly, as float64 operands
    U.grouping_operation, U.gather_operation = _grouping, _gather
    U.furthest_point_sample = lambda xyz, n: fps(xyz.detach().float().contiguous(), n)
    U.ball_query = lambda r, ns, xyz, new: bq(r, ns, xyz.detach().float().contiguous(), new.detach().float().contiguous())
    try:
        yield
    finally:
        U.grouping_operation, U.gather_operation, U.furthest_point_sample, U.ball_query = saved
        Lc.control_points = cpt


class _Oracle64(OracleAgent):
    def _load(self, batch):
        d = super()._load(batch)
        return {k: (v.double() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in d.items()}

    def update_parameters(self, batch, noise_u=None):
        with f64_ops():   # injected noise is float32 data; everything it meets is float64, so the sum promotes
            return super().update_parameters(batch, noise_u=noise_u)


def f64_twin(agent):
    """Float64 copy of an OracleAgent in its current state (call BEFORE the step that is to be refereed)."""
    t = copy.deepcopy(agent)
    t.__class__ = _Oracle64
    mods = [t.feat, t.policy, t.policy_target] + ([t.critic, t.critic_target] if t.has_critic else [])
    for m in mods:
        m.double()
    t.policy.action_scale, t.policy.action_bias = t.policy.action_scale.double(), t.policy.action_bias.double()
    t.policy_target.action_scale, t.policy_target.action_bias = t.policy_target.action_scale.double(), t.policy_target.action_bias.double()
    opts = [t.feat_opt, t.enc_opt, t.venc_opt, t.policy_opt] + ([t.critic_opt] if t.has_critic else [])
    for o in opts:
        for st in o.state.values():
            for k, v in st.items():
                if torch.is_tensor(v) and v.is_floating_point() and v.dim() > 0:
                    st[k] = v.double()
    return t


def referee(name, cuda, ora32, f64, k=4.0, rel_floor=2e-6, abs_floor=1e-9):
    """Assert |cuda - f64| <= k * |oracle32 - f64| + floors (scalars or tensors, max-norm)."""
    c, o, r = (torch.as_tensor(np.asarray(x, dtype=np.float64)) for x in (cuda, ora32, f64))
    if bool(torch.isnan(r).all()):
        assert bool(torch.isnan(c).all()), (name, "f64 is NaN, CUDA is not")
        return 0.0, 0.0
    ec, eo = float((c - r).abs().max()), float((o - r).abs().max())
    scale = float(r.abs().max())
    assert ec <= k * eo + rel_floor * scale + abs_floor, \
        "%s: |cuda-f64| = %.3e exceeds %g x |oracle32-f64| = %.3e (+ floor %.1e); scale %.3e" % (name, ec, k, eo, rel_floor * scale + abs_floor, scale)
    return ec / (scale + 1e-300), eo / (scale + 1e-300)


def referee_elems(name, cuda, ora32, f64, k=5.0, big=1e-3, floors=(2e-7, 1e-6), frac_slack=0.01, scales=None):
    """Element-wise referee over a LIST of tensors (gradients, post-step parameters).

    Two fp32 evaluations of a ReLU / max-pool network differ from the float64 one in two ways: rounding (every element,
    ~1e-6 of the tensor's scale) and ROUTING — an element whose pre-activation is within rounding of a kink is sent the
    other way, which moves a handful of gradient elements by O(1e-3..1e-1) of the scale on EITHER side (measured:
    tests/diag/diag_grad_f64.py — the fp32 oracle and the CUDA kernels show the same bimodal picture, and which of them
    is hit depends on the input).  A max- or L2-norm over the tensor is therefore decided by a few unlucky elements.
    What is asserted instead, with e = |x - f64| / scale(tensor) over ALL elements:
        median(e_cuda) <= k median(e_ora) + floor,   q90(e_cuda) <= k q90(e_ora) + floor      (rounding level)
        frac(e_cuda > big) <= 4 frac(e_ora > big) + frac_slack                               (routing is rare)
    ``scales``: per-tensor scale (default max |f64|)."""
    ec, eo = [], []
    for i, (c, o, r) in enumerate(zip(cuda, ora32, f64)):
        c, o, r = (torch.as_tensor(np.asarray(x, dtype=np.float64)).flatten() for x in (c, o, r))
        sc = float(r.abs().max()) if scales is None else float(scales[i])
        if sc == 0.0:
            continue
        ec.append((c - r).abs() / sc)
        eo.append((o - r).abs() / sc)
    ec, eo = torch.cat(ec), torch.cat(eo)
    out = {}
    for q, fl in zip((0.5, 0.9), floors):
        qc, qo = float(torch.quantile(ec, q)), float(torch.quantile(eo, q))
        assert qc <= k * qo + fl, "%s: q%d element error vs f64: cuda %.3e > %g x oracle32 %.3e + %.0e" % (name, int(q * 100), qc, k, qo, fl)
        out["q%d" % int(q * 100)] = (qc, qo)
    fc, fo = float((ec > big).float().mean()), float((eo > big).float().mean())
    assert fc <= 4 * fo + frac_slack, "%s: %.3f%% of the elements are > %g off f64 (oracle32: %.3f%%)" % (name, 100 * fc, big, 100 * fo)
    out["frac_big"] = (fc, fo)
    out["max"] = (float(ec.max()), float(eo.max()))
    return out


def hybrid_actor_eval(pre_state, post_state, batch, update_step, cfg_over, dtype=torch.float32):
    """The ACTOR half of an even DDPG step (ddpg.py:164-177 + agent.py:127-139) evaluated by the oracle on prescribed
    weights: policy + policy encoder from ``pre_state`` (they are stepped only at the end of the step), value encoder +
    critic from ``post_state`` (the critic phase has already stepped them when F5 runs).  Returns the actor-critic loss,
    pi, d(ac)/d(pi), and the gradients of the total actor loss w.r.t. policy / policy-encoder parameters and of the
    actor-critic loss w.r.t. the critic parameters (what B2 accumulates onto the clipped critic gradients)."""
    from oracle import losses_cpu as L

    hyb = copy.deepcopy(pre_state)
    for k in post_state["state_feat"]:
        if "value_encoder" in k:
            hyb["state_feat"][k] = post_state["state_feat"][k].detach().cpu().clone()
    hyb["critic"] = {k: v.detach().cpu().clone() for k, v in post_state["critic"].items()}
    o = OracleAgent("DDPG", seed=1, **cfg_over)
    o.load_state_dicts(hyb)
    o.update_step = update_step
    if dtype == torch.float64:
        o = f64_twin(o)
    o.feat.train(), o.policy.train(), o.critic.train()
    ctx = f64_ops() if dtype == torch.float64 else contextlib.nullcontext()
    with ctx:
        d = o._load(batch)
        mix = o._mix_policy_ratio()
        pf = o.features(d["cloud"], d["time"], value=False)
        pi, _, _, aux = o.policy.sample(pf, eps=torch.zeros(pf.shape[0], 6, dtype=dtype))
        vpf = o.features(d["cloud"], d["time"], pi, value=True)
        q1, q2, _ = o.critic(vpf)
        sel = ~d["expert_reward_mask"]
        ac = -mix * torch.min(q1.squeeze(-1)[sel], q2.squeeze(-1)[sel]).mean()
        loss = L.pose_bc_loss(pi[d["expert_mask"]], d["expert_action"][d["expert_mask"]]) * (1 - mix) + ac
        if o.cfg["policy_aux"]:
            gm = d["reward_mask"]
            loss = loss + L.goal_pred_loss(aux[gm, :7], d["goal"][gm, :7])
        dpi = torch.autograd.grad(ac, pi, retain_graph=True)[0]
        for p in list(o.policy.parameters()) + list(o.feat.parameters()) + list(o.critic.parameters()):
            p.grad = None
        loss.backward()
    g = lambda mod: {k: (p.grad.detach().clone() if p.grad is not None else None) for k, p in mod.named_parameters()}  # noqa: E731
    return dict(ac=float(ac), pi=pi.detach(), dpi=dpi.detach(), policy=g(o.policy), encoder=g(o.feat.encoder), critic=g(o.critic),
                q1=q1.detach(), q2=q2.detach())
