"""GPU: the even (actor-critic) DDPG step — F5, ``actor_critic_loss``, the dX-only backward through the value encoder to
d(pi), the accumulated critic gradients, B2 — held to the oracle with a FLOAT64 REFEREE wherever two fp32 evaluations
cannot be expected to agree to 1e-4.

Three instruments (tests/f64ref.py):
  * the plain 1e-4 comparison with the fp32 oracle (north_star's bar) — used for everything computed from identical
    weights through well-conditioned arithmetic;
  * ``hybrid_actor_eval``: the oracle's actor half evaluated on the CUDA agent's OWN post-critic-update weights, so that
    ``actor_critic_loss`` is compared at 1e-4 from identical weights (the in-step Adam update legitimately leaves the two
    sides ~lr apart in noise-gradient directions, which is what the old 3e-3 tolerance papered over);
  * ``referee``: |cuda - f64| <= k |oracle32 - f64| + floor for gradients (ReLU / max-pool kinks: the fp32 oracle itself
    is ~1e-2 away from the float64 gradient d(ac)/d(pi)) and for post-step parameters.
"""
import gc
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

CHAOTIC = ("actor_critic_loss", "critic_grad", "policy_param", "critic_param")


def _close(a, b, rtol=1e-4, atol=1e-6):
    return (np.isnan(a) and np.isnan(b)) or abs(a - b) <= atol + rtol * abs(b)


def _chaotic_ok(where, key, m, o, r):
    """The four scalars computed AFTER the in-step Adam update of the critic + value encoder (ddpg.py:141-143 precedes the
    actor forward, :164-177).  Adam's first steps move every weight by ~lr * g / (|g| + eps): a weight whose gradient is within
    rounding of zero, or one gradient element routed differently through a ReLU / max-pool kink, lands up to 2 lr away on ANY two
    fp32 evaluations, and the actor-critic term / the accumulated critic gradients are evaluated on those weights.  The float64
    twin shows which side happens to sit closer to the exact trajectory is input-dependent (measured: the fp32 oracle 1.6e-5 and
    CUDA 5e-3 from float64 on one configuration, the reverse ordering on others), while the post-step PARAMETER errors of the two
    sides are statistically indistinguishable (referee_elems below: equal medians, 90th percentiles and outlier fractions).
    What the kernels are held to is therefore (a) the same quantities from IDENTICAL weights at 1e-4 (_actor_half_checks) and
    (b) here, a sanity bound of a few learning-rate effects."""
    bound = 2e-2 * max(abs(o), abs(r) if r is not None else 0.0) + 1e-6
    assert abs(m - o) <= bound, (where, key, m, o, r)


def _clone_state(sd):
    return {n: {k: v.detach().cpu().clone() for k, v in d.items()} for n, d in sd.items()}


def _snap_clipped_b1(ora):
    """Record the oracle's critic gradients right after clip_grad_norm_ (ddpg.py:141), i.e. before B2 accumulates the
    actor-loss gradients onto them (agent.py:242-259 reads critic_grad after B2)."""
    box = {}
    orig = ora.critic_opt.step

    def step(*a, **k):
        box["g"] = {n: p.grad.detach().clone() for n, p in ora.critic.named_parameters() if p.grad is not None}
        return orig(*a, **k)

    ora.critic_opt.step = step
    return box, (lambda: setattr(ora.critic_opt, "step", orig))


def _actor_half_checks(name, mine, m, pre, batch, step_no, over, box, with_f64, report, small=False):
    """``mine`` just ran an EVEN step from the synchronised state ``pre``; ``m`` = its scalars."""
    from tests.f64ref import hybrid_actor_eval, referee, referee_elems

    post = _clone_state(mine.state_dicts())
    h32 = hybrid_actor_eval(pre, post, batch, step_no, over, torch.float32)
    # actor_critic_loss from IDENTICAL weights: north_star's 1e-4
    assert _close(m["actor_critic_loss"], h32["ac"], rtol=1e-4), (name, "actor_critic_loss", m["actor_critic_loss"], h32["ac"])
    assert float((mine.pc.pi.cpu() - h32["pi"]).abs().max() / h32["pi"].abs().max()) < 1e-4
    q = torch.stack([h32["q1"].view(-1), h32["q2"].view(-1)], 1)
    qm = torch.stack([mine.cc5.qa[:, 0], mine.cc5.qa[:, 4]], 1).cpu()
    assert float((qm - q).abs().max() / q.abs().max()) < 1e-4, (name, "Q(s, pi(s)) from identical weights")
    # accumulated critic gradient statistic: clipped B1 gradients (identical pre-step weights on both sides) + actor part
    exp_cg = max(float((box["g"][k] + (h32["critic"][k] if h32["critic"].get(k) is not None else 0)).abs().max()) for k in box["g"])
    if not with_f64:
        assert _close(m["critic_grad"], exp_cg, rtol=1e-3), (name, "critic_grad", m["critic_grad"], exp_cg)
        return
    h64 = hybrid_actor_eval(pre, post, batch, step_no, over, torch.float64)
    report["ac"] = referee(name + ":actor_critic_loss", m["actor_critic_loss"], h32["ac"], h64["ac"], rel_floor=2e-5)
    # d(ac)/d(pi): B x 6 numbers, each a sum over every point of the cloud -> routing flips (f64ref.referee_elems) reach all
    # (at B = 8 BatchNorm1d normalises over 8 samples, so one routing flip in one sample moves all 48 numbers: wider floors there;
    # the full-size configurations are held to the tight ones)
    report["dpi"] = referee_elems(name + ":d(ac)/d(pi)", [mine.dpi_ac.cpu().numpy()], [h32["dpi"].numpy()], [h64["dpi"].numpy()],
                                  floors=(1e-3, 5e-3) if small else (2e-5, 1e-4), big=5e-2, frac_slack=0.05)
    exp_cg64 = max(float((box["g"][k].double() + (h64["critic"][k] if h64["critic"].get(k) is not None else 0)).abs().max()) for k in box["g"])
    report["critic_grad"] = referee(name + ":critic_grad", m["critic_grad"], exp_cg, exp_cg64, k=10.0, rel_floor=1e-4)
    for which, mod in (("policy", mine.policy), ("encoder", mine._extractor.encoder)):
        keys = [k for k, _ in mod.named_parameters() if h32[which].get(k) is not None and not k.endswith(("1.0.bias", "1.3.bias"))]
        pm = dict(mod.named_parameters())
        report["grad_" + which] = referee_elems("%s:actor-loss gradients of the %s" % (name, which), [pm[k].grad.detach().cpu().numpy() for k in keys],
                                                [h32[which][k].numpy() for k in keys], [h64[which][k].numpy() for k in keys],
                                                floors=(1e-4, 1e-3) if small else (2e-7, 1e-6))


@pytest.mark.parametrize("over", [dict(), dict(policy_aux=False, critic_aux=False, extra_latent=3)])
def test_even_step_small_with_f64_referee(cuda, over):
    """B = 8, N = 512, four teacher-forced steps (odd, even, odd, even).  Every returned scalar is within 1e-4 of the fp32
    oracle, or — for the four quantities downstream of the in-step Adam update / of a max over gradient elements — no
    further from the float64 twin of the same step than 4x the fp32 oracle itself; post-step parameters likewise, tensor
    by tensor (this replaces the blanket 2.5e-3 bound: only tensors where the ORACLE is that far from float64 get it)."""
    from gaddpg_b200 import agent as ag, synthetic
    from gaddpg_b200.config import LOSS_KEYS
    from oracle.ddpg_cpu import OracleAgent
    from tests.f64ref import f64_twin, referee, referee_elems
    from tests.test_agent_gpu import _sync_from_oracle

    B, N = 8, 512
    ch = 6 if over.get("extra_latent") == 3 else 4
    ora = OracleAgent("DDPG", seed=123456, **over)
    mine = ag.make_agent("DDPG", seed=123456, **over)
    rs = np.random.RandomState(9)
    for step in range(4):
        _sync_from_oracle(mine, ora)
        pre = _clone_state(ora.state_dicts())
        twin = f64_twin(ora)
        batch = synthetic.make_batch(B, N, step=step, channels=ch)
        u = rs.rand(B, 6).astype(np.float32)
        box, undo = _snap_clipped_b1(ora)
        step_no = ora.update_step
        o = ora.update_parameters(batch, noise_u=u)
        undo()
        ora.step_scheduler()
        r = twin.update_parameters(batch, noise_u=u)
        m = mine.update_parameters(batch, mine.update_step, 0, noise_u=u)
        mine.step_scheduler(mine.update_step)
        even = step_no % 2 == 0
        for k in LOSS_KEYS:
            if _close(m[k], o[k], rtol=1e-4):
                continue
            assert k in CHAOTIC, (step, k, m[k], o[k])          # everything else: 1e-4, no excuses
            if even:
                _chaotic_ok("step %d" % step, k, m[k], o[k], r[k])
            else:   # odd steps: only the max-|.| statistics, whose arg-max element may be a noise-gradient weight (+- lr)
                assert k != "actor_critic_loss" and _close(m[k], o[k], rtol=5e-4), (step, k, m[k], o[k])
        # post-step parameters against the float64 twin, in LEARNING-RATE UNITS (Adam's first steps move every weight by
        # ~lr * g / (|g| + eps): a weight whose gradient is rounding noise lands anywhere within +-lr on any fp32 side)
        sm, so, sr = mine.state_dicts(), ora.state_dicts(), twin.state_dicts()
        lr_of = {"policy": 3e-4, "critic": 3e-4, "state_feat": 1e-3}
        for net in ("policy", "critic", "state_feat"):
            keys = [k for k in so[net] if "running" not in k and "num_batches" not in k]
            for k in so[net]:
                if "num_batches_tracked" in k:
                    assert int(sm[net][k]) == int(so[net][k])
            null = [k for k in keys if k.endswith(("1.0.bias", "1.3.bias"))]     # Linear biases in front of BatchNorm1d
            live = [k for k in keys if k not in null]
            A = lambda d, ks: [d[net][k].detach().cpu().double().numpy() for k in ks]  # noqa: E731
            # the float64 referee covers what is stepped from gradients of IDENTICAL weights: everything on odd steps; on even
            # steps the critic and the value encoder only (the policy side is stepped with the actor-critic gradient, which is
            # evaluated on the post-Adam critic: _chaotic_ok)
            ref_keys = [k for k in live if not even or net == "critic" or "value_encoder" in k]
            rep = None
            if ref_keys:
                rep = referee_elems("step %d %s parameters (lr units)" % (step, net), A(sm, ref_keys), A(so, ref_keys), A(sr, ref_keys), k=5.0,
                                    big=0.25, floors=(1e-4, 1e-3), frac_slack=0.002, scales=[lr_of[net]] * len(ref_keys))
            worst = max(float(np.abs(a - b).max()) for a, b in zip(A(sm, live), A(so, live)))
            assert worst <= 2.0 * lr_of[net] * 1.01, (step, net, worst)                       # the mechanistic bound: one lr per side
            for k_ in null:   # exactly-zero true gradient: pure rounding noise through Adam on both sides — only the 2 lr bound
                assert float((sm[net][k_].detach().cpu() - so[net][k_]).abs().max()) <= 2.5e-3, (step, net, k_)
            print("step %d %-10s param error / lr vs f64 (cuda, oracle32): %s; max |cuda - oracle32| = %.2e" % (step, net, rep, worst))
        for net in ("policy_target", "critic_target"):
            for k in so[net]:
                assert float((sm[net][k].detach().cpu() - so[net][k]).abs().max()) < 1e-6, (step, net, k)
        for k in so["state_feat"]:
            if "running" in k:
                a, b = sm["state_feat"][k].detach().cpu().double(), so["state_feat"][k].double()
                # odd steps: two passes from identical weights; even steps: the third value-encoder pass (F5) runs on the post-Adam
                # weights (see _chaotic_ok), which shows at the 1e-4 level in its batch statistics
                assert float((a - b).abs().max()) <= 1e-6 + (2e-3 if even else 1e-4) * float(b.abs().max()), (step, k)
        if even:
            rep = {}
            _actor_half_checks("small step %d" % step, mine, m, pre, batch, step_no, over, box, True, rep, small=True)
            print("even step %d referee (cuda err, oracle32 err) relative to scale: %s" % (step, rep))


@pytest.mark.parametrize("name,B,over", [
    ("cfg2", 256, dict(extra_latent=3, policy_aux=False, critic_aux=False)),   # BASELINE config 2 (the bench workload)
    ("cfg3", 512, dict()),                                                       # BASELINE config 3: goal-aux + grasp-aux losses on
])
def test_full_size_even_step_matches_oracle(cuda, name, B, over):
    """BASELINE.json's full-size configurations, teacher-forced SECOND step (update_step 2: the policy_update_gap branch,
    ddpg.py:170-177): F5 through the value encoder at M ~ 423 k rows, ``actor_critic_loss``, the dX-only backward with the
    per-sample action-channel gradient (sa1_dbc), the accumulated critic gradients and B2 on the production kernels
    (tcgen05 SA1/SA2/SA3, sparse pool backward, multi-stream schedule)."""
    from gaddpg_b200 import agent as ag, synthetic
    from gaddpg_b200.config import LOSS_KEYS
    from oracle.ddpg_cpu import OracleAgent
    from tests.f64ref import f64_twin, referee
    from tests.test_agent_gpu import _sync_from_oracle

    N = 4096
    torch.set_num_threads(os.cpu_count())
    ch = 6 if over.get("extra_latent") == 3 else 4
    ora = OracleAgent("DDPG", seed=123456, **over)
    mine = ag.make_agent("DDPG", seed=123456, **over)
    rs = np.random.RandomState(5)
    ora.update_parameters(synthetic.make_batch(B, N, step=0, channels=ch), noise_u=rs.rand(B, 6).astype(np.float32))  # step 1 (odd)
    ora.step_scheduler()
    _sync_from_oracle(mine, ora)
    assert mine.update_step == 2
    pre = _clone_state(ora.state_dicts())
    # float64 referee at B = 256 only: the float64 autograd graph of a B = 512 step needs > 30 GB of host memory
    with_f64 = B <= 256
    twin = f64_twin(ora) if with_f64 else None
    batch = synthetic.make_batch(B, N, step=1, channels=ch)
    u = rs.rand(B, 6).astype(np.float32)
    box, undo = _snap_clipped_b1(ora)
    o = ora.update_parameters(batch, noise_u=u)
    undo()
    m = mine.update_parameters(batch, mine.update_step, 0, noise_u=u)
    r = twin.update_parameters(batch, noise_u=u) if with_f64 else None
    del twin
    gc.collect()
    assert o["actor_critic_loss"] != 0.0 and m["actor_critic_loss"] != 0.0
    for k in LOSS_KEYS:
        if _close(m[k], o[k], rtol=1e-4):
            continue
        assert k in CHAOTIC, (name, k, m[k], o[k])     # everything else: 1e-4 from identical weights, no excuses
        _chaotic_ok(name, k, m[k], o[k], r[k] if with_f64 else None)
    rel = lambda a, b: float((a.cpu().double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))  # noqa: E731
    assert rel(mine.y, ora.last["y"]) < 1e-4
    assert rel(mine.cc1.qa[:, 0], ora.last["q1"].view(-1)) < 1e-4 and rel(mine.cc1.qa[:, 4], ora.last["q2"].view(-1)) < 1e-4
    rep = {}
    _actor_half_checks(name, mine, m, pre, batch, 2, over, box, with_f64, rep)
    print("%s even step: referee (cuda err, oracle32 err) relative to scale: %s" % (name, rep))
    if not with_f64:
        # max |param| after Adam: each side moved its largest element by at most one learning-rate unit (3e-4)
        for k in ("policy_param", "critic_param"):
            assert abs(m[k] - o[k]) <= 2 * 3e-4, (name, k, m[k], o[k])
