"""GPU known-answer tests for the schedule- and mask-dependent corners of the update (SURVEY.md §8(c) list), against
the oracle — which tests/test_oracle_golden.py::test_oracle_schedule_points_match_reference_when_present pins to the
unmodified reference at exactly these points."""
import numpy as np
import pytest
import torch

from tests.test_agent_gpu import _close, _param_report, _sync_from_oracle

pytestmark = pytest.mark.gpu
TOL = {"actor_critic_loss": 3e-3, "critic_grad": 3e-2, "policy_param": 5e-4, "critic_param": 5e-4}


def _pair(**over):
    from gaddpg_b200 import agent as ag
    from oracle.ddpg_cpu import OracleAgent

    return OracleAgent("DDPG", seed=123456, **over), ag.make_agent("DDPG", seed=123456, **over)


def _step_both(ora, mine, batch, u):
    from gaddpg_b200.config import LOSS_KEYS

    o = ora.update_parameters(batch, noise_u=u)
    m = mine.update_parameters(batch, mine.update_step, 0, noise_u=u)
    for k in LOSS_KEYS:
        assert _close(m[k], o[k], rtol=TOL.get(k, 1e-4)), (k, m[k], o[k])
    return o, m


def test_q2_hard_target_copy_at_interval(cuda):
    """utils.py:750-770 / agent.py:204-209: at update_step % 3000 == 0 linear4-6 of the critic target become exact
    copies of the online critic, linear1-3 keep Polyak-averaging, linear7/8/extra_pred are never copied."""
    from gaddpg_b200 import synthetic

    ora, mine = _pair()
    rs = np.random.RandomState(0)
    ora.update_parameters(synthetic.make_batch(8, 512, step=0), noise_u=rs.rand(8, 6).astype(np.float32))
    ora.update_step = 3000
    _sync_from_oracle(mine, ora)
    init_aux = {k: v.clone() for k, v in mine.critic_target.state_dict().items() if k[:7] in ("linear7", "linear8") or "extra" in k}
    _step_both(ora, mine, synthetic.make_batch(8, 512, step=1), rs.rand(8, 6).astype(np.float32))
    assert mine.update_step == ora.update_step == 3001
    tgt, onl = mine.critic_target.state_dict(), mine.critic.state_dict()
    for k in tgt:
        if k[:7] in ("linear4", "linear5", "linear6"):
            assert torch.equal(tgt[k], onl[k]), k                       # hard copy
        elif k[:7] in ("linear1", "linear2", "linear3"):
            assert not torch.equal(tgt[k], onl[k]), k                   # tau = 1e-4 Polyak
        else:
            assert torch.equal(tgt[k], init_aux[k]), k                  # never touched
    rep = _param_report(mine, ora)
    assert rep["critic_target"] < 1e-6 and rep["policy_target"] < 1e-6, rep
    # the next step (3001) must not hard-copy again
    _sync_from_oracle(mine, ora)
    _step_both(ora, mine, synthetic.make_batch(8, 512, step=2), rs.rand(8, 6).astype(np.float32))
    assert _param_report(mine, ora)["critic_target"] < 1e-6


def test_mix_and_noise_schedule_past_first_milestone(cuda):
    """update_step > 4000: mix_policy_ratio 0.1 -> 0.2 (capped by ddpg_coefficients[4]) and TD3 noise ratio 3.0 -> 2.5
    (agent.py:127-139, ddpg.py:61-88); the captured graphs are keyed by the schedule index, so both regimes coexist."""
    from gaddpg_b200 import synthetic

    ora, mine = _pair()
    rs = np.random.RandomState(1)
    for step0 in (1, 4001, 4002):
        ora.update_step = step0
        _sync_from_oracle(mine, ora)
        _step_both(ora, mine, synthetic.make_batch(8, 512, step=step0 % 7), rs.rand(8, 6).astype(np.float32))


def test_empty_masks_give_nan_like_the_reference(cuda):
    """loss.py:17-23 / ddpg.py:119-130: a mean over an empty row selection is NaN in the reference — reproduced, not
    fixed: no positive return -> both goal-auxiliary losses NaN; every row perturbed -> critic loss NaN."""
    from gaddpg_b200 import synthetic

    ora, mine = _pair()
    batch = synthetic.make_batch(8, 512, step=0)
    batch["return_batch"][:] = 0.0
    u = np.full((8, 6), 0.25, np.float32)
    o = ora.update_parameters(batch, noise_u=u)
    m = mine.update_parameters(batch, 1, 0, noise_u=u)
    for k in ("policy_grasp_aux_loss", "critic_grasp_aux_loss"):
        assert np.isnan(o[k]) and np.isnan(m[k]), (k, o[k], m[k])
    assert m["reward_mask_num"] == o["reward_mask_num"] == 0.0
    ora, mine = _pair(policy_aux=False, critic_aux=False)
    batch = synthetic.make_batch(8, 512, step=1)
    batch["perturb_flag_batch"][:] = 1.0
    o = ora.update_parameters(batch, noise_u=u)
    m = mine.update_parameters(batch, 1, 0, noise_u=u)
    assert np.isnan(o["critic_loss"]) and np.isnan(m["critic_loss"]), (o["critic_loss"], m["critic_loss"])


def test_perturbed_rows_are_excluded_from_the_td_loss(cuda):
    """agent.py:229 + ddpg.py:119-130: rows with perturb_flag >= 1 do not enter the smooth-L1 mean; changing their
    reward must not change the critic loss, changing an included row's reward must."""
    from gaddpg_b200 import synthetic

    base = synthetic.make_batch(8, 512, step=3)
    base["perturb_flag_batch"][:] = 0.0
    base["perturb_flag_batch"][[2, 5]] = 1.0
    u = np.full((8, 6), 0.5, np.float32)
    losses = []
    for edit in (None, 2, 0):
        _, mine = _pair(policy_aux=False, critic_aux=False)
        b = {k: (v.copy() if hasattr(v, "copy") else v) for k, v in base.items()}
        if edit is not None:
            b["reward_batch"][edit] += 5.0
        losses.append(mine.update_parameters(b, 1, 0, noise_u=u)["critic_loss"])
    assert losses[0] == losses[1] and losses[0] != losses[2], losses
