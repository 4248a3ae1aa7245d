"""GPU parity of the fused agents against the CPU oracle (itself pinned bit-exactly to the unmodified reference,
oracle/make_golden.py): identical seeds, replay batches and TD3 noise; every returned scalar, the FPS / ball-query
indices, the network outputs and the post-step parameters must agree.
Tolerance: 1e-4 relative (north_star) on scalars/outputs; parameters are compared with an absolute slack of a few
learning-rate units because Adam turns rounding-level gradient differences into O(lr) parameter differences for
weights whose true gradient is ~0."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _close(a, b, rtol=1e-4, atol=1e-6):
    return (np.isnan(a) and np.isnan(b)) or abs(a - b) <= atol + rtol * abs(b)


def _param_report(agent, ora):
    """max |diff| per network, normalised by lr-units (3e-4 heads, 1e-3 encoders)."""
    out = {}
    sa, so = agent.state_dicts(), ora.state_dicts()
    for net in so:
        worst = 0.0
        for k in so[net]:
            a, b = sa[net][k].detach().cpu().double(), so[net][k].detach().cpu().double()
            if "num_batches_tracked" in k:
                assert int(a) == int(b), (net, k, int(a), int(b))
                continue
            worst = max(worst, float((a - b).abs().max()))
        out[net] = worst
    return out


def _run_pair(policy, B, N, steps, use_graph, **over):
    from gaddpg_b200 import agent as ag, synthetic
    from gaddpg_b200.config import LOSS_KEYS
    from oracle.ddpg_cpu import OracleAgent

    ora = OracleAgent(policy, seed=123456, **over)
    mine = ag.make_agent(policy, seed=123456, **over)
    mine.use_graph = use_graph
    # same seed => same initial weights as the oracle (and therefore as the reference constructors)
    for net, d in ora.state_dicts().items():
        for k, v in d.items():
            assert torch.equal(mine.state_dicts()[net][k].cpu(), v), (net, k)
    rs = np.random.RandomState(42)
    hist = []
    for step in range(steps):
        batch = synthetic.make_batch(B, N, step=step)
        u = rs.rand(B, 6).astype(np.float32)
        o = ora.update_parameters(batch, noise_u=u)
        ora.step_scheduler()
        m = mine.update_parameters(batch, mine.update_step, 0, noise_u=u)
        mine.step_scheduler(mine.update_step)
        for k in LOSS_KEYS:
            assert _close(m[k], o[k], rtol=2e-4 if step else 1e-4), (policy, step, k, m[k], o[k])
        hist.append((o, m))
    rep = _param_report(mine, ora)
    return ora, mine, rep, hist


@pytest.mark.parametrize("use_graph", [False, True])
def test_ddpg_steps_match_oracle(cuda, use_graph):
    ora, mine, rep, hist = _run_pair("DDPG", 8, 512, 4, use_graph)
    # 4 Adam steps: encoders lr 1e-3, heads 3e-4; allow 1.5 lr-units of drift on ill-conditioned weights
    assert rep["policy"] < 1.5 * 3e-4 * 4 and rep["critic"] < 1.5 * 3e-4 * 4, rep
    assert rep["state_feat"] < 1.5 * 1e-3 * 4, rep
    assert rep["policy_target"] < 1e-6 and rep["critic_target"] < 1e-6, rep
    assert hist[1][1]["actor_critic_loss"] != 0.0 and hist[0][1]["actor_critic_loss"] == 0.0


def test_ddpg_first_step_outputs_and_indices(cuda):
    """After ONE step from identical weights nothing has drifted yet: compare the internals tightly."""
    from gaddpg_b200 import agent as ag, synthetic
    from oracle.ddpg_cpu import OracleAgent
    from oracle.pointnet2_ops_cpu import pointnet2_utils as U

    B, N = 8, 512
    ora = OracleAgent("DDPG", seed=123456)
    mine = ag.make_agent("DDPG", seed=123456)
    mine.use_graph = False
    batch = synthetic.make_batch(B, N, step=0)
    u = np.random.RandomState(1).rand(B, 6).astype(np.float32)
    o = ora.update_parameters(batch, noise_u=u)
    m = mine.update_parameters(batch, 1, 0, noise_u=u)
    rel = lambda a, b: float((a.cpu().double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))  # noqa: E731
    assert rel(mine.y, ora.last["y"]) < 1e-4
    assert rel(mine.cc1.qa[:, 0], ora.last["q1"].view(-1)) < 1e-4 and rel(mine.cc1.qa[:, 4], ora.last["q2"].view(-1)) < 1e-4
    assert rel(mine.ctx_v1.feat[:, :513], ora.last["value_feat"]) < 1e-4
    assert rel(mine.pc.pi, ora.last["pi"]) < 1e-4
    # bit-exact indices against the golden fixture written next to the reference-pinned oracle
    fx = np.load(os.path.join(GOLDEN, "index_b8_n512.npz"))
    g = mine.geom_s.lv
    assert np.array_equal(g[0].fps_idx.cpu().numpy(), fx["fps1"]) and np.array_equal(g[0].bq_idx.cpu().numpy(), fx["bq1"])
    assert np.array_equal(g[1].fps_idx.cpu().numpy(), fx["fps2"]) and np.array_equal(g[1].bq_idx.cpu().numpy(), fx["bq2"])
    rep = _param_report(mine, ora)
    assert rep["policy"] < 1.2 * 3e-4 and rep["critic"] < 1.2 * 3e-4 and rep["state_feat"] < 1.2 * 1e-3, rep


def test_golden_scalars_from_reference(cuda):
    """The fixtures were produced by the UNMODIFIED reference (== oracle) in the build container."""
    from gaddpg_b200 import agent as ag, synthetic
    from gaddpg_b200.config import LOSS_KEYS

    for policy in ("DDPG", "BC"):
        fx = np.load(os.path.join(GOLDEN, "%s_b8_n512.npz" % policy.lower()))
        mine = ag.make_agent(policy, seed=int(fx["seed"]))
        for step in range(int(fx["steps"])):
            m = mine.update_parameters(synthetic.make_batch(8, 512, step=step), mine.update_step, 0, noise_u=fx["noise"][step])
            mine.step_scheduler(mine.update_step)
            for i, k in enumerate(LOSS_KEYS):
                assert _close(m[k], float(fx["scalars"][step, i]), rtol=2e-4 if step else 1e-4), (policy, step, k, m[k], fx["scalars"][step, i])
        cloud = synthetic.make_batch(1, 512, step=99)["point_state_batch"][0]
        mean, logp, act, aux = mine.select_action([[cloud, None]], remain_timestep=7, eps=np.zeros((1, 6), np.float32))
        assert np.allclose(mean, fx["sel_mean"], rtol=1e-3, atol=1e-5), (mean, fx["sel_mean"])
        assert np.allclose(aux, fx["sel_aux"], rtol=1e-3, atol=1e-5)


def test_bc_and_no_aux_configs(cuda):
    _, _, rep, _ = _run_pair("BC", 8, 512, 3, True)
    assert rep["policy"] < 1.5 * 3e-4 * 3 and rep["state_feat"] < 1.5 * 1e-3 * 3, rep
    # BASELINE cfg2: DDPG without the auxiliary heads (policy extra_pred_dim=1 unused, no critic aux branch)
    _, _, rep, _ = _run_pair("DDPG", 8, 512, 2, True, policy_aux=False, critic_aux=False)
    assert rep["policy"] < 1.5 * 3e-4 * 2 and rep["critic"] < 1.5 * 3e-4 * 2, rep


def test_six_channel_cloud_variant(cuda):
    """extra_latent=3 ("x6-ch" clouds): policy encoder sees 6 channels, value encoder 6 + the first 4 action channels."""
    from gaddpg_b200 import agent as ag, synthetic
    from gaddpg_b200.config import LOSS_KEYS
    from oracle.ddpg_cpu import OracleAgent

    ora = OracleAgent("DDPG", seed=7, extra_latent=3)
    mine = ag.make_agent("DDPG", seed=7, extra_latent=3)
    for step in range(2):
        batch = synthetic.make_batch(8, 512, step=step, channels=6)
        u = np.random.RandomState(step).rand(8, 6).astype(np.float32)
        o = ora.update_parameters(batch, noise_u=u)
        m = mine.update_parameters(batch, mine.update_step, 0, noise_u=u)
        for k in LOSS_KEYS:
            assert _close(m[k], o[k], rtol=2e-4), (step, k, m[k], o[k])
