"""Pin the oracle against the UNMODIFIED reference and write tests/golden/ (build container only).

    python -m oracle.make_golden            # check + (re)write fixtures
    python -m oracle.make_golden --check    # check only

For DDPG and BC at B=8, N=512 (BASELINE config 1 shape) it builds the reference agent the way
core/train_test_offline.py:305-349 does (through oracle/refstack.py) and ``OracleAgent`` under the same
seed, asserts bit-identical initial weights, runs STEPS update steps on the same synthetic batches with
the same RNG state and asserts every returned scalar and every parameter/buffer is IDENTICAL.  Only then
are the fixtures written: per-step scalars, the TD3 noise, index-op outputs on the first batch,
per-tensor parameter digests and a select_action known answer.  tests/test_oracle_golden.py replays them
with the oracle alone (that is what runs on the GPU box, where /root/reference does not exist).
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gaddpg_b200 import synthetic  # noqa: E402
from oracle import refstack  # noqa: E402
from oracle.ddpg_cpu import LOSS_KEYS, OracleAgent  # noqa: E402
from oracle.pointnet2_ops_cpu import pointnet2_utils as U  # noqa: E402

B, N, STEPS, SEED = 8, 512, 4, 123456
GOLDEN = os.path.join(ROOT, "tests", "golden")


def ref_state_dicts(ref):
    d = {"policy": ref.policy.state_dict(), "policy_target": ref.policy_target.state_dict(),
         "state_feat": ref.state_feature_extractor.state_dict()}
    if hasattr(ref, "critic"):
        d["critic"] = ref.critic.state_dict()
        d["critic_target"] = ref.critic_target.state_dict()
    return d


def assert_same_weights(a, b, what):
    for k in a:
        assert list(a[k].keys()) == list(b[k].keys()), (what, k)
        for n in a[k]:
            assert torch.equal(a[k][n], b[k][n]), (what, k, n)


def digest(sd):
    """name -> (sum, abs-sum) in float64: small, order-independent fingerprint of every tensor."""
    out = {}
    for k, d in sd.items():
        for n, v in d.items():
            v = v.double()
            out[k + "/" + n] = np.array([float(v.sum()), float(v.abs().sum())])
    return out


def run(policy, write):
    ns, ref, _ = refstack.make_reference_agent(policy, seed=SEED)
    ora = OracleAgent(policy, seed=SEED)
    assert_same_weights(ref_state_dicts(ref), ora.state_dicts(), policy + " init")
    scalars = np.zeros((STEPS, len(LOSS_KEYS)))
    noise = np.zeros((STEPS, B, 6), np.float32)
    for step in range(STEPS):
        batch = synthetic.make_batch(B, N, step=step)
        torch.manual_seed(1000 + step)
        r = ref.update_parameters(batch, ref.update_step, 0)
        ref.step_scheduler(ref.update_step)
        torch.manual_seed(1000 + step)
        if policy == "DDPG":
            torch.randn(B, 6)
            noise[step] = torch.rand(B, 6).numpy()  # the draw get_noise_delta will make (utils.py:575)
            torch.manual_seed(1000 + step)
        o = ora.update_parameters(batch)
        ora.step_scheduler()
        for i, k in enumerate(LOSS_KEYS):
            assert r[k] == o[k] or (np.isnan(r[k]) and np.isnan(o[k])), (policy, step, k, r[k], o[k])
            scalars[step, i] = o[k]
    assert_same_weights(ref_state_dicts(ref), ora.state_dicts(), policy + " after %d steps" % STEPS)
    # explicit-noise path of the oracle must reproduce the generator path
    ora2 = OracleAgent(policy, seed=SEED)
    for step in range(STEPS):
        o2 = ora2.update_parameters(synthetic.make_batch(B, N, step=step), noise_u=noise[step])
        ora2.step_scheduler()
        assert all(o2[k] == scalars[step, i] or np.isnan(o2[k]) for i, k in enumerate(LOSS_KEYS)), (policy, step)
    assert_same_weights(ora.state_dicts(), ora2.state_dicts(), policy + " explicit noise")
    # select_action known answer (eval-mode BN, batch of one) vs the reference
    cloud = synthetic.make_batch(1, N, step=99)["point_state_batch"][0]
    torch.manual_seed(77)
    ra = ref.select_action([[cloud, np.zeros((1,), np.float32)]], remain_timestep=7)
    torch.manual_seed(77)
    oa = ora.select_action(cloud, 7)
    for x, y in zip(ra, oa):
        assert np.array_equal(np.asarray(x), np.asarray(y)), (policy, "select_action")
    print("[make_golden] %s: oracle == unmodified reference over %d steps (scalars, weights, select_action)" % (policy, STEPS))
    if write:
        fx = dict(scalars=scalars, noise=noise, loss_keys=np.array(LOSS_KEYS), B=B, N=N, steps=STEPS, seed=SEED,
                  sel_mean=oa[0], sel_logp=np.float32(oa[1]), sel_action=oa[2], sel_aux=oa[3])
        for k, v in digest(ora.state_dicts()).items():
            fx["digest:" + k] = v
        np.savez_compressed(os.path.join(GOLDEN, "%s_b%d_n%d.npz" % (policy.lower(), B, N)), **fx)


def checkpoint_roundtrip(policy="DDPG"):
    """Pin the oracle's save_model / load_model (agent.py:282-431) to the unmodified reference, both directions:
    reference saves after 2 steps -> oracle loads -> third step identical; oracle saves -> reference loads ->
    fourth step identical (weights, Adam moments and schedules all travel through the files)."""
    import tempfile

    ns, ref, _ = refstack.make_reference_agent(policy, seed=SEED)
    ora = OracleAgent(policy, seed=SEED + 1)          # different weights: everything must come from the files
    def both(step, r_agent, o_agent):
        batch = synthetic.make_batch(B, N, step=step)
        torch.manual_seed(1000 + step)
        r = r_agent.update_parameters(batch, r_agent.update_step, 0)
        r_agent.step_scheduler(r_agent.update_step)
        torch.manual_seed(1000 + step)
        o = o_agent.update_parameters(batch)
        o_agent.step_scheduler()
        return r, o
    for step in range(2):
        batch = synthetic.make_batch(B, N, step=step)
        torch.manual_seed(1000 + step)
        ref.update_parameters(batch, ref.update_step, 0)
        ref.step_scheduler(ref.update_step)
    with tempfile.TemporaryDirectory() as d:
        ref.save_model(ref.update_step, d, surfix="pin")
        assert ora.load_model(d, surfix="pin") == ref.update_step
        # the reference hard-updates its targets on load; do the same on the saver so both sides continue equal
        ref.load_model(d, surfix="pin")
        r, o = both(2, ref, ora)
        for k in LOSS_KEYS:
            assert r[k] == o[k] or (np.isnan(r[k]) and np.isnan(o[k])), ("ref->oracle", k, r[k], o[k])
        assert_same_weights(ref_state_dicts(ref), ora.state_dicts(), "ref->oracle checkpoint")
        ora.save_model(ora.update_step, d, surfix="pin2")
        ns2, ref2, _ = refstack.make_reference_agent(policy, seed=SEED + 2)
        assert ref2.load_model(d, surfix="pin2") == ora.update_step
        ora.load_model(d, surfix="pin2")
        r, o = both(3, ref2, ora)
        for k in LOSS_KEYS:
            assert r[k] == o[k] or (np.isnan(r[k]) and np.isnan(o[k])), ("oracle->ref", k, r[k], o[k])
        assert_same_weights(ref_state_dicts(ref2), ora.state_dicts(), "oracle->ref checkpoint")
    print("[make_golden] %s: checkpoint files interchange with the unmodified reference in both directions" % policy)


REPLAY_N, REPLAY_CAP, REPLAY_B = 128, 200, 16


def replay_episodes():
    """The episode stream of the replay fixture: lengths 3..22, ~30 % failures, every 4th episode on-policy (not
    expert); 200 slots so the ring wraps (buffer_start_idx = 0) before the last samples are drawn."""
    from gaddpg_b200 import synthetic
    rs = np.random.RandomState(777)
    return [synthetic.make_episode(int(rs.randint(3, 23)), REPLAY_N, seed=e, success=bool(rs.rand() > 0.3), expert=e % 4 != 3)
            for e in range(24)]


def replay_fixture(write):
    """Pin oracle/replay_cpu.py to the UNMODIFIED reference BaseMemory (replay_memory.py): same episodes in, same numpy
    seed => every key of every sampled minibatch bit-identical, before and after the ring wraps."""
    import importlib

    from oracle.replay_cpu import OracleMemory
    ns = refstack.load()
    ns.config.process_cfg()
    cfg = ns.config.cfg
    old = cfg.RL_TRAIN.uniform_num_pts
    cfg.RL_TRAIN.uniform_num_pts = REPLAY_N
    try:
        ref = importlib.import_module("core.replay_memory").BaseMemory(REPLAY_CAP, cfg, "expert")
    finally:
        cfg.RL_TRAIN.uniform_num_pts = old
    ora = OracleMemory(REPLAY_CAP, uniform_num_pts=REPLAY_N, episode_max_len=cfg.RL_MAX_STEP, gamma=cfg.RL_TRAIN.gamma,
                       buffer_start_idx=cfg.RL_TRAIN.buffer_start_idx, RL=cfg.RL_TRAIN.RL)
    fx, rounds = {}, 0
    for e, ep in enumerate(replay_episodes()):
        ref.add_episode(ep)
        ora.add_episode(ep)
        assert ref.cur_idx == ora.cur_idx and ref.is_full == ora.is_full and ref.upper_idx() == ora.upper_idx(), e
        if ora.upper_idx() <= ora.episode_max_len + 1:
            continue
        np.random.seed(500 + e)
        r = ref.sample(REPLAY_B)
        np.random.seed(500 + e)
        o = ora.sample(REPLAY_B)
        assert set(r.keys()) == set(o.keys()), (sorted(r.keys()), sorted(o.keys()))
        for k in r:
            a, b = np.asarray(r[k]), np.asarray(o[k])
            assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), (e, k, a.dtype, b.dtype, a.shape, b.shape)
        for k in ("batch_idx", "time_batch", "return_batch", "reward_batch", "mask_batch", "expert_flag_batch",
                  "perturb_flag_batch", "action_batch", "goal_batch", "next_goal_batch", "next_return_batch"):
            fx["r%d:%s" % (rounds, k)] = np.asarray(o[k])
        fx["r%d:cloud_digest" % rounds] = np.array([o["point_state_batch"].sum(), np.abs(o["next_point_state_batch"]).sum()])
        fx["r%d:episode" % rounds] = np.array(e)
        rounds += 1
    for name in ("returns", "episode_map", "reward", "timestep", "terminal"):
        assert np.array_equal(getattr(ref, name), getattr(ora, name)), name
    assert ora.is_full, "the fixture is meant to wrap the ring"
    fx["rounds"], fx["returns"], fx["episode_map"] = np.array(rounds), ora.returns, ora.episode_map
    print("[make_golden] replay: oracle == unmodified BaseMemory over %d sampled minibatches (ring wrapped: %s)" % (rounds, ora.is_full))
    if write:
        np.savez_compressed(os.path.join(GOLDEN, "replay_n%d.npz" % REPLAY_N), **fx)


def pin_at_step(step0, steps=2, policy="DDPG", empty_goal_mask=False):
    """Schedule-dependent behaviour the 4-step fixtures never reach, pinned to the unmodified reference by jumping
    ``update_step``: the Q2 hard target copy at multiples of target_update_interval = 3000 (agent.py:204-209,
    utils.py:750-770) and the mix / TD3-noise ratios past the first mix milestone (4000).  ``empty_goal_mask``: a batch
    without a single positive return -> the goal-auxiliary losses are NaN on both sides (loss.py:17-23)."""
    ns, ref, _ = refstack.make_reference_agent(policy, seed=SEED)
    ora = OracleAgent(policy, seed=SEED)
    ref.update_step = ora.update_step = step0
    for step in range(steps):
        batch = synthetic.make_batch(B, N, step=step)
        if empty_goal_mask:
            batch["return_batch"][:] = 0.0
        torch.manual_seed(1000 + step)
        r = ref.update_parameters(batch, ref.update_step, 0)
        torch.manual_seed(1000 + step)
        o = ora.update_parameters(batch)
        for k in LOSS_KEYS:
            assert r[k] == o[k] or (np.isnan(r[k]) and np.isnan(o[k])), (policy, step0, step, k, r[k], o[k])
        if empty_goal_mask:
            assert np.isnan(o["policy_grasp_aux_loss"]) and np.isnan(o["critic_grasp_aux_loss"])
            return
    assert ref.update_step == ora.update_step == step0 + steps
    a, b = ref_state_dicts(ref), ora.state_dicts()
    for k in a:
        for n in a[k]:
            x, y = a[k][n], b[k][n]
            assert torch.equal(x, y) or (torch.isnan(x) == torch.isnan(y)).all(), (k, n)
    print("[make_golden] %s: oracle == unmodified reference for steps %d..%d" % (policy, step0, step0 + steps - 1))


def replay_save_load_pin():
    """BaseMemory.save / load (replay_memory.py:274-356) interchange .npz files with the oracle in both directions; the
    minibatches sampled after loading are bit-identical (load re-derives the returns and drops the last stored slot)."""
    import importlib
    import tempfile

    from oracle.replay_cpu import OracleMemory
    ns = refstack.load()
    ns.config.process_cfg()
    cfg = ns.config.cfg
    BaseMemory = importlib.import_module("core.replay_memory").BaseMemory

    def ref_mem():
        old = cfg.RL_TRAIN.uniform_num_pts
        cfg.RL_TRAIN.uniform_num_pts = REPLAY_N
        try:
            return BaseMemory(REPLAY_CAP, cfg, "expert")
        finally:
            cfg.RL_TRAIN.uniform_num_pts = old

    def same(a, b, what):
        np.random.seed(3)
        x = a.sample(REPLAY_B)
        np.random.seed(3)
        y = b.sample(REPLAY_B)
        assert a.cur_idx == b.cur_idx and a.is_full == b.is_full and a.total_env_step == b.total_env_step, what
        for k in x:
            assert np.array_equal(np.asarray(x[k]), np.asarray(y[k])), (what, k)

    ref, ora = ref_mem(), OracleMemory(REPLAY_CAP, uniform_num_pts=REPLAY_N)
    for ep in replay_episodes()[:10]:
        ref.add_episode(ep), ora.add_episode(ep)
    with tempfile.TemporaryDirectory() as d:
        ref.save(d)
        ora2, ref2 = OracleMemory(REPLAY_CAP, uniform_num_pts=REPLAY_N), ref_mem()
        ora2.load(d, ref.save_data_name), ref2.load(d)
        same(ref2, ora2, "reference file -> both")
    with tempfile.TemporaryDirectory() as d:
        ora.save(d, ref.save_data_name)
        ora3, ref3 = OracleMemory(REPLAY_CAP, uniform_num_pts=REPLAY_N), ref_mem()
        ora3.load(d, ref.save_data_name), ref3.load(d)
        same(ref3, ora3, "oracle file -> both")
    print("[make_golden] replay: save/load files interchange with the unmodified BaseMemory in both directions")


def pin_variant(extra_latent=3, channels=6, steps=2, **overrides):
    """The benchmark's channel variant (BASELINE "x6-ch": ``extra_latent: 3`` => 6-channel clouds, value encoder
    hard-sliced to 10 of 12 channels, networks.py:206-207,238) and the aux-off switch of cfg2, against the unmodified
    reference: same seeds and batches => identical scalars and weights."""
    ns, ref, _ = refstack.make_reference_agent("DDPG", extra_latent=extra_latent, seed=SEED, overrides=overrides)
    ora = OracleAgent("DDPG", seed=SEED, extra_latent=extra_latent, **overrides)
    assert_same_weights(ref_state_dicts(ref), ora.state_dicts(), "variant init")
    for step in range(steps):
        batch = synthetic.make_batch(B, N, step=step, channels=channels)
        torch.manual_seed(1000 + step)
        r = ref.update_parameters(batch, ref.update_step, 0)
        ref.step_scheduler(ref.update_step)
        torch.manual_seed(1000 + step)
        o = ora.update_parameters(batch)
        ora.step_scheduler()
        for k in LOSS_KEYS:
            assert r[k] == o[k] or (np.isnan(r[k]) and np.isnan(o[k])), (extra_latent, overrides, step, k, r[k], o[k])
    assert_same_weights(ref_state_dicts(ref), ora.state_dicts(), "variant after %d steps" % steps)
    print("[make_golden] DDPG extra_latent=%d %s: oracle == unmodified reference over %d steps" % (extra_latent, overrides, steps))


def index_fixture(write):
    """FPS / ball-query outputs of the oracle's C code on the first synthetic batch and on tie-heavy clouds."""
    cloud = torch.from_numpy(synthetic.make_batch(B, N, step=0)["point_state_batch"])
    xyz = cloud[:, :3, 6:].transpose(1, 2).contiguous()
    f1 = U.fps_raw(xyz, 32)
    c1 = torch.gather(xyz, 1, f1.long().unsqueeze(-1).expand(-1, -1, 3))
    b1, n1 = U.ball_query_raw(0.02, 64, xyz, c1, return_cnt=True)
    f2 = U.fps_raw(c1, 32)
    c2 = torch.gather(c1, 1, f2.long().unsqueeze(-1).expand(-1, -1, 3))
    b2, n2 = U.ball_query_raw(0.04, 128, c1, c2, return_cnt=True)
    if write:
        np.savez_compressed(os.path.join(GOLDEN, "index_b%d_n%d.npz" % (B, N)), fps1=f1.numpy(), bq1=b1.numpy().astype(np.int16),
                            cnt1=n1.numpy(), fps2=f2.numpy(), bq2=b2.numpy().astype(np.int8), cnt2=n2.numpy())


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    os.makedirs(GOLDEN, exist_ok=True)
    run("DDPG", not a.check)
    run("BC", not a.check)
    index_fixture(not a.check)
    checkpoint_roundtrip("DDPG")
    replay_fixture(not a.check)
    replay_save_load_pin()
    pin_variant(3, 6, policy_aux=False, critic_aux=False)
    pin_variant(3, 6)
    pin_at_step(2999)
    pin_at_step(4001)
    pin_at_step(1, empty_goal_mask=True)
    print("[make_golden] done")
