"""CPU: the replay oracle (oracle/replay_cpu.py) against the committed fixture tests/golden/replay_n128.npz — written by
oracle/make_golden.py::replay_fixture only after every key of 22 sampled minibatches matched the UNMODIFIED reference
``BaseMemory`` (replay_memory.py) bit for bit — and, when the reference tree is present, against the reference again."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "replay_n128.npz")
KEYS = ("batch_idx", "time_batch", "return_batch", "reward_batch", "mask_batch", "expert_flag_batch", "perturb_flag_batch",
        "action_batch", "goal_batch", "next_goal_batch", "next_return_batch")


def replay_rounds(mem, sample=None):
    """Feed the fixture's episode stream into ``mem``; after each episode (once enough slots are filled) seed numpy and
    sample, exactly as make_golden.replay_fixture does.  Yields (round, episode number, minibatch)."""
    from oracle.make_golden import REPLAY_B, replay_episodes

    rounds = 0
    for e, ep in enumerate(replay_episodes()):
        mem.add_episode(ep)
        if mem.upper_idx() <= mem.episode_max_len + 1:
            continue
        np.random.seed(500 + e)
        yield rounds, e, (sample or mem.sample)(REPLAY_B)
        rounds += 1


def test_oracle_replay_reproduces_golden():
    from oracle.make_golden import REPLAY_CAP, REPLAY_N
    from oracle.replay_cpu import OracleMemory

    fx = np.load(GOLDEN)
    mem = OracleMemory(REPLAY_CAP, uniform_num_pts=REPLAY_N)
    n = 0
    for r, e, batch in replay_rounds(mem):
        assert int(fx["r%d:episode" % r]) == e
        for k in KEYS:
            want, got = fx["r%d:%s" % (r, k)], np.asarray(batch[k])
            assert got.dtype == want.dtype and np.array_equal(got, want), (r, k)
        dig = np.array([batch["point_state_batch"].sum(), np.abs(batch["next_point_state_batch"]).sum()])
        assert np.allclose(dig, fx["r%d:cloud_digest" % r], rtol=1e-12)
        n += 1
    assert n == int(fx["rounds"]) and mem.is_full
    assert np.array_equal(mem.returns, fx["returns"]) and np.array_equal(mem.episode_map, fx["episode_map"])


def test_oracle_replay_semantics():
    """Hand-checkable cases of replay_memory.py:209-272: discounted returns, episode map, next index clamped at the
    episode end, remaining-time remap, uint8 wrap of batch_idx."""
    from gaddpg_b200 import synthetic
    from oracle.replay_cpu import OracleMemory

    mem = OracleMemory(400, uniform_num_pts=128)
    for e, n in enumerate((5, 7, 30)):
        mem.add_episode(synthetic.make_episode(n, 128, seed=e, success=True))
    assert mem.cur_idx == 42 and not mem.is_full
    assert list(mem.episode_map[:5]) == [4] * 5 and list(mem.episode_map[5:12]) == [11] * 7
    # sic: add_episode multiplies the running return by gamma**i (not gamma), so step i from the end carries
    # gamma**(i(i+1)/2) (replay_memory.py:226-229) — reproduced, not fixed
    assert np.allclose(mem.returns[:5], [0.95 ** 10, 0.95 ** 6, 0.95 ** 3, 0.95, 1.0], rtol=1e-6)
    idx = np.array([0, 4, 5, 11, 12, 41])
    d = mem.gather(idx)
    inc = np.minimum(mem.episode_map[idx], idx + 1)
    assert list(inc) == [1, 4, 6, 11, 13, 41]                       # terminal transitions point at themselves
    assert np.array_equal(d["next_point_state_batch"], mem.point_state[inc])
    assert list(d["time_batch"]) == [5, 1, 7, 1, 30, 1]             # steps remaining, counting the current one
    assert d["point_state_batch"].dtype == np.float64 and d["time_batch"].dtype == np.float32
    assert np.uint8(np.array([300]))[0] == 44 and mem.gather(np.array([300]))["batch_idx"][0] == 44
    # a failed episode is dropped when RL is off (replay_memory.py:214-215)
    bc = OracleMemory(100, uniform_num_pts=128, RL=False)
    bc.add_episode(synthetic.make_episode(6, 128, seed=9, success=False))
    assert bc.cur_idx == 0
    # an all-zero cloud is not stored (replay_memory.py:185-189)
    ep = synthetic.make_episode(3, 128, seed=3)
    ep[1]["point_state"] = np.zeros_like(ep[1]["point_state"])
    bc.RL = True
    bc.add_episode(ep)
    assert bc.cur_idx == 2


def test_oracle_replay_matches_unmodified_reference_when_present():
    from oracle import refstack

    if not refstack.available():
        pytest.skip("/root/reference not present (GPU box): covered by the committed fixture")
    from oracle import make_golden

    make_golden.replay_fixture(write=False)


def test_oracle_replay_save_load_matches_reference_when_present():
    from oracle import refstack

    if not refstack.available():
        pytest.skip("/root/reference not present (GPU box)")
    from oracle import make_golden

    make_golden.replay_save_load_pin()


def test_replay_config_from_reference_cfg():
    """ReplayMemoryB200 takes the reference's constructor arguments (buffer_size, cfg): the cfg fields BaseMemory reads
    resolve to the same values (checked against the unmodified experiments/config.py when present)."""
    from types import SimpleNamespace

    from gaddpg_b200.replay_memory import replay_config

    assert replay_config(None) == dict(uniform_num_pts=1024, gamma=0.95, buffer_start_idx=0, RL=True, episode_max_len=20,
                                       save_data_name="data_buffer.npz")
    cfg = SimpleNamespace(RL_TRAIN=dict(uniform_num_pts=512, gamma=0.9, buffer_start_idx=3, RL=False), RL_MAX_STEP=30,
                          RL_SAVE_DATA_NAME="x.npz")
    assert replay_config(cfg) == dict(uniform_num_pts=512, gamma=0.9, buffer_start_idx=3, RL=False, episode_max_len=30, save_data_name="x.npz")
    assert replay_config(cfg, gamma=0.5, episode_max_len=7)["gamma"] == 0.5 and replay_config(cfg, episode_max_len=7)["episode_max_len"] == 7
    with pytest.raises(NotImplementedError):
        replay_config(SimpleNamespace(RL_TRAIN=dict(use_image=True)))
    with pytest.raises(NotImplementedError):
        replay_config(SimpleNamespace(RL_TRAIN=dict(self_supervision=True)))
    from oracle import refstack

    if refstack.available():
        ns = refstack.load()
        ns.config.process_cfg()
        c, ref = replay_config(ns.config.cfg), ns.config.cfg
        assert c["uniform_num_pts"] == ref.RL_TRAIN.uniform_num_pts and c["gamma"] == ref.RL_TRAIN.gamma
        assert c["episode_max_len"] == ref.RL_MAX_STEP and c["save_data_name"] == ref.RL_SAVE_DATA_NAME
        assert c["buffer_start_idx"] == ref.RL_TRAIN.buffer_start_idx and c["RL"] == ref.RL_TRAIN.RL
