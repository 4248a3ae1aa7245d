"""GPU: the device-resident replay buffer (ReplayMemoryB200 + gaddpg_replay_gather, csrc/replay.cu) against the replay
oracle (pinned bit-exactly to the unmodified reference BaseMemory) — byte movement, so everything is BIT-EXACT:
clouds equal float32(reference float64 cloud), the small fields, next-state indices, remaining time."""
import ctypes

import numpy as np
import pytest
import torch

from tests.test_replay_cpu import GOLDEN, KEYS, replay_rounds

pytestmark = pytest.mark.gpu

DEVICE_KEYS = ("point_state_batch", "next_point_state_batch", "action_batch", "expert_action_batch", "goal_batch", "reward_batch",
               "return_batch", "mask_batch", "time_batch", "expert_flag_batch", "perturb_flag_batch", "collide_batch", "grasp_batch",
               "next_action_batch", "next_expert_action_batch", "next_goal_batch", "next_return_batch")


def _same(dev_batch, ora_batch, where):
    for k in DEVICE_KEYS:
        got = dev_batch[k].cpu().numpy()
        want = np.asarray(ora_batch[k]).astype(np.float32)         # float64 clouds: the cast Agent.prepare_data performs
        assert got.shape == want.shape and np.array_equal(got, want), (where, k)
    assert np.array_equal(dev_batch["batch_idx"], ora_batch["batch_idx"]) and dev_batch["batch_idx"].dtype == np.uint8
    assert dev_batch["grasp_sample_batch"].shape == (0, 4, 4)


def test_device_replay_matches_oracle_and_golden(cuda):
    from gaddpg_b200.replay_memory import ReplayMemoryB200
    from oracle.make_golden import REPLAY_CAP, REPLAY_N
    from oracle.replay_cpu import OracleMemory

    fx = np.load(GOLDEN)
    ora = OracleMemory(REPLAY_CAP, uniform_num_pts=REPLAY_N)
    mem = ReplayMemoryB200(REPLAY_CAP, uniform_num_pts=REPLAY_N)
    ora_rounds = list(replay_rounds(ora))
    n = 0
    for (r, e, batch), (_, _, want) in zip(replay_rounds(mem), ora_rounds):
        _same(batch, want, "round %d" % r)
        for k in KEYS:   # and the committed fixture written from the unmodified reference's output
            got = batch[k] if isinstance(batch[k], np.ndarray) else batch[k].cpu().numpy()
            assert np.array_equal(got, fx["r%d:%s" % (r, k)]), (r, k)
        n += 1
    assert n == int(fx["rounds"]) and mem.is_full and mem.cur_idx == ora.cur_idx
    assert np.array_equal(mem.returns, ora.returns) and np.array_equal(mem.episode_map, ora.episode_map)
    assert np.array_equal(mem.point_state.cpu().numpy(), ora.point_state.astype(np.float32))


def test_next_index_and_time_edge_cases(cuda):
    """Terminal transitions point at themselves, the last slot of the store, duplicates inside one minibatch, a failed
    episode with RL off, an all-zero cloud."""
    from gaddpg_b200 import synthetic
    from gaddpg_b200.replay_memory import ReplayMemoryB200
    from oracle.replay_cpu import OracleMemory

    cap = 43
    ora, mem = OracleMemory(cap, uniform_num_pts=128), ReplayMemoryB200(cap, uniform_num_pts=128)
    for e, n in enumerate((5, 7, 30)):
        ep = synthetic.make_episode(n, 128, seed=e)
        ora.add_episode(ep), mem.add_episode(ep)
    assert not mem.is_full and not ora.is_full and mem.cur_idx == ora.cur_idx == 42
    idx = np.array([0, 4, 5, 11, 12, 41, 41, 0, 40])
    d, w = mem.gather(idx), ora.gather(idx)
    _same(d, w, "edge")
    assert d["increment_idx"].cpu().tolist() == [1, 4, 6, 11, 13, 41, 41, 1, 41]
    assert d["time_batch"].cpu().tolist() == [5, 1, 7, 1, 30, 1, 1, 5, 2]
    with pytest.raises(IndexError):
        mem.gather(np.array([cap]))
    # sic (replay_memory.py:223-232): an episode that ends exactly on the last slot wraps cur_idx to buffer_start_idx
    # BEFORE the returns / episode map are written, so they stay 0 for that episode — reproduced, not fixed
    ora2, mem2 = OracleMemory(42, uniform_num_pts=128), ReplayMemoryB200(42, uniform_num_pts=128)
    for e, n in enumerate((5, 7, 30)):
        ep = synthetic.make_episode(n, 128, seed=e)
        ora2.add_episode(ep), mem2.add_episode(ep)
    assert mem2.cur_idx == ora2.cur_idx == 0 and list(mem2.episode_map[12:]) == [0] * 30 == list(ora2.episode_map[12:])
    d2, w2 = mem2.gather(idx), ora2.gather(idx)
    _same(d2, w2, "exact wrap")
    assert d2["increment_idx"].cpu().tolist() == [1, 4, 6, 11, 0, 0, 0, 1, 0]
    assert mem.gather(np.zeros(0, dtype=np.int64))["point_state_batch"].shape == (0, 4, 134)
    bc = ReplayMemoryB200(50, uniform_num_pts=128, RL=False)
    bc.add_episode(synthetic.make_episode(6, 128, seed=9, success=False))
    assert bc.cur_idx == 0
    ep = synthetic.make_episode(3, 128, seed=3)
    ep[1]["point_state"] = np.zeros_like(ep[1]["point_state"])
    bc.add_episode(ep)
    assert bc.cur_idx == 2
    with pytest.raises(ValueError):
        bc.push(dict(ep[0], point_state=np.ones((4, 200))))


def test_c_abi_gather_paths(cuda):
    """Raw C-ABI: clouds-only call (no record table), the scalar path for rows that are not a multiple of 16 bytes,
    index clamping, argument errors."""
    from gaddpg_b200.capi import current_stream, lib

    g = torch.Generator().manual_seed(0)
    for row in (4 * 134, 7 * 33):                                  # vector path | scalar path (231 floats)
        cap, B = 64, 9
        store = torch.randn(cap, row, generator=g).cuda()
        emap = torch.tensor([min(cap - 1, (i // 8) * 8 + 7) for i in range(cap)], dtype=torch.int32).cuda()
        idx = torch.tensor([0, 7, 8, 63, 62, 5, 5, -3, 99], dtype=torch.int32).cuda()
        so, no = torch.zeros(B, row).cuda(), torch.zeros(B, row).cuda()
        inc = torch.zeros(B, dtype=torch.int32).cuda()
        lib.gaddpg_replay_gather(store.data_ptr(), row, None, 0, 0, emap.data_ptr(), cap, idx.data_ptr(), B, so.data_ptr(),
                                 no.data_ptr(), None, inc.data_ptr(), None, None, current_stream())
        cl = idx.clamp(0, cap - 1).long()
        want_inc = torch.minimum(emap[cl].long(), cl + 1)
        assert inc.cpu().tolist() == want_inc.cpu().tolist() == [1, 7, 9, 63, 63, 6, 6, 1, 63]
        assert torch.equal(so, store[cl]) and torch.equal(no, store[want_inc])
    # field-major scatter of the current record (soa_map): column c of sample b -> soa_out[base[c] + b * stride[c]]
    cap, B, W = 32, 5, 32
    store = torch.zeros(cap, 4).cuda()
    rec = torch.arange(cap * W, dtype=torch.float32).view(cap, W).cuda()
    emap = torch.full((cap,), cap - 1, dtype=torch.int32).cuda()
    idx = torch.tensor([3, 0, 31, 7, 7], dtype=torch.int32).cuda()
    m = torch.full((2, W), -1, dtype=torch.int32)
    m[0, 0:6], m[1, 0:6] = torch.arange(6, dtype=torch.int32), 6               # a 6-wide field at offset 0
    m[0, 20], m[1, 20] = 40, 1                                                 # a scalar field at offset 40
    out = torch.full((64,), -7.0).cuda()
    so, no = torch.zeros(B, 4).cuda(), torch.zeros(B, 4).cuda()
    lib.gaddpg_replay_gather(store.data_ptr(), 4, rec.data_ptr(), W, 22, emap.data_ptr(), cap, idx.data_ptr(), B, so.data_ptr(),
                             no.data_ptr(), None, None, m.cuda().data_ptr(), out.data_ptr(), current_stream())
    want = torch.full((64,), -7.0)
    for b, i in enumerate(idx.cpu().tolist()):
        want[b * 6: b * 6 + 6] = rec[i, 0:6].cpu()
        want[40 + b] = rec[i, 20].cpu()
    assert torch.equal(out.cpu(), want)
    raw = lib.load().gaddpg_replay_gather
    z = ctypes.c_void_p(0)
    assert raw(z, 8, z, 0, 0, z, 4, z, 1, z, z, z, z, z, z, z) == -1     # GADDPG_ERR_ARG: null pointers
    assert b"replay_gather" in lib.load().gaddpg_last_error()


def test_update_from_device_replay_equals_update_from_host_batch(cuda):
    """The dict out of ReplayMemoryB200.sample drives DDPGB200.update_parameters exactly like the host dict of the
    reference layout: same scalars bit for bit (the device batch never touches the host)."""
    from gaddpg_b200 import agent as ag, synthetic
    from gaddpg_b200.replay_memory import ReplayMemoryB200
    from oracle.replay_cpu import OracleMemory

    N, B = 512, 8
    ora, mem = OracleMemory(300, uniform_num_pts=N), ReplayMemoryB200(300, uniform_num_pts=N)
    for e in range(12):
        ep = synthetic.make_episode(8 + e, N, seed=100 + e, success=e % 3 != 0)
        ora.add_episode(ep), mem.add_episode(ep)
    from gaddpg_b200.replay_memory import ReplayBatch

    a, b, c = (ag.make_agent("DDPG", seed=123456) for _ in range(3))
    rs = np.random.RandomState(1)
    for step in range(4):       # eager, graph capture, graph replay x2
        np.random.seed(step)
        host = ora.sample(B)
        np.random.seed(step)
        lazy = mem.sample(B)                   # never materialised: gathered straight into the agent's input buffers
        assert isinstance(lazy, ReplayBatch) and not lazy.materialised
        u = rs.rand(B, 6).astype(np.float32)
        ra = a.update_parameters(host, a.update_step, 0, noise_u=u)
        rb = b.update_parameters(lazy, b.update_step, 0, noise_u=torch.from_numpy(u).cuda())
        assert not lazy.materialised
        np.random.seed(step)
        full = mem.sample(B).materialise()     # the dict-of-device-tensors path (what code written for BaseMemory sees)
        assert "point_state_batch" in full and full.materialised
        rc = c.update_parameters(full, c.update_step, 0, noise_u=u)
        for k in ra:
            assert ra[k] == rb[k] == rc[k] or (np.isnan(ra[k]) and np.isnan(rb[k]) and np.isnan(rc[k])), (step, k, ra[k], rb[k], rc[k])
    # the lazy path filled the same input buffers the host path filled
    assert torch.equal(a.cloud, b.cloud) and torch.equal(a.next_cloud, b.next_cloud) and torch.equal(a.vec, b.vec)


def test_full_size_gather_roundtrip(cuda):
    """BASELINE size (B = 256 clouds of 4 x 4102 floats): every gathered row equals the stored row it names."""
    from gaddpg_b200.replay_memory import ReplayMemoryB200

    cap, N, B = 2048, 4096, 256
    mem = ReplayMemoryB200(cap, uniform_num_pts=N)
    mem.point_state.copy_(torch.randn(cap, 4, N + 6, device="cuda"))
    mem.episode_map[:] = (np.arange(cap) // 16) * 16 + 15
    mem.timestep[:] = np.arange(cap) % 16 + 1
    mem.cur_idx, mem.is_full = 0, True
    mem._mark(0, cap)
    np.random.seed(0)
    d = mem.sample(B)
    idx = torch.from_numpy(np.asarray(d["increment_idx"].cpu())).long()
    assert torch.equal(d["next_point_state_batch"], mem.point_state[idx.cuda()])
    np.random.seed(0)
    bi = torch.from_numpy(mem.draw_indices(B)).long()
    assert torch.equal(d["point_state_batch"], mem.point_state[bi.cuda()])
    assert torch.equal(idx, torch.minimum((bi // 16) * 16 + 15, bi + 1))
    assert torch.equal(d["time_batch"].cpu(), (16 + 1 - (bi % 16 + 1)).float())


def test_save_load_interchange_with_reference_format(cuda, tmp_path):
    """replay_memory.py:274-356: the .npz the reference writes (oracle.save is pinned to it) loads into the device store
    and vice versa; minibatches sampled after loading are bit-identical (load re-derives the returns and, like the
    reference, leaves the last stored slot out)."""
    from gaddpg_b200.replay_memory import ReplayMemoryB200
    from oracle.make_golden import REPLAY_CAP, REPLAY_N, replay_episodes
    from oracle.replay_cpu import OracleMemory

    ora, mem = OracleMemory(REPLAY_CAP, uniform_num_pts=REPLAY_N), ReplayMemoryB200(REPLAY_CAP, uniform_num_pts=REPLAY_N)
    for ep in replay_episodes()[:10]:
        ora.add_episode(ep), mem.add_episode(ep)
    d1, d2 = tmp_path / "from_oracle", tmp_path / "from_device"
    ora.save(str(d1), mem.save_data_name)
    mem.save(str(d2))
    for src in (d1, d2):
        o2, m2 = OracleMemory(REPLAY_CAP, uniform_num_pts=REPLAY_N), ReplayMemoryB200(REPLAY_CAP, uniform_num_pts=REPLAY_N)
        o2.load(str(src), mem.save_data_name), m2.load(str(src))
        assert o2.cur_idx == m2.cur_idx and o2.is_full == m2.is_full and o2.total_env_step == m2.total_env_step
        assert np.array_equal(o2.returns, m2.returns) and np.array_equal(o2.episode_map, m2.episode_map)
        np.random.seed(3)
        want = o2.sample(16)
        np.random.seed(3)
        got = m2.sample(16)
        _same(got, want, str(src))
