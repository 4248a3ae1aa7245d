"""GPU: rows f2 / f4 of SURVEY.md §8 — the double-buffered feed loop (feed.FeedLoop, the role of trainer.py:202-293 /
train_test_offline.py:117-127), batched inference (test_realworld_ros_final.py:1257-1261) and the device= argument."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _fill(mem, N, episodes=12):
    from gaddpg_b200 import synthetic

    for e in range(episodes):   # mixed expert / on-policy, successful / failed rollouts: no empty loss masks (NaN, a16)
        mem.add_episode(synthetic.make_episode(14, N, seed=e, success=e % 4 != 0, expert=e % 3 != 0))


def _same(a, b):
    """Loss-dict lists equal bit for bit (NaN == NaN: an empty mask gives NaN on both sides, like the reference)."""
    assert len(a) == len(b)
    for i, (x, y) in enumerate(zip(a, b)):
        for k in x:
            assert x[k] == y[k] or (np.isnan(x[k]) and np.isnan(y[k])), (i, k, x[k], y[k])


def test_pipelined_feed_loop_keeps_the_gpu_busy_and_changes_nothing(cuda):
    """Same seeds => the pipelined loop (step i+1 enqueued, index draw included, before step i's scalars are read)
    returns the synchronous loop's loss dicts bit for bit; and the GPU idle gap between consecutive steps, measured with
    CUDA events on the stream, collapses: the device never waits for the host's index draw / staging."""
    from gaddpg_b200 import agent as ag
    from gaddpg_b200.feed import FeedLoop
    from gaddpg_b200.replay_memory import ReplayMemoryB200

    B, N, K = 64, 1024, 24
    res, gaps = {}, {}
    for pipelined in (False, True):
        mem = ReplayMemoryB200(256, uniform_num_pts=N)
        _fill(mem, N)
        agent = ag.make_agent("DDPG", seed=123456)
        torch.manual_seed(5)                      # TD3 noise is drawn on the device inside update_parameters
        np.random.seed(3)
        loop = FeedLoop(agent, mem, B, measure_gaps=True)
        loop.train_iter(6, pipelined=pipelined)   # warm-up: eager pass + graph capture for both step parities
        loop._starts.clear(), loop._ends.clear()
        res[pipelined] = loop.train_iter(K, pipelined=pipelined)
        gaps[pipelined] = np.array(loop.gaps_us())
    _same(res[True], res[False])
    med_sync, med_pipe = float(np.median(gaps[False])), float(np.median(gaps[True]))
    print("GPU idle gap between steps: synchronous %.0f us, pipelined %.0f us (median of %d)" % (med_sync, med_pipe, len(gaps[True])))
    assert med_pipe < 40.0, gaps[True]
    assert med_pipe < 0.5 * med_sync, (med_pipe, med_sync)


def test_host_memory_prefetch_thread_matches_synchronous_loop(cuda):
    """A BaseMemory-style host replay (float64 numpy dict out of sample()): the prefetch thread converts / pins minibatch
    i+1 while step i runs; the sequence of minibatches and therefore every returned scalar is unchanged."""
    from gaddpg_b200 import agent as ag, synthetic
    from gaddpg_b200.feed import FeedLoop
    from oracle.replay_cpu import OracleMemory

    B, N = 16, 512
    out = {}
    for pipelined in (False, True):
        mem = OracleMemory(256, uniform_num_pts=N)
        _fill(mem, N)
        agent = ag.make_agent("DDPG", seed=123456)
        torch.manual_seed(5)
        np.random.seed(3)
        out[pipelined] = FeedLoop(agent, mem, B).train_iter(8, pipelined=pipelined)
    _same(out[True], out[False])


@pytest.mark.parametrize("policy", ["DDPG", "BC"])
def test_prep_streams_for_static_device_batches_change_nothing(cuda, policy):
    """Per-slot input buffers + prep streams (agent.static_device_batches: dicts of finished DEVICE tensors, bench.py's `value`
    leg): the staging copies and the FPS / ball-query / row-table kernels of minibatch i+1 run beside step i, and every returned
    scalar — twelve steps over three distinct minibatches, both step parities — equals the synchronous loop's bit for bit."""
    from gaddpg_b200 import agent as ag, synthetic

    B, N = 32, 1024
    batches = []
    for i in range(3):
        b = synthetic.make_batch(B, N, step=20 + i)
        t = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(cuda) for k, v in b.items()
             if k not in ("grasp_sample_batch", "batch_idx")}
        t["noise_u"] = torch.from_numpy(np.random.RandomState(i).rand(B, 6).astype(np.float32)).to(cuda)
        batches.append(t)
    torch.cuda.synchronize()
    out = {}
    for mode in ("sync", "pipe"):
        agent = ag.make_agent(policy, seed=123456)
        agent.static_device_batches = True
        res, pending = [], None
        for i in range(12):
            b = batches[i % 3]
            h = agent.update_parameters(b, agent.update_step, 0, noise_u=b["noise_u"], defer=(mode == "pipe"))
            agent.step_scheduler(agent.update_step)
            if mode == "sync":
                res.append(h)
                continue
            if pending is not None:
                res.append(pending.result())
            pending = h
        if pending is not None:
            res.append(pending.result())
        out[mode] = res
    _same(out["pipe"], out["sync"])


def test_batched_select_action_and_extract_feature(cuda):
    """test_realworld_ros_final.py:1257-1261: a batch of view-point clouds through extract_feature + policy.sample.  Row b
    of the batched call equals the single-cloud ``select_action`` of cloud b, and both match the oracle within 1e-4."""
    from gaddpg_b200 import agent as ag, synthetic
    from oracle.ddpg_cpu import OracleAgent
    from tests.test_agent_gpu import _sync_from_oracle

    ora = OracleAgent("DDPG", seed=123456)
    mine = ag.make_agent("DDPG", seed=123456)
    for step in range(2):
        ora.update_parameters(synthetic.make_batch(8, 512, step=step), noise_u=np.full((8, 6), 0.5, np.float32))
    _sync_from_oracle(mine, ora)
    Bv = 5
    clouds = synthetic.make_batch(Bv, 512, step=77)["point_state_batch"]
    eps = np.random.RandomState(1).randn(Bv, 6).astype(np.float32)
    remain = np.array([3, 10, 10, 7, 1], np.float32)
    got = [mine.select_action_batch(clouds, remain, eps=eps) for _ in range(3)]     # eager, capture, replay
    for g in got[1:]:
        assert all(np.array_equal(a, b) for a, b in zip(got[0], g))
    pi, logp, act, aux = got[0]
    assert pi.shape == (Bv, 6) and logp.shape == (Bv,) and act.shape == (Bv, 6) and aux.shape == (Bv, 7)
    for b in range(Bv):
        want = ora.select_action(clouds[b], float(remain[b]), eps=torch.from_numpy(eps[b:b + 1]))
        for w, g, name in zip(want, (pi[b], logp[b], act[b], aux[b]), ("mean", "logp", "sample", "aux")):
            assert np.allclose(np.asarray(g), np.asarray(w), rtol=1e-4, atol=1e-5), (b, name, g, w)
        one = mine.select_action([[clouds[b], None]], remain_timestep=float(remain[b]), eps=eps[b:b + 1])
        assert np.allclose(one[0], pi[b], rtol=1e-5, atol=1e-6) and np.allclose(one[3], aux[b], rtol=1e-5, atol=1e-6)
    # the reference's own call style: agent.extract_feature(img, clouds, value=False, time_batch=...) -> policy.sample
    feat = mine.extract_feature(None, torch.from_numpy(clouds).cuda(), value=False, time_batch=torch.from_numpy(remain).cuda())
    assert feat.shape == (Bv, 513)
    mean2, _, _, aux2 = mine.policy.sample(feat, eps=eps)
    assert np.allclose(mean2.cpu().numpy(), pi, rtol=1e-5, atol=1e-6) and np.allclose(aux2.cpu().numpy(), aux, rtol=1e-5, atol=1e-6)
    # value features: cloud (+) action through the value encoder (eval-mode BatchNorm)
    ora.feat.eval()
    a = synthetic.make_batch(Bv, 512, step=78)["action_batch"]
    with torch.no_grad():
        want_v = ora.features(torch.from_numpy(clouds), torch.from_numpy(remain), torch.from_numpy(a), value=True)
    got_v = mine.extract_feature(None, clouds, action_batch=a, value=True, time_batch=remain)
    assert float((got_v.cpu() - want_v).abs().max() / want_v.abs().max()) < 1e-4


def test_migrate_model_copies_bc_checkpoint_to_ddpg_names(tmp_path, cuda):
    """utils.py:319-334: BC_* files become DDPG_* files; a DDPG agent then loads the behaviour-cloned actor + encoder."""
    from gaddpg_b200 import agent as ag, synthetic
    from gaddpg_b200.checkpoint import migrate_model

    bc = ag.make_agent("BC", seed=3)
    bc.update_parameters(synthetic.make_batch(8, 512, step=0), 1, 0)
    src, dst = str(tmp_path / "bc"), str(tmp_path / "ddpg")
    bc.save_model(bc.update_step, output_dir=src)
    made = migrate_model(src, dst)
    assert sorted(d.split("/")[-1] for _, d in made) == ["DDPG_actor_PandaYCBEnv_latest", "DDPG_state_feat_PandaYCBEnv_latest"]
    dd = ag.make_agent("DDPG", seed=4)
    assert dd.load_model(dst) == bc.update_step
    for (k, a), (_, b) in zip(bc.policy.state_dict().items(), dd.policy.state_dict().items()):
        assert torch.equal(a, b), k
    for (k, a), (_, b) in zip(bc._extractor.state_dict().items(), dd._extractor.state_dict().items()):
        assert torch.equal(a, b), k


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_device_argument_is_honoured(cuda):
    """an agent built for cuda:1 while cuda:0 is current must run on cuda:1 with ITS constant tables."""
    from gaddpg_b200 import agent as ag, synthetic

    torch.cuda.set_device(0)
    a0 = ag.make_agent("DDPG", seed=123456, device="cuda:0")
    a1 = ag.make_agent("DDPG", seed=123456, device="cuda:1")
    batch = synthetic.make_batch(8, 512, step=0)
    u = np.random.RandomState(0).rand(8, 6).astype(np.float32)
    r0 = a0.update_parameters(batch, 1, 0, noise_u=u)
    r1 = a1.update_parameters(batch, 1, 0, noise_u=u)
    assert torch.cuda.current_device() == 0
    assert r0 == r1 and a1.cloud.device.index == 1
