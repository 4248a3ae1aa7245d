"""CPU property tests (hypothesis): the packed-key closed form the CUDA FPS kernels use equals the oracle's literal
launch-shape simulation on arbitrary sizes with ties and invalid points; ball-query invariants; and the lazy
``ReplayBatch`` contract (pure Python: exercised with a stub memory, no GPU)."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle.pointnet2_ops_cpu import pointnet2_utils as U
from tests.fps_closed_form import fps_closed_form


@settings(max_examples=40, deadline=None, derandomize=True)
@given(n=st.integers(2, 700), m=st.integers(1, 40), seed=st.integers(0, 10_000), grid=st.sampled_from([0, 4, 16]),
       zero_frac=st.sampled_from([0.0, 0.1, 0.9]))
def test_fps_closed_form_equals_simulation(n, m, seed, grid, zero_frac):
    rs = np.random.RandomState(seed)
    xyz = rs.uniform(0.05, 0.4, (n, 3)).astype(np.float32)
    if grid:                                   # snap to a coarse lattice: many exactly equal distances and duplicates
        xyz = (np.round(xyz * grid) / grid + 0.125).astype(np.float32)
    xyz[rs.rand(n) < zero_frac] = 0.0          # |p|^2 <= 1e-3: skipped by the upstream kernel
    m = min(m, n)
    want = U.fps_raw(torch.from_numpy(xyz[None]), m).numpy()[0]
    got = fps_closed_form(xyz, m, U.opt_n_threads(n))
    assert np.array_equal(got, want)
    assert want[0] == 0
    valid = (xyz.astype(np.float64) ** 2).sum(1) > 1.1e-3
    assert all(valid[i] or i == 0 for i in want)          # an invalid point is only ever emitted as the index-0 filler


@settings(max_examples=30, deadline=None, derandomize=True)
@given(n=st.integers(1, 300), m=st.integers(1, 20), ns=st.sampled_from([1, 4, 64]), r=st.sampled_from([0.02, 0.1, 0.5]),
       seed=st.integers(0, 10_000))
def test_ball_query_invariants(n, m, ns, r, seed):
    rs = np.random.RandomState(seed)
    xyz = rs.uniform(0, 0.5, (1, n, 3)).astype(np.float32)
    new = rs.uniform(0, 0.5, (1, m, 3)).astype(np.float32)
    idx, cnt = U.ball_query_raw(r, ns, torch.from_numpy(xyz), torch.from_numpy(new), return_cnt=True)
    idx, cnt = idx.numpy()[0], cnt.numpy()[0]
    r2 = np.float32(r) * np.float32(r)
    for j in range(m):
        d = xyz[0] - new[0, j]
        d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        sure = np.nonzero(d2 < r2 * np.float32(0.999))[0]       # clear of the rounding band of the fused multiply-adds
        h = cnt[j]
        assert 0 <= h <= ns and np.all(np.diff(idx[j, :h]) > 0)                  # ascending scan order, no repeats
        assert np.all(idx[j, h:] == (idx[j, 0] if h else 0))                     # padded with the first hit (0 if none)
        assert set(sure[:max(h - 2, 0)]).issubset(set(idx[j, :h])) or len(sure) == 0
        if h < ns:
            assert set(sure).issubset(set(idx[j, :h]))                           # fewer than nsample hits: every sure hit is listed


class _StubMemory:
    def __init__(self):
        self.calls = 0

    def gather(self, batch_idx):
        self.calls += 1
        return {"point_state_batch": "cloud%s" % [int(i) for i in batch_idx], "reward_batch": "r"}


def test_replay_batch_is_lazy_and_dict_compatible():
    from gaddpg_b200.replay_memory import ReplayBatch

    mem = _StubMemory()
    b = ReplayBatch(mem, [3, 1])
    b["noise_u"] = "u"                          # attaching a key does not force the gather
    assert b["noise_u"] == "u" and mem.calls == 0 and not b.materialised
    assert b["point_state_batch"] == "cloud[3, 1]" and mem.calls == 1 and b.materialised
    assert "reward_batch" in b and set(b.keys()) == {"noise_u", "point_state_batch", "reward_batch"} and len(b) == 3
    assert dict(b.items())["reward_batch"] == "r" and b.get("missing", 7) == 7 and mem.calls == 1
    b2 = ReplayBatch(mem, [0])
    assert sorted(b2) == ["point_state_batch", "reward_batch"] and mem.calls == 2    # iteration materialises once
    assert isinstance(b2, dict)


def test_lazy_batches_are_snapshotted_before_a_write():
    """BaseMemory.sample copies at sample time (replay_memory.py:166-176); a lazy ReplayBatch must therefore
    be gathered before the buffer is written, and the dirty bookkeeping keeps a wrapped ring as two ranges."""
    from gaddpg_b200.replay_memory import ReplayBatch, ReplayMemoryB200

    class M(_StubMemory):
        _settle = ReplayMemoryB200._settle
        _mark = ReplayMemoryB200._mark

        def __init__(self):
            super().__init__()
            self._lazy, self._dirty = {}, []

    mem = M()
    b = ReplayBatch(mem, [5, 6])
    assert len(mem._lazy) == 1 and mem.calls == 0
    mem._settle()                                  # what push / add_episode / load / mark_dirty call first
    assert mem.calls == 1 and b.materialised and not mem._lazy
    assert b["point_state_batch"] == "cloud[5, 6]" and mem.calls == 1
    mem._mark(90, 100), mem._mark(0, 4), mem._mark(3, 7)      # ring wrapped: tail + head, not [0, 100)
    assert mem._dirty == [(0, 7), (90, 100)]
    for k in range(10, 80, 10):
        mem._mark(k, k + 1)
    assert len(mem._dirty) <= 4 and mem._dirty[0][0] == 0 and mem._dirty[-1][1] == 100
