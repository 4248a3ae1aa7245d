"""CPU: the oracle's ``regularize_pc_point_count`` (oracle/utils_cpu.py) — known answers, and identity with the
UNMODIFIED reference function (core/utils.py:784-812) when the reference tree is present."""
import numpy as np
import pytest

from oracle.utils_cpu import regularize_pc_point_count


def raw_cloud(n, seed, channels=4):
    rs = np.random.RandomState(seed)
    pc = rs.uniform(-0.2, 0.2, (n, channels))
    pc[:, 2] += 0.6
    pc[rs.choice(n, max(n // 50, 1), replace=False), :3] = 0.0     # a few origin points: FPS must never select them
    return pc


def test_branches():
    pc = raw_cloud(3000, 0)
    out = regularize_pc_point_count(pc, 1024, use_farthest_point=True)
    assert out.shape == (1024, 4) and out.dtype == np.float32
    rows = {tuple(r) for r in pc.astype(np.float32)}
    assert all(tuple(r) in rows for r in out) and np.array_equal(out[0], pc[0].astype(np.float32))   # FPS starts at point 0
    assert len({tuple(r) for r in out}) == 1024 and (np.abs(out[1:, :3]).sum(1) > 0).all()
    np.random.seed(1)
    sub = regularize_pc_point_count(pc, 1024)
    assert sub.shape == (1024, 4) and sub.dtype == np.float64 and len({tuple(r) for r in sub}) == 1024
    np.random.seed(2)
    up = regularize_pc_point_count(pc[:700], 1024)
    assert up.shape == (1024, 4) and np.array_equal(up[:700], pc[:700])
    assert regularize_pc_point_count(pc[:1024], 1024) is not None and regularize_pc_point_count(pc[:1024], 1024).shape == (1024, 4)


def test_matches_unmodified_reference_when_present():
    from oracle import refstack

    if not refstack.available():
        pytest.skip("/root/reference not present (GPU box)")
    ref = refstack.load().utils.regularize_pc_point_count
    for n, m, fp, seed in ((3000, 1024, True, 0), (9000, 1024, True, 1), (520, 64, True, 2), (3000, 1024, False, 3), (700, 1024, False, 4)):
        pc = raw_cloud(n, seed)
        np.random.seed(seed)
        want = ref(pc.copy(), m, use_farthest_point=fp)
        np.random.seed(seed)
        got = regularize_pc_point_count(pc.copy(), m, use_farthest_point=fp)
        assert want.dtype == got.dtype and np.array_equal(want, got), (n, m, fp)
