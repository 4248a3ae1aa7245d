"""GPU parity (bit-exact): CUDA FPS / ball query / gather / group / row table through the C ABI vs the
CPU oracle (oracle/pointnet2_cpu.c) on the shared adversarial cases and on full-size synthetic clouds."""
import numpy as np
import pytest
import torch

from tests.cases import ball_query_cases, fps_cases

pytestmark = pytest.mark.gpu


def _U():
    from oracle.pointnet2_ops_cpu import pointnet2_utils as U

    return U


def _ops():
    from gaddpg_b200 import ops

    return ops


def test_fps_bit_exact_cases(cuda):
    U, ops = _U(), _ops()
    for name, (xyz, m) in fps_cases().items():
        ref = U.fps_raw(torch.from_numpy(xyz), m)
        got = ops.furthest_point_sample(torch.from_numpy(xyz).to(cuda), m).cpu()
        assert torch.equal(got, ref), name


def test_ball_query_bit_exact_cases(cuda):
    U, ops = _U(), _ops()
    for name, (xyz, new, r, ns) in ball_query_cases().items():
        ref, rcnt = U.ball_query_raw(r, ns, torch.from_numpy(xyz), torch.from_numpy(new), return_cnt=True)
        got, cnt = ops.ball_query(r, ns, torch.from_numpy(xyz).to(cuda), torch.from_numpy(new).to(cuda), return_cnt=True)
        assert torch.equal(got.cpu(), ref), name
        assert torch.equal(cnt.cpu(), rcnt), name


@pytest.mark.parametrize("B,N", [(8, 512), (4, 1024), (3, 2048), (16, 4096), (2, 8192)])
def test_fused_fps_ballquery_on_channel_major_cloud(cuda, B, N):
    """The fused kernel reads the reference's (B,C,N+6) cloud in place; SA1 (r=.02, ns=64) then SA2 level."""
    U, ops = _U(), _ops()
    from gaddpg_b200 import synthetic

    cloud = torch.from_numpy(synthetic.make_batch(B, N, step=3)["point_state_batch"])
    xyz = cloud[:, :3, 6:].transpose(1, 2).contiguous()
    ref_idx = U.fps_raw(xyz, 32)
    ref_new = torch.gather(xyz, 1, ref_idx.long().unsqueeze(-1).expand(-1, -1, 3))
    ref_bq, ref_cnt = U.ball_query_raw(0.02, 64, xyz, ref_new, return_cnt=True)
    g = ops.fps_ballquery_cloud(cloud.to(cuda), 6, 32, 0.02, 64)
    assert torch.equal(g.fps_idx.cpu(), ref_idx)
    assert torch.equal(g.new_xyz.cpu(), ref_new)
    assert torch.equal(g.bq_idx.cpu(), ref_bq)
    assert torch.equal(g.bq_cnt.cpu(), ref_cnt)
    # SA2 level on the 32 centroids
    ref_idx2 = U.fps_raw(ref_new, 32)
    ref_new2 = torch.gather(ref_new, 1, ref_idx2.long().unsqueeze(-1).expand(-1, -1, 3))
    ref_bq2, ref_cnt2 = U.ball_query_raw(0.04, 128, ref_new, ref_new2, return_cnt=True)
    g2 = ops.fps_ballquery_xyz(g.new_xyz, 32, 0.04, 128)
    assert torch.equal(g2.fps_idx.cpu(), ref_idx2)
    assert torch.equal(g2.new_xyz.cpu(), ref_new2)
    assert torch.equal(g2.bq_idx.cpu(), ref_bq2)
    assert torch.equal(g2.bq_cnt.cpu(), ref_cnt2)
    # compact row table: expanding it must reproduce the padded ball-query rows exactly
    for gg, ns in ((g, 64), (g2, 128)):
        rt = ops.row_table(gg.bq_cnt, gg.bq_idx, ns)
        seg_off = rt.seg_off.cpu().numpy()
        cnt = np.maximum(gg.bq_cnt.cpu().numpy().reshape(-1), 1)
        assert np.array_equal(seg_off, np.concatenate([[0], np.cumsum(cnt)]))
        M = int(seg_off[-1])
        row_seg, row_src, row_w = rt.row_seg.cpu().numpy()[:M], rt.row_src.cpu().numpy()[:M], rt.row_w.cpu().numpy()[:M]
        bq = gg.bq_idx.cpu().numpy().reshape(-1, ns)
        assert np.array_equal(row_seg, np.repeat(np.arange(len(cnt)), cnt))
        assert np.array_equal(row_src, np.concatenate([bq[s, : cnt[s]] for s in range(len(cnt))]))
        assert np.isclose(row_w.sum(), len(cnt) * ns)
        first = seg_off[:-1]
        assert np.array_equal(row_w[first], ns - cnt + 1)


def test_gather_group_and_grads(cuda):
    U, ops = _U(), _ops()
    rs = np.random.RandomState(5)
    B, C, N, m, s = 3, 7, 50, 6, 9
    feats = torch.from_numpy(rs.randn(B, C, N).astype(np.float32))
    idx1 = torch.from_numpy(rs.randint(0, N, (B, m)).astype(np.int32))
    idx2 = torch.from_numpy(rs.randint(0, 5, (B, m, s)).astype(np.int32))  # heavy duplicates
    g1 = torch.from_numpy(rs.randn(B, C, m).astype(np.float32))
    g2 = torch.from_numpy(rs.randn(B, C, m, s).astype(np.float32))
    fr = feats.clone().requires_grad_(True)
    o1 = U.gather_operation(fr, idx1)
    o1.backward(g1)
    fd = feats.to(cuda).requires_grad_(True)
    d1 = ops.gather_operation(fd, idx1.to(cuda))
    d1.backward(g1.to(cuda))
    assert torch.equal(d1.cpu(), o1) and torch.allclose(fd.grad.cpu(), fr.grad, atol=1e-6)
    fr = feats.clone().requires_grad_(True)
    o2 = U.grouping_operation(fr, idx2)
    o2.backward(g2)
    fd = feats.to(cuda).requires_grad_(True)
    d2 = ops.grouping_operation(fd, idx2.to(cuda))
    d2.backward(g2.to(cuda))
    assert torch.equal(d2.cpu(), o2) and torch.equal(fd.grad.cpu(), fr.grad)  # same (j,l) summation order


def test_errors_are_loud(cuda):
    ops = _ops()
    with pytest.raises(RuntimeError):
        ops.furthest_point_sample(torch.zeros(1, 60000, 3, device=cuda), 4)  # N > 51072 unsupported, no fallback
    with pytest.raises(RuntimeError):
        ops.fps_ballquery_xyz(torch.zeros(1, 9000, 3, device=cuda), 32, 0.02, 64)  # fused kernel: N <= 8192
    with pytest.raises((RuntimeError, AssertionError)):
        ops.furthest_point_sample(torch.zeros(1, 10, 3), 4)  # CPU tensor


@pytest.mark.parametrize("B,N,m", [(2, 9000, 1024), (1, 20000, 1024), (1, 8193, 33), (3, 4096, 3000), (1, 51072, 16)])
def test_fps_large_clouds_bit_exact(cuda, B, N, m):
    """Clouds beyond the register-resident kernel (N > 8192 or npoint > 2730) take fps_large_kernel: same Spec S1
    arithmetic and tie-break, indices bit-exact against the oracle's literal launch-shape simulation — including exact
    ties (duplicated points) and invalid (near-origin) points."""
    U, ops = _U(), _ops()
    rs = np.random.RandomState(N + m)
    xyz = rs.uniform(-0.2, 0.4, (B, N, 3)).astype(np.float32)
    xyz[:, N // 2: N // 2 + N // 8] = xyz[:, : N // 8]                  # exact duplicates -> tie-break by (bitrev, k)
    xyz[:, rs.choice(N, N // 40, replace=False)] = 0.0                  # |p|^2 <= 1e-3: never selected
    want = U.fps_raw(torch.from_numpy(xyz), m)
    got = ops.furthest_point_sample(torch.from_numpy(xyz).to(cuda), m)
    assert torch.equal(got.cpu(), want)


def test_regularize_pc_point_count_matches_oracle(cuda):
    """core/utils.py:784-812 through the CUDA ops: FPS down-sampling (small and > 8192-point clouds) bit-exact against
    the oracle restatement; the numpy branches reproduce the same seeded draws."""
    from gaddpg_b200.ops import regularize_pc_point_count
    from oracle.utils_cpu import regularize_pc_point_count as oracle_fn
    from tests.test_regularize_cpu import raw_cloud

    for n, m, fp, seed in ((3000, 1024, True, 0), (9000, 1024, True, 1), (20000, 1024, True, 5), (520, 64, True, 2),
                           (3000, 1024, False, 3), (700, 1024, False, 4), (1024, 1024, True, 6)):
        pc = raw_cloud(n, seed)
        np.random.seed(seed)
        want = oracle_fn(pc.copy(), m, use_farthest_point=fp)
        np.random.seed(seed)
        got = regularize_pc_point_count(pc.copy(), m, use_farthest_point=fp)
        assert want.dtype == got.dtype and want.shape == got.shape and np.array_equal(want, got), (n, m, fp)
