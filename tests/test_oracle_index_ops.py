"""CPU: the oracle's literal FPS simulation vs the packed-key closed form the CUDA kernel uses, and
known-answer tests for ball query / grouping."""
import numpy as np
import torch

from oracle.pointnet2_ops_cpu import pointnet2_utils as U
from tests.cases import ball_query_cases, fps_cases
from tests.fps_closed_form import fps_closed_form


def test_fps_closed_form_matches_launch_shape_simulation():
    for name, (xyz, m) in fps_cases().items():
        N = xyz.shape[1]
        bs = U.opt_n_threads(N)
        ref = U.fps_raw(torch.from_numpy(xyz), m).numpy()
        for b in range(xyz.shape[0]):
            got = fps_closed_form(xyz[b], m, bs)
            assert np.array_equal(got, ref[b]), (name, b, got, ref[b])


def test_fps_block_size_rule():
    assert [U.opt_n_threads(n) for n in (1, 31, 32, 500, 512, 1024, 4096, 8192)] == [1, 16, 32, 256, 512, 512, 512, 512]


def test_fps_known_answers():
    # 4 collinear points: 0 first, then the farthest, then the one maximising the min-distance
    xyz = torch.tensor([[[0.1, 0, 0], [0.2, 0, 0], [0.45, 0, 0], [0.9, 0, 0]]])
    assert U.fps_raw(xyz, 4).tolist() == [[0, 3, 2, 1]]
    # all points invalid -> index 0 repeated
    assert U.fps_raw(torch.full((1, 10, 3), 0.01), 4).tolist() == [[0, 0, 0, 0]]
    # exact tie between k=1 and k=2 (bs=4: bitrev(1)=2 > bitrev(2)=1 -> k=2 wins)
    xyz = torch.tensor([[[0.5, 0.5, 0.5], [0.75, 0.5, 0.5], [0.25, 0.5, 0.5], [0.5, 0.5, 0.5]]])
    assert U.fps_raw(xyz, 2).tolist() == [[0, 2]]


def test_ball_query_known_answers():
    c = ball_query_cases()
    xyz, new, r, ns = c["boundary"]
    idx, cnt = U.ball_query_raw(r, ns, torch.from_numpy(xyz), torch.from_numpy(new), return_cnt=True)
    # hits: k=0 (d=0), k=2 (just inside); k=1 (d2 == r^2) and k=3 (outside) are rejected; 4..7 at origin hit
    assert idx[0, 0].tolist() == [0, 2, 4, 5] and cnt[0, 0].item() == 4
    xyz, new, r, ns = c["nohit_rows"]
    idx, cnt = U.ball_query_raw(r, ns, torch.from_numpy(xyz), torch.from_numpy(new), return_cnt=True)
    assert (idx[:, :5] == 0).all() and (cnt[:, :5] == 0).all() and (cnt[:, 5:] > 0).all()
    xyz, new, r, ns = c["sparse_r0.005_ns64"]
    idx, cnt = U.ball_query_raw(r, ns, torch.from_numpy(xyz), torch.from_numpy(new), return_cnt=True)
    b, j = 0, 3
    h = cnt[b, j].item()
    assert 0 < h < ns
    row = idx[b, j]
    assert (row[h:] == row[0]).all() and (row[:h][1:] > row[:h][:-1]).all()
    d2 = ((torch.from_numpy(xyz[b]) - torch.from_numpy(new[b, j])) ** 2).sum(-1)
    assert set(row[:h].tolist()) == set(torch.nonzero(d2 < np.float32(r) * np.float32(r)).flatten().tolist()[:h])


def test_group_and_grad_with_duplicates():
    feats = torch.arange(2 * 3 * 5, dtype=torch.float32).reshape(2, 3, 5).requires_grad_(True)
    idx = torch.tensor([[[0, 0, 4], [2, 2, 2]], [[1, 3, 1], [4, 0, 0]]], dtype=torch.int32)
    out = U.grouping_operation(feats, idx)
    assert out.shape == (2, 3, 2, 3)
    assert torch.equal(out[1, 2], feats[1, 2][idx[1].long()])
    out.sum().backward()
    want = torch.zeros(2, 3, 5)
    for b in range(2):
        for k in idx[b].flatten().tolist():
            want[b, :, k] += 1
    assert torch.equal(feats.grad, want)
