"""Python operator surface over the C ABI: the ``pointnet2_ops.pointnet2_utils`` ops the reference calls
(/root/reference/core/utils.py:32,795-796; used inside PointnetSAModule, networks.py:66-81), plus the
fused per-level geometry op.  CUDA tensors only — a CPU tensor raises, like upstream's
``AT_ASSERT(false, "CPU not supported")``."""
from types import SimpleNamespace

import torch

from .capi import current_stream, lib, ptr


def _chk(t, dtype, nd):
    if not t.is_cuda:
        raise RuntimeError("gaddpg_b200: CPU tensors are not supported (no CPU fallback)")
    assert t.dtype == dtype and t.dim() == nd, (t.dtype, tuple(t.shape))
    return t.contiguous()


def fps_raw(xyz, npoint):
    xyz = _chk(xyz, torch.float32, 3)
    B, N, _ = xyz.shape
    idx = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
    lib.gaddpg_fps(ptr(xyz), B, N, npoint, ptr(idx), current_stream())
    return idx


def ball_query(radius, nsample, xyz, new_xyz, return_cnt=False):
    xyz = _chk(xyz, torch.float32, 3)
    new_xyz = _chk(new_xyz, torch.float32, 3)
    B, N, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.empty(B, m, nsample, dtype=torch.int32, device=xyz.device)
    cnt = torch.empty(B, m, dtype=torch.int32, device=xyz.device)
    lib.gaddpg_ball_query(ptr(new_xyz), ptr(xyz), B, N, m, float(radius), nsample, ptr(idx), ptr(cnt), current_stream())
    return (idx, cnt) if return_cnt else idx


class _FPS(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        out = fps_raw(xyz, npoint)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, g):
        return None, None


furthest_point_sample = _FPS.apply


class _Gather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        features = _chk(features, torch.float32, 3)
        idx = _chk(idx, torch.int32, 2)
        B, C, N = features.shape
        m = idx.shape[1]
        out = torch.empty(B, C, m, dtype=torch.float32, device=features.device)
        lib.gaddpg_gather_points(ptr(features), ptr(idx), B, C, N, m, ptr(out), current_stream())
        ctx.save_for_backward(idx)
        ctx.N = N
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        g = g.contiguous()
        B, C, m = g.shape
        out = torch.empty(B, C, ctx.N, dtype=torch.float32, device=g.device)
        lib.gaddpg_gather_points_grad(ptr(g), ptr(idx), B, C, ctx.N, m, ptr(out), current_stream())
        return out, None


gather_operation = _Gather.apply


class _Group(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        features = _chk(features, torch.float32, 3)
        idx = _chk(idx, torch.int32, 3)
        B, C, N = features.shape
        _, m, s = idx.shape
        out = torch.empty(B, C, m, s, dtype=torch.float32, device=features.device)
        lib.gaddpg_group_points(ptr(features), ptr(idx), B, C, N, m, s, ptr(out), current_stream())
        ctx.save_for_backward(idx)
        ctx.N = N
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        g = g.contiguous()
        B, C, m, s = g.shape
        out = torch.empty(B, C, ctx.N, dtype=torch.float32, device=g.device)
        lib.gaddpg_group_points_grad(ptr(g), ptr(idx), B, C, ctx.N, m, s, ptr(out), current_stream())
        return out, None


grouping_operation = _Group.apply


def _fps_bq(base_ptr, sb, sk, sc, B, N, m, radius, nsample, device):
    g = SimpleNamespace(
        fps_idx=torch.empty(B, m, dtype=torch.int32, device=device),
        new_xyz=torch.empty(B, m, 3, dtype=torch.float32, device=device),
        bq_idx=torch.empty(B, m, nsample, dtype=torch.int32, device=device),
        bq_cnt=torch.empty(B, m, dtype=torch.int32, device=device),
    )
    lib.gaddpg_fps_ballquery(base_ptr, sb, sk, sc, B, N, m, float(radius), nsample, ptr(g.fps_idx), ptr(g.new_xyz),
                             ptr(g.bq_idx), ptr(g.bq_cnt), current_stream())
    return g


def fps_ballquery_cloud(cloud, skip, m, radius, nsample):
    """One SA level's geometry straight from the reference's channel-major cloud (B, C, skip+N):
    rows 0..2 are x,y,z; the first ``skip`` (hand) columns are dropped (networks.py:234-235)."""
    cloud = _chk(cloud, torch.float32, 3)
    B, C, Np = cloud.shape
    N = Np - skip
    return _fps_bq(cloud.data_ptr() + 4 * skip, C * Np, 1, Np, B, N, m, radius, nsample, cloud.device)


def fps_ballquery_xyz(xyz, m, radius, nsample):
    xyz = _chk(xyz, torch.float32, 3)
    B, N, _ = xyz.shape
    return _fps_bq(xyz.data_ptr(), N * 3, 3, 1, B, N, m, radius, nsample, xyz.device)


def row_table(bq_cnt, bq_idx, nsample, n_src=None):
    """Compact (duplicate-folded) row list of one SA level; see include/gaddpg_b200.h gaddpg_row_table."""
    S = bq_cnt.numel()
    cap = S * min(nsample, n_src if n_src is not None else nsample)
    dev = bq_cnt.device
    rt = SimpleNamespace(
        S=S, nsample=nsample, cap=cap,
        seg_off=torch.empty(S + 1, dtype=torch.int32, device=dev),
        row_seg=torch.empty(cap, dtype=torch.int32, device=dev),
        row_src=torch.empty(cap, dtype=torch.int32, device=dev),
        row_w=torch.empty(cap, dtype=torch.float32, device=dev),
    )
    lib.gaddpg_row_table(ptr(bq_cnt), ptr(bq_idx), S, nsample, ptr(rt.seg_off), ptr(rt.row_seg), ptr(rt.row_src),
                         ptr(rt.row_w), current_stream())
    return rt


def regularize_pc_point_count(pc, npoints, use_farthest_point=False):
    """Drop-in for ``core.utils.regularize_pc_point_count`` (/root/reference/core/utils.py:784-812), the second FPS call
    site of the reference (environment-side resampling of a raw cloud to ``uniform_num_pts`` points).

    ``pc`` is an (N, C) numpy array (xyz first).  N > npoints: farthest-point down-sampling on the GPU
    (``gaddpg_fps`` + ``gaddpg_gather_points``; clouds above 8192 points take the shared-memory-table FPS kernel) or a
    uniform choice without replacement; N < npoints: over-sampling with replacement.  The random branches use numpy's
    global state exactly like the reference, so a seeded run reproduces its output."""
    import numpy as np

    if pc.shape[0] > npoints:
        if use_farthest_point:
            t = torch.from_numpy(np.ascontiguousarray(pc)).cuda()[None].float()
            idx = furthest_point_sample(t[..., :3].contiguous(), npoints)
            new = gather_operation(t.transpose(1, 2).contiguous(), idx).contiguous()
            pc = new[0].T.detach().cpu().numpy()
        else:
            center_indexes = np.random.choice(range(pc.shape[0]), size=npoints, replace=False)
            pc = pc[center_indexes, :]
    else:
        required = npoints - pc.shape[0]
        if required > 0:
            index = np.random.choice(range(pc.shape[0]), size=required)
            pc = np.concatenate((pc, pc[index, :]), axis=0)
    return pc
