"""CPU ``pointnet2_ops.pointnet2_utils`` (oracle; SURVEY.md §8 Spec S1-S3).

Each autograd.Function mirrors one op of the upstream ``pointnet2_ops._ext`` module
(SURVEY.md §8(b) "Operator boundary being replaced"); tensors are CPU float32/int32.
"""
import ctypes

import numpy as np
import torch
from torch import nn

from .. import build as _build

_lib = ctypes.CDLL(_build.build())
_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)


def _fp(t):
    return ctypes.cast(t.data_ptr(), _f32p)


def _ip(t):
    return ctypes.cast(t.data_ptr(), _i32p) if t is not None else None


def _chk_f32(t, nd):
    assert t.dtype == torch.float32 and t.dim() == nd and t.device.type == "cpu", (t.dtype, t.shape)
    return t.contiguous()


_lib.oracle_opt_n_threads.restype = ctypes.c_int


def opt_n_threads(n):
    return int(_lib.oracle_opt_n_threads(ctypes.c_int(n)))


def fps_raw(xyz, npoint, block_size=0):
    xyz = _chk_f32(xyz, 3)
    B, N, _ = xyz.shape
    idx = torch.zeros(B, npoint, dtype=torch.int32)
    temp = torch.empty(B, N, dtype=torch.float32)
    _lib.oracle_fps(B, N, npoint, _fp(xyz), _fp(temp), _ip(idx), int(block_size))
    return idx


def ball_query_raw(radius, nsample, xyz, new_xyz, return_cnt=False):
    xyz = _chk_f32(xyz, 3)
    new_xyz = _chk_f32(new_xyz, 3)
    B, N, _ = xyz.shape
    m = new_xyz.shape[1]
    idx = torch.zeros(B, m, nsample, dtype=torch.int32)
    cnt = torch.zeros(B, m, dtype=torch.int32)
    _lib.oracle_ball_query(B, N, m, ctypes.c_float(radius), nsample, _fp(new_xyz), _fp(xyz), _ip(idx), _ip(cnt))
    return (idx, cnt) if return_cnt else idx


class FurthestPointSampling(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xyz, npoint):
        out = fps_raw(xyz, npoint)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class GatherOperation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        features = _chk_f32(features, 3)
        idx = idx.contiguous()
        B, C, N = features.shape
        m = idx.shape[1]
        out = torch.empty(B, C, m, dtype=torch.float32)
        _lib.oracle_gather_points(B, C, N, m, _fp(features), _ip(idx), _fp(out))
        ctx.save_for_backward(idx)
        ctx.N = N
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        B, C, m = grad_out.shape
        g = torch.empty(B, C, ctx.N, dtype=torch.float32)
        _lib.oracle_gather_points_grad(B, C, ctx.N, m, _fp(grad_out), _ip(idx), _fp(g))
        return g, None


gather_operation = GatherOperation.apply


class GroupingOperation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        features = _chk_f32(features, 3)
        idx = idx.contiguous()
        B, C, N = features.shape
        _, m, s = idx.shape
        out = torch.empty(B, C, m, s, dtype=torch.float32)
        _lib.oracle_group_points(B, C, N, m, s, _fp(features), _ip(idx), _fp(out))
        ctx.save_for_backward(idx)
        ctx.N = N
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        B, C, m, s = grad_out.shape
        g = torch.empty(B, C, ctx.N, dtype=torch.float32)
        _lib.oracle_group_points_grad(B, C, ctx.N, m, s, _fp(grad_out), _ip(idx), _fp(g))
        return g, torch.zeros_like(idx)


grouping_operation = GroupingOperation.apply


class BallQuery(torch.autograd.Function):
    @staticmethod
    def forward(ctx, radius, nsample, xyz, new_xyz):
        out = ball_query_raw(radius, nsample, xyz, new_xyz)
        ctx.mark_non_differentiable(out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        return None, None, None, None


ball_query = BallQuery.apply


class QueryAndGroup(nn.Module):
    """Spec S3: ball query, group xyz (relative to the centroid) and features, concat [dxyz, feats]."""

    def __init__(self, radius, nsample, use_xyz=True):
        super().__init__()
        self.radius, self.nsample, self.use_xyz = radius, nsample, use_xyz

    def forward(self, xyz, new_xyz, features=None):
        idx = ball_query(self.radius, self.nsample, xyz, new_xyz)
        xyz_trans = xyz.transpose(1, 2).contiguous()
        grouped_xyz = grouping_operation(xyz_trans, idx)  # (B, 3, npoint, nsample)
        grouped_xyz = grouped_xyz - new_xyz.transpose(1, 2).unsqueeze(-1)
        if features is not None:
            grouped_features = grouping_operation(features, idx)
            if self.use_xyz:
                new_features = torch.cat([grouped_xyz, grouped_features], dim=1)
            else:
                new_features = grouped_features
        else:
            assert self.use_xyz
            new_features = grouped_xyz
        return new_features


class GroupAll(nn.Module):
    """Spec S3 ``npoint=None``: one group holding every point, absolute xyz first."""

    def __init__(self, use_xyz=True):
        super().__init__()
        self.use_xyz = use_xyz

    def forward(self, xyz, new_xyz, features=None):
        grouped_xyz = xyz.transpose(1, 2).unsqueeze(2)
        if features is not None:
            grouped_features = features.unsqueeze(2)
            if self.use_xyz:
                new_features = torch.cat([grouped_xyz, grouped_features], dim=1)
            else:
                new_features = grouped_features
        else:
            new_features = grouped_xyz
        return new_features
