"""Host-side orchestration of the fused CUDA encoder and heads (forward + hand-written backward).

What runs here replaces, for one point-cloud encoder, ``PointNetFeature.encode`` (/root/reference/core/
networks.py:217-220: SA1 -> SA2 -> SA3 -> Linear/BN1d/ReLU x2) and, for the heads, ``QNetwork.forward``
(:280-300) / ``GaussianPolicy.forward`` (:339-351) together with their autograd graphs.  Every tensor the
kernels touch is pre-allocated (contexts, scratch) so a whole update step is a fixed launch sequence with no
allocation and no host sync — capturable in a CUDA graph.

Data layout in HBM (all fp32, row-major):
  cloud            (B, C, 6+N)   the reference's channel-major replay layout, consumed in place
  geometry         per cloud and SA level: fps_idx (B,32), new_xyz (B,32,3), bq_idx (B,32,ns), bq_cnt (B,32) and the
                   compact row table seg_off (S+1), row_seg/row_src/row_w (M): shared by every encoder pass on that cloud
  activations      per SA level Y0,Y1,Y2 (M, C_l) = PRE-BatchNorm outputs of the three 1x1 convs over the compact
                   rows only; BN+ReLU are applied by the consumer's prologue, never materialised
  pooled features  (B*32, 128), (B*32, 256), (B, 512)  point-major ("channels last")
  parameters       one arena per network: [params | grads | Adam m | Adam v] (nets.Arena)
"""
import ctypes
import os as _os
from types import SimpleNamespace as NS

import numpy as np
import torch

from . import nets
from .capi import current_stream, lib
from .structs import (EPI_DMASK, EPI_STORE, OP_BNBWD, OP_BNBWD_POOL, OP_BNRELU, OP_PLAIN, STAT_SLOTS, BNTail, NTGroup, NTProblem, Operand,
                      TNProblem, check_sizes, dp, op_bnbwd, op_bnbwd_pool, op_bnrelu, op_plain)

BN_EPS, BN_MOMENTUM = 1e-5, 0.1
_checked = set()   # device indices whose __constant__ tables (action scale / bias, control points) are uploaded


def _init_once(device=None):
    """Per DEVICE: the constant tables live in each GPU's constant bank, and cudaMemcpyToSymbol targets the current one."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _checked:
        if not _checked:
            check_sizes()
        scale = ((nets.ACTION_HIGH - nets.ACTION_LOW) / 2.0).astype(np.float32)
        bias = ((nets.ACTION_HIGH + nets.ACTION_LOW) / 2.0).astype(np.float32)
        cp = np.array([[0, 0, 0], [0, 0, 0], [0.053, -0.0, 0.075], [-0.053, 0, 0.075], [0.053, -0.0, 0.105],
                       [-0.053, 0, 0.105]], dtype=np.float32)
        a = np.pi / 2  # utils.py:826-829: float64 matmul with rotZ(pi/2)[:3,:3], then cast to float32
        rz = np.array([[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]])
        cpz = np.ascontiguousarray(np.matmul(cp, rz).astype(np.float32))
        with torch.cuda.device(idx):
            lib.gaddpg_heads_init(scale.ctypes.data, bias.ctypes.data, cpz.ctypes.data)
        _checked.add(idx)


def _f(device, *shape):
    return torch.zeros(*shape, dtype=torch.float32, device=device)


def _pad4(k):
    return (k + 3) // 4 * 4


class Workspace:
    """Scratch shared by all passes on one device (stream-ordered reuse)."""

    def __init__(self, device):
        _init_once(device)
        self.device = device
        self.stats = _f(device, STAT_SLOTS * 2 * 1024 + 4)
        # ticket word of the BatchNorm tails (gaddpg_bn_tail.counter): zero between launches, one per statistics buffer
        self.counter = self.stats.data_ptr() + 4 * STAT_SLOTS * 2 * 1024
        self.tn_bytes = int(lib.gaddpg_gemm_tn_workspace_bytes())
        self.tn = _f(device, self.tn_bytes // 4)
        self.red = _f(device, 2048)
        self.sa1_ws_bytes = 0
        self.sa1_ws = None
        self._keys = None
        self._sa1f = None

    def keys(self, S, C):
        """Zero-initialised key table of the fused max-pool, (S, C) uint64; the finalize kernel
        re-zeroes what it reads, so one table per stream serves every pass."""
        n = S * C
        if self._keys is None or self._keys.numel() < n:
            self._keys = torch.zeros(n, dtype=torch.int64, device=self.device)
        return self._keys

    def sa1f(self, S, cap):
        """Scratch of the fused SA1 chain (gaddpg_sa1_fused_fwd): raw per-(group, channel) extremes + rows, the partial
        results of groups that straddle two 128-row tiles (one side-table row per tile), and the group -> partial map (all -1
        between calls)."""
        grid = (cap + 127) // 128 + 1   # one side-table row per 128-row tile + a spare row
        b = self._sa1f
        if b is None or b.S < S or b.grid < grid:
            i32 = dict(dtype=torch.int32, device=self.device)
            b = NS(S=S, grid=grid, ext=_f(self.device, S, 128), arg=torch.zeros(S, 128, **i32), part_ext=_f(self.device, grid, 128),
                   part_arg=torch.zeros(grid, 128, **i32), seg_part=torch.full((S,), -1, **i32))
            self._sa1f = b
        return b

    def sa1(self, B, cap):
        jmax = -(-(-(-cap // max(B, 1))) // 256)
        need = (STAT_SLOTS * 64 * 16 + B * 64 * (1 + jmax)) * 4
        if self.sa1_ws is None or self.sa1_ws_bytes < need:
            self.sa1_ws = _f(self.device, need // 4)
            self.sa1_ws_bytes = need
        return self.sa1_ws


class SideStream:
    """A second CUDA stream with its own scratch.  Used (a) for the weight-gradient products of a backward pass, which
    are off the dX critical path, and (b) for a whole independent encoder chain (F2 -> F3 beside F1 -> F4).  Fork /
    join are plain stream waits, so the same code runs eagerly and under CUDA-graph capture (parallel graph branches)."""

    def __init__(self, device):
        self.stream = torch.cuda.Stream(device=device)
        self.ws = Workspace(device)

    def fork(self):
        """Make the side stream wait for everything issued so far on the current stream."""
        self.stream.wait_stream(torch.cuda.current_stream())

    def run(self, fn):
        self.fork()
        with torch.cuda.stream(self.stream):
            fn(self.ws)

    def join(self):
        torch.cuda.current_stream().wait_stream(self.stream)


class BNStage:
    """Staging copy of one encoder's running-statistics arena for ONE pass: with several passes of the same encoder in
    flight at once, each pass writes its batch mean / unbiased variance here and ``apply`` folds them into the real
    running statistics afterwards, pass by pass in the reference's order (F1, F3 | F2, F4) — the values are the
    ones the in-kernel update would have produced."""

    def __init__(self, ef, device):
        self.ef = ef
        self.flat = torch.zeros_like(ef.buffers.p)
        self.rm = {k: ef.buffers.view(k + ".rm", (L.N,), self.flat) for k, L in ef.layers.items()}
        self.rv = {k: ef.buffers.view(k + ".rv", (L.N,), self.flat) for k, L in ef.layers.items()}

    def apply(self):
        ef = self.ef
        lib.gaddpg_bn_running_update(dp(ef.buffers.p), dp(self.flat), ef.buffers.n, BN_MOMENTUM, dp(ef.nbt), ef.nbt.numel(),
                                     current_stream())


# ------------------------------------------------------------------------------------------------
# thin launch helpers
# ------------------------------------------------------------------------------------------------
def _tag(kind, M_max, M_dev, flops_per_row, bytes_per_row, bytes_fixed=0.0):
    from . import capi

    if capi.PROFILE is not None:
        capi.PROFILE.tag = dict(kind=kind, M_max=M_max, M_dev=M_dev, flops_per_row=flops_per_row, bytes_per_row=bytes_per_row,
                                bytes_fixed=bytes_fixed)


def nt(problems, amode, emode):
    g = NTGroup()
    for i, p in enumerate(problems):
        g.p[i] = p
    p0 = problems[0]
    # algorithmic work of the launch: 2*N*K flops per row; bytes = A row (+Y for BN-backward, +Yprev for the mask
    # epilogue) read once + C row written once, weights read once
    rd = p0.K * (2 if amode == OP_BNBWD else 1) + (p0.N if emode == EPI_DMASK else 0)
    from . import capi
    path = ("ffma", "tc", "kc", "sk")[lib.gaddpg_gemm_nt_path(ctypes.byref(g), len(problems), amode, emode)] if capi.PROFILE is not None else ""
    _tag("%snt[%dx%d,a%d,e%d]" % (path + ":" if path else "", p0.N, p0.K, amode, emode), p0.M_max, p0.M_dev,
         sum(2.0 * q.N * q.K for q in problems), 4.0 * len(problems) * (rd + p0.N), 4.0 * sum(q.N * q.K for q in problems))
    lib.gaddpg_gemm_nt(ctypes.byref(g), len(problems), amode, emode, current_stream())


def nt_problem(A, Bw, ldb, C, ldc, M_max, M_dev, N, K, bias=None, relu=0, stats=None, srw=None, Yprev=None, ldyp=0,
               pbn=None, pool_keys=None, pool_seg=None, pool_gamma=None, no_store=0, w_split=None, tail=None):
    p = NTProblem(A=A, Bw=dp(Bw) if torch.is_tensor(Bw) else Bw, ldb=ldb, bias=dp(bias), C=dp(C) if torch.is_tensor(C) else C,
                  ldc=ldc, M_max=M_max, M_dev=M_dev, N=N, K=K, relu=relu, stats=dp(stats), srw=dp(srw),
                  Yprev=dp(Yprev) if torch.is_tensor(Yprev) else Yprev, ldyp=ldyp)
    if pbn is not None:
        p.psc, p.psh, p.pmean, p.prstd = dp(pbn.scale), dp(pbn.shift), dp(pbn.mean), dp(pbn.rstd)
    if pool_keys is not None:
        p.pool_keys, p.pool_seg, p.pool_gamma, p.no_store = dp(pool_keys), dp(pool_seg), dp(pool_gamma), no_store
    if w_split is not None:
        p.Bw_hi, p.Bw_lo = w_split
    if tail is not None:
        p.tail = tail
    return p


SPARSE_POOL = True  # SA1 max-pool backward without the dense gradient tensor (False: dense gaddpg_pool_bwd everywhere)
SIDE = None  # SideStream carrying the weight-gradient products of the backward pass being issued (None: same stream)


def tn(ws, P, Q, pmode, qmode, M_max, M_dev, N, K, dW, ldd, Ntrue, Ktrue, rot=0, dbias=None, accumulate=0):
    prob = TNProblem(P=P, Q=Q, M_max=M_max, M_dev=M_dev, N=N, K=K)

    def launch(w):
        _tag("tn[%dx%d,p%d,q%d]" % (N, K, pmode, qmode), M_max, M_dev, 2.0 * N * K,
             4.0 * (N * (2 if pmode == OP_BNBWD else 1) + K), 4.0 * N * K)
        lib.gaddpg_gemm_tn(ctypes.byref(prob), pmode, qmode, dp(dW), ldd, Ntrue, Ktrue, rot, dp(dbias), accumulate,
                           dp(w.tn), w.tn_bytes, current_stream())

    if SIDE is None:
        launch(ws)
    else:
        SIDE.run(launch)  # reads only buffers that are complete at this point and that the dX chain never rewrites


class side_dw:
    """``with side_dw(side): <backward calls>`` — weight gradients of the enclosed backward passes go to ``side``;
    joined on exit, so gradients are complete (and the scratch reusable) when the block ends."""

    def __init__(self, side):
        self.side = side

    def __enter__(self):
        global SIDE
        self.prev, SIDE = SIDE, self.side
        return self

    def __exit__(self, *exc):
        global SIDE
        if self.side is not None:
            self.side.join()
        SIDE = self.prev


# BatchNorm finalize as the tail of the kernel that produced the statistics (gaddpg_bn_tail; csrc/bn_tail.cuh): built, parity-green
# (tests/test_bn_tail_gpu.py) and measured at cfg2 (profiles/r2_bn_tail.md): 252 instead of 322 launches per step, but the step
# is 2.4 % SLOWER (211.4 vs 216.7 steps/s, A/B in one session): what a finalize costs on the dependency chain is its two
# dependent L2 round trips (slots, then constants / running statistics) plus the ticket, not its launch — and one CTA does
# them no faster than the C/32-CTA kernel.  OFF by default; GADDPG_BN_TAIL=1 selects it.
FUSED_BN_TAIL = _os.environ.get("GADDPG_BN_TAIL", "0") == "1"


def fwd_tail(ws, C, count, bnp, st, train, stage=None):
    """gaddpg_bn_tail of a training-mode forward BatchNorm: the kernel that writes the statistics of this layer into ws.stats also
    finalizes them (scale / shift / mean / rstd into ``st``, running statistics into the layer or the pass's BNStage).  None in
    eval mode (no batch statistics: call bn_fwd) or with FUSED_BN_TAIL off (then bn_fwd_after runs the separate kernel)."""
    if not train or not FUSED_BN_TAIL:
        return None
    t = BNTail(kind=1, count=float(count), a=dp(bnp.gamma), b=dp(bnp.beta), eps=BN_EPS, o0=dp(st.scale), o1=dp(st.shift),
               o2=dp(st.mean), o3=dp(st.rstd), counter=ws.counter)
    if stage is not None:   # deferred running-stat update: stage the batch statistics (momentum 1 = verbatim)
        t.momentum, t.running_mean, t.running_var = 1.0, dp(stage.rm[bnp.key]), dp(stage.rv[bnp.key])
    else:
        t.momentum, t.running_mean, t.running_var, t.num_batches_tracked = BN_MOMENTUM, dp(bnp.rm), dp(bnp.rv), dp(bnp.nbt)
    return t


def bn_fwd_after(tail, ws, C, count, bnp, st, train, stage=None):
    """The separate finalize launch, unless the producer carried ``tail``."""
    if tail is None:
        bn_fwd(ws, C, count, bnp, st, train, stage)


def bwd_tail(ws, C, count, bnp, st, bb, want_grads, accumulate=0):
    """gaddpg_bn_tail of a BatchNorm backward: m1 / m2 / g into ``bb``, dgamma / dbeta into the gradient arena."""
    if not FUSED_BN_TAIL:
        return None
    return BNTail(kind=2, accumulate=accumulate, count=float(count), a=dp(bnp.gamma), b=dp(st.rstd), o0=dp(bb.g), o1=dp(bb.m1),
                  o2=dp(bb.m2), dgamma=dp(bnp.dgamma) if want_grads else None, dbeta=dp(bnp.dbeta) if want_grads else None,
                  counter=ws.counter)


def _tp(tail):
    return None if tail is None else ctypes.byref(tail)


def bn_fwd(ws, C, count, bnp, st, train, stage=None):
    if stage is not None and train:  # deferred running-stat update: stage the batch statistics (momentum 1 = verbatim)
        lib.gaddpg_bn_finalize_fwd(dp(ws.stats), C, float(count), dp(bnp.gamma), dp(bnp.beta), BN_EPS, 1.0, dp(stage.rm[bnp.key]),
                                   dp(stage.rv[bnp.key]), None, 1, dp(st.scale), dp(st.shift), dp(st.mean), dp(st.rstd),
                                   current_stream())
        return
    lib.gaddpg_bn_finalize_fwd(dp(ws.stats), C, float(count), dp(bnp.gamma), dp(bnp.beta), BN_EPS, BN_MOMENTUM, dp(bnp.rm),
                               dp(bnp.rv), dp(bnp.nbt), 1 if train else 0, dp(st.scale), dp(st.shift), dp(st.mean),
                               dp(st.rstd), current_stream())


def bn_bwd(ws, C, count, bnp, st, bb, want_grads, accumulate=0):
    lib.gaddpg_bn_finalize_bwd(dp(ws.stats), C, float(count), dp(bnp.gamma), dp(st.rstd), dp(bb.g), dp(bb.m1), dp(bb.m2),
                               dp(bnp.dgamma) if want_grads else None, dp(bnp.dbeta) if want_grads else None, accumulate,
                               current_stream())


# ------------------------------------------------------------------------------------------------
# geometry: FPS + ball query + compact row tables for both SA levels of one cloud
# ------------------------------------------------------------------------------------------------
class Geometry:
    def __init__(self, B, N, device, npoint=32, r1=0.02, ns1=64, r2=0.04, ns2=128):
        self.B, self.N, self.npoint = B, N, npoint
        self.r = (r1, r2)
        self.ns = (ns1, ns2)
        i32 = dict(dtype=torch.int32, device=device)
        S = B * npoint
        self.lv = []
        for ns, nsrc in ((ns1, N), (ns2, npoint)):
            cap = S * min(ns, nsrc)
            self.lv.append(NS(fps_idx=torch.zeros(B, npoint, **i32), new_xyz=_f(device, B, npoint, 3),
                              bq_idx=torch.zeros(B, npoint, ns, **i32), bq_cnt=torch.zeros(B, npoint, **i32),
                              seg_off=torch.zeros(S + 1, **i32), row_seg=torch.zeros(cap, **i32),
                              row_src=torch.zeros(cap, **i32), row_w=_f(device, cap), cap=cap, S=S, ns=ns, nsrc=nsrc))
        for l in self.lv:
            l.M_dev = l.seg_off.data_ptr() + 4 * l.S

    def build(self, cloud, skip):
        """cloud (B, C, skip+N) float32 cuda, channel rows 0..2 = x,y,z (networks.py:234-242)."""
        B, C, Np = cloud.shape
        assert B == self.B and Np - skip == self.N and cloud.is_contiguous() and cloud.dtype == torch.float32
        st = current_stream()
        l1, l2 = self.lv
        lib.gaddpg_fps_ballquery(cloud.data_ptr() + 4 * skip, C * Np, 1, Np, B, self.N, self.npoint, self.r[0], l1.ns,
                                 dp(l1.fps_idx), dp(l1.new_xyz), dp(l1.bq_idx), dp(l1.bq_cnt), st)
        lib.gaddpg_row_table(dp(l1.bq_cnt), dp(l1.bq_idx), l1.S, l1.ns, dp(l1.seg_off), dp(l1.row_seg), dp(l1.row_src),
                             dp(l1.row_w), st)
        lib.gaddpg_fps_ballquery(dp(l1.new_xyz), self.npoint * 3, 3, 1, B, self.npoint, self.npoint, self.r[1], l2.ns,
                                 dp(l2.fps_idx), dp(l2.new_xyz), dp(l2.bq_idx), dp(l2.bq_cnt), st)
        lib.gaddpg_row_table(dp(l2.bq_cnt), dp(l2.bq_idx), l2.S, l2.ns, dp(l2.seg_off), dp(l2.row_seg), dp(l2.row_src),
                             dp(l2.row_w), st)
        return self


# ------------------------------------------------------------------------------------------------
# flat views of one encoder's parameters (+ derived weight layouts)
# ------------------------------------------------------------------------------------------------
class EncoderFlat:
    """Arena + per-layer views of an ``nets.make_encoder_params`` module tree."""

    def __init__(self, enc, device, grad_pool=None):
        sa_mods, fc = enc[0], enc[1]
        order, bufs = [], []
        self.bn_modules = []
        for i, sa in enumerate(sa_mods):
            seq = sa.mlps[0]
            for l in range(3):
                conv, bn = seq[3 * l], seq[3 * l + 1]
                order += [("sa%d.%d.W" % (i, l), conv.weight, False), ("sa%d.%d.gamma" % (i, l), bn.weight, False),
                          ("sa%d.%d.beta" % (i, l), bn.bias, False)]
                bufs.append(("sa%d.%d" % (i, l), bn))
        for j, (li, bi) in enumerate(((0, 1), (3, 4))):
            lin, bn = fc[li], fc[bi]
            order += [("fc%d.W" % j, lin.weight, False), ("fc%d.b" % j, lin.bias, False), ("fc%d.gamma" % j, bn.weight, False),
                      ("fc%d.beta" % j, bn.bias, False)]
            bufs.append(("fc%d" % j, bn))
        self.arena = nets.Arena(order, device, grad_pool=grad_pool)
        # the SA1 layers come first in the arena: their gradients are the LAST to be produced by a backward pass, so
        # [sa1_end, n) can be all-reduced while the SA1 backward still runs (agent._reduce_early / _reduce_late)
        self.sa1_end = self.arena.offsets["sa1.0.W"]
        # running statistics: float arena (no optimiser state) + int64 counters
        border = []
        for name, bn in bufs:
            border += [(name + ".rm", bn.running_mean, False), (name + ".rv", bn.running_var, False)]
        self.buffers = nets.Arena(border, device, with_opt=False)
        self.nbt = torch.zeros(len(bufs), dtype=torch.int64, device=device)
        for k, (name, bn) in enumerate(bufs):
            self.nbt[k] = int(bn.num_batches_tracked)
            bn.running_mean = self.buffers.view(name + ".rm", bn.running_mean.shape)
            bn.running_var = self.buffers.view(name + ".rv", bn.running_var.shape)
            bn.num_batches_tracked = self.nbt[k]
        A = self.arena
        self.layers = {}
        derived = []  # (key, N, K, rot, need_wp)
        k = 0
        for i, sa in enumerate(sa_mods):
            for l in range(3):
                conv = sa.mlps[0][3 * l]
                N, K = conv.weight.shape[0], conv.weight.shape[1]
                key = "sa%d.%d" % (i, l)
                self.layers[key] = self._layer(key, N, K, k, bias=False, rot=3 if (l == 0 and i > 0) else 0)
                k += 1
        for j in range(2):
            lin = fc[(0, 3)[j]]
            N, K = lin.weight.shape
            self.layers["fc%d" % j] = self._layer("fc%d" % j, N, K, k, bias=True, rot=0)
            k += 1
        self._build_derived(device)

    def _layer(self, key, N, K, k, bias, rot):
        A, Bf = self.arena, self.buffers
        L = NS(key=key, N=N, K=K, Kp=_pad4(K), rot=rot,
               W=A.view(key + ".W", (N, K)), dW=A.gview(key + ".W", (N, K)),
               bias=A.view(key + ".b", (N,)) if bias else None, dbias=A.gview(key + ".b", (N,)) if bias else None,
               gamma=A.view(key + ".gamma", (N,)), beta=A.view(key + ".beta", (N,)),
               dgamma=A.gview(key + ".gamma", (N,)), dbeta=A.gview(key + ".beta", (N,)),
               rm=Bf.view(key + ".rm", (N,)), rv=Bf.view(key + ".rv", (N,)), nbt=self.nbt[k:k + 1])
        return L

    def _build_derived(self, device):
        """Wf: forward B operand [N,Kp] (the parameter itself when K%4==0 and no rotation); WT: [Kp,N] for dX."""
        total, plan = 0, []
        for key, L in self.layers.items():
            need_wp = (L.K % 4 != 0) or L.rot != 0
            wp_off = total
            if need_wp:
                total += L.N * L.Kp
            wt_off = total
            total += L.Kp * L.N
            plan.append((L, need_wp, wp_off, wt_off))
        self.derived = _f(device, max(total, 4))
        jobs = []
        for L, need_wp, wp_off, wt_off in plan:
            L.Wf = self.derived[wp_off: wp_off + L.N * L.Kp].view(L.N, L.Kp) if need_wp else L.W
            L.WT = self.derived[wt_off: wt_off + L.Kp * L.N].view(L.Kp, L.N)
            jobs.append([L.W.data_ptr(), L.N, L.K, L.rot, L.Wf.data_ptr() if need_wp else 0, L.Kp, L.WT.data_ptr(), L.N])
        self.jobs = torch.tensor(jobs, dtype=torch.int64, device=device)
        # hi/lo TF32 split of the three SA1 conv weights, fetched by TMA in the fused SA1 chain (csrc/sa1_fused.cu)
        L0 = self.layers["sa0.0"]
        self.sa1f_ok = (L0.K <= 16 and (self.layers["sa0.0"].N, self.layers["sa0.1"].N, self.layers["sa0.2"].N) == (64, 64, 128)
                        and self.layers["sa0.1"].K == 64 and self.layers["sa0.2"].K == 64)
        self.sa1f_w = _f(device, int(lib.gaddpg_sa1f_wsplit_floats())) if self.sa1f_ok else None
        if self.sa1f_ok:   # split jobs (rot = -1) of the same batched launch: [W0 hi | W0 lo | W1 hi | W1 lo | W2 hi | W2 lo]
            w, off, L = self.sa1f_w.data_ptr(), 0, self.layers
            for key, kp in (("sa0.0", 16), ("sa0.1", 64), ("sa0.2", 64)):
                Lk = L[key]
                jobs.append([Lk.W.data_ptr(), Lk.N, Lk.K, -1, w + 4 * off, kp, w + 4 * (off + Lk.N * kp), kp])
                Lk.W_hi, Lk.W_lo = w + 4 * off, w + 4 * (off + Lk.N * kp)   # TMA sources of the forward row-GEMMs (K = Kp = 64)
                off += 2 * Lk.N * kp
            self.jobs = torch.tensor(jobs, dtype=torch.int64, device=device)
        self.refresh_derived()

    def refresh_derived(self):
        lib.gaddpg_wprep_batched(dp(self.jobs), self.jobs.shape[0], current_stream())


_refresh_cache = {}


def refresh_many(flats):
    """Derived weight layouts of several networks in ONE launch (their job tables concatenated once)."""
    key = tuple(id(f) for f in flats)
    j = _refresh_cache.get(key)
    if j is None:
        j = _refresh_cache[key] = torch.cat([f.jobs for f in flats]).contiguous()
    lib.gaddpg_wprep_batched(dp(j), j.shape[0], current_stream())


OPT_CHUNK = 4096
_OPT_DTYPE = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("target", "<u8"), ("tau_vec", "<u8"), ("dyn", "<u8"),
                       ("clip", "<u8"), ("absmax_p", "<u8"), ("absmax_g", "<u8"), ("n", "<i8"), ("chunk0", "<i8"), ("eps", "<f4"),
                       ("weight_decay", "<f4"), ("tau", "<f4"), ("kind", "<i4"), ("write_back", "<i4"), ("pad", "<i4")])
assert _OPT_DTYPE.itemsize == 120


class OptJobs:
    """Device job table of gaddpg_optim_multi: element ranges of the parameter arenas with what one optimiser phase does to
    them (Adam [+ Polyak target] [+ abs-max], Polyak only, statistics only) — the whole phase is one launch."""

    def __init__(self, device):
        self.rows, self.device, self.chunks, self.table = [], device, 0, None

    def _add(self, kind, n, **kw):
        if n <= 0:
            return
        r = np.zeros((), dtype=_OPT_DTYPE)
        r["kind"], r["n"], r["chunk0"] = kind, n, self.chunks
        for k, v in kw.items():
            r[k] = v
        self.rows.append(r)
        self.chunks += (n + OPT_CHUNK - 1) // OPT_CHUNK

    def adam(self, arena, off, n, dyn_ptr, eps, wd, clip=0, write_back=0, target=None, tau=0.0, absmax_p=0, absmax_g=0):
        q = lambda t: 0 if t is None else t.data_ptr() + 4 * off  # noqa: E731
        self._add(0, n, p=q(arena.p), g=q(arena.g), m=q(arena.m), v=q(arena.v), target=q(target), dyn=dyn_ptr, clip=clip, eps=eps,
                  weight_decay=wd, tau=tau, write_back=write_back, absmax_p=absmax_p, absmax_g=absmax_g)

    def polyak(self, source, target, off, n, tau=0.0, tau_vec=None, grads=None, absmax_p=0, absmax_g=0):
        q = lambda t: 0 if t is None else t.data_ptr() + 4 * off  # noqa: E731
        self._add(1, n, p=q(source), target=q(target), tau=tau, tau_vec=q(tau_vec), g=q(grads), absmax_p=absmax_p, absmax_g=absmax_g)

    def launch(self):
        if self.table is None:
            self.table = torch.from_numpy(np.stack(self.rows).view(np.uint8).reshape(len(self.rows), -1).copy()).to(self.device)
        lib.gaddpg_optim_multi(dp(self.table), len(self.rows), self.chunks, current_stream())


# ------------------------------------------------------------------------------------------------
# per-pass context (activations kept for backward) and backward scratch
# ------------------------------------------------------------------------------------------------
def _bnstate(device, C):
    return NS(scale=_f(device, C), shift=_f(device, C), mean=_f(device, C), rstd=_f(device, C))


def _bnbwd(device, C):
    return NS(g=_f(device, C), m1=_f(device, C), m2=_f(device, C))


class EncoderCtx:
    def __init__(self, B, geom_caps, widths, device):
        """geom_caps = (M1cap, M2cap); widths = [(64,64,128), (128,128,256), (256,256,512)]"""
        self.B = B
        caps = (geom_caps[0], geom_caps[1], B * 32)
        self.sa = []
        for i in range(3):
            w = widths[i]
            S = B * 32 if i < 2 else B
            self.sa.append(NS(Y=[_f(device, caps[i], c) for c in w], bn=[_bnstate(device, c) for c in w],
                              out=_f(device, S, w[2]), arg=torch.zeros(S, w[2], dtype=torch.int32, device=device),
                              G=None, cap=caps[i]))
        self.sa[1].G = _f(device, caps[1], _pad4(widths[0][2] + 3))
        self.sa[2].G = _f(device, caps[2], _pad4(widths[1][2] + 3))
        self.fc = NS(Y=[_f(device, B, 1024), _f(device, B, 512)], bn=[_bnstate(device, 1024), _bnstate(device, 512)])
        self.feat = _f(device, B, 516)
        self.bcbias = _f(device, B, 64)
        self.bc = None
        self.Cp = self.Cb = 0


class BwdScratch:
    def __init__(self, B, geom_caps, widths, device):
        caps = (geom_caps[0], geom_caps[1], B * 32)
        self.D = [[_f(device, caps[i], c) for c in widths[i]] for i in range(3)]
        self.bb = [[_bnbwd(device, c) for c in widths[i]] for i in range(3)]
        self.dG = [None, _f(device, caps[1], _pad4(widths[0][2] + 3)), _f(device, caps[2], _pad4(widths[1][2] + 3))]
        self.dout = [_f(device, B * 32, widths[0][2]), None, _f(device, B, 512)]  # SA1 pooled grad, -, SA3 pooled grad
        self.Dfc = [_f(device, B, 1024), _f(device, B, 512)]
        self.bbfc = [_bnbwd(device, 1024), _bnbwd(device, 512)]
        self.dbc = _f(device, B, 8)
        # sparse SA1 pool backward: pooled gradient after the pooled ReLU mask + one arg-max bit per (row, channel)
        self.E = _f(device, B * 32, widths[0][2])
        self.mask = torch.zeros(caps[0], widths[0][2] // 32, dtype=torch.int32, device=device)


WIDTHS = [(64, 64, 128), (128, 128, 256), (256, 256, 512)]


# ------------------------------------------------------------------------------------------------
# encoder forward / backward
# ------------------------------------------------------------------------------------------------
def encoder_forward(ws, ef, geom, cloud, skip, Cp, bc, ctx, time=None, time_offset=0.0, train=True, bn_stage=None, keep=True):
    """One pass of an encoder over ``cloud`` (B, C, skip+N).  Per-point input channels are cloud rows [0, Cp);
    ``bc`` (B, Cb) are per-sample constant channels appended after them (the action for the value encoder).
    Writes ctx.feat (B, 516) = [z(512) | time+offset | 0 0 0] and keeps the activations backward needs.
    ``bn_stage`` (BNStage): defer the running-statistics update of this pass (see BNStage).  ``keep=False``: the pass is
    never differentiated (target chain, select_action), so the last SA1 / SA2 layer outputs are not written at all —
    their max-pool is fused into the producing kernel."""
    B, C, Np = cloud.shape
    st = current_stream()
    l1, l2 = geom.lv
    Cb = 0 if bc is None else bc.shape[1]
    ctx.bc, ctx.Cp, ctx.Cb, ctx.cloud, ctx.skip, ctx.geom = bc, Cp, Cb, cloud, skip, geom
    L = ef.layers
    # ---- SA1: layer 0 straight from the cloud, layers 1-2 as row-GEMMs, pool over ball groups
    s = ctx.sa[0]
    W0 = L["sa0.0"]
    assert W0.K == 3 + Cp + Cb, (W0.K, Cp, Cb)
    fused = FUSED_SA1 or (FUSED_SA1_EVAL and not train and not keep)
    if fused and ef.sa1f_ok and (FUSED_SA1_KEEP or not keep) and lib.gaddpg_get_tensor_core() >= 3:
        _sa1_fused_forward(ws, ef, geom, cloud, skip, Cp, bc, Cb, ctx, train, bn_stage, keep)
    else:
        _sa1_unfused_forward(ws, ef, geom, cloud, skip, Cp, bc, Cb, ctx, train, bn_stage, keep)
    _encoder_forward_upper(ws, ef, geom, ctx, time, time_offset, train, bn_stage, keep)
    return ctx.feat


# SA1 shared MLP + max-pool as the TMA-fed recompute chain of csrc/sa1_fused.cu (three phases, no activation through HBM on
# forward-only passes).  Built, bit-for-rounding identical to the unfused kernels (tests/test_sa1_fused_gpu.py) and measured
# (profiles/r2_sa1_fused.md): at cfg2 the chain is bound by the per-tile dependency conv0 -> E -> conv1 -> E -> conv2 through
# ONE shared operand buffer (227 KB of shared memory hold the three hi/lo weight matrices + one 64 KB operand) and by the
# SIMT cost of the per-(row, channel) statistics / extreme pass — 329 us against 272 us for the unfused launches — so it is
# OFF by default; GADDPG_FUSED_SA1=1 (or engine.FUSED_SA1 = True) selects it.
FUSED_SA1 = _os.environ.get("GADDPG_FUSED_SA1", "0") == "1"
FUSED_SA1_KEEP = True   # ... also for passes that are differentiated (the pre-BN outputs are stored by TMA, never re-read)
# Eval-mode passes (select_action / select_action_batch / extract_feature: running statistics, so ONE phase instead of three)
# do use the chain by default: same latency as the unfused launches (scripts/bench_select_action.py: 228 vs 226 us of graph
# time per action at 4096 points) with two launches fewer and no activation written.
FUSED_SA1_EVAL = _os.environ.get("GADDPG_FUSED_SA1_EVAL", "1") == "1"


def _sa1_fused_forward(ws, ef, geom, cloud, skip, Cp, bc, Cb, ctx, train, bn_stage, keep):
    """SA1 through gaddpg_sa1_fused_fwd: phase k writes the batch statistics of conv k-1 (finalised in between), phase 3
    also the per-(ball, channel) extremes; forward-only passes never write an activation.  Eval mode (running statistics)
    needs phase 3 only."""
    B, C, Np = cloud.shape
    st = current_stream()
    l1 = geom.lv[0]
    s, L = ctx.sa[0], ef.layers
    lay = [L["sa0.0"], L["sa0.1"], L["sa0.2"]]
    count = B * geom.npoint * l1.ns
    b = ws.sa1f(l1.S, l1.cap)

    def phase(k):
        lib.gaddpg_sa1_fused_fwd(k, dp(cloud), C * Np, Np, skip, Cp, dp(bc), Cb, dp(l1.new_xyz), geom.npoint, dp(l1.seg_off),
                                 dp(l1.row_seg), dp(l1.row_src), dp(l1.row_w), l1.cap, l1.M_dev, dp(ef.sa1f_w),
                                 dp(s.bn[0].scale), dp(s.bn[0].shift), dp(s.bn[1].scale), dp(s.bn[1].shift), dp(lay[2].gamma),
                                 dp(ws.stats), dp(s.Y[k - 1]) if keep else None, dp(b.ext), dp(b.arg), dp(b.part_ext),
                                 dp(b.part_arg), dp(b.seg_part), st)

    if train:
        for k in (1, 2, 3):
            phase(k)
            bn_fwd(ws, lay[k - 1].N, count, lay[k - 1], s.bn[k - 1], True, bn_stage)
    else:
        for k in (1, 2, 3):
            bn_fwd(ws, lay[k - 1].N, count, lay[k - 1], s.bn[k - 1], False, None)
        phase(3)
    lib.gaddpg_sa1_pool_finalize(dp(b.ext), dp(b.arg), dp(b.part_ext), dp(b.part_arg), dp(b.seg_part), dp(lay[2].gamma),
                                 dp(s.bn[2].scale), dp(s.bn[2].shift), l1.S, dp(s.out), dp(s.arg), st)


def _sa1_unfused_forward(ws, ef, geom, cloud, skip, Cp, bc, Cb, ctx, train, bn_stage, keep):
    B, C, Np = cloud.shape
    st = current_stream()
    l1 = geom.lv[0]
    s, L = ctx.sa[0], ef.layers
    W0 = L["sa0.0"]
    t0 = fwd_tail(ws, 64, B * geom.npoint * l1.ns, W0, s.bn[0], train, bn_stage)
    lib.gaddpg_sa1_l1_fwd(dp(cloud), C * Np, Np, skip, Cp, dp(bc), Cb, B, dp(l1.new_xyz), geom.npoint, dp(l1.seg_off),
                          dp(l1.row_seg), dp(l1.row_src), dp(l1.row_w), l1.cap, l1.M_dev, dp(W0.W), W0.K, dp(ctx.bcbias),
                          dp(s.Y[0]), dp(ws.stats) if train else None, _tp(t0), st)
    bn_fwd_after(t0, ws, 64, B * geom.npoint * l1.ns, W0, s.bn[0], train, bn_stage)
    _mlp_tail_forward(ws, [L["sa0.1"], L["sa0.2"]], s, 1, l1.cap, l1.M_dev, l1.row_w, B * geom.npoint * l1.ns, train, bn_stage,
                      pool=NS(lv=l1, keep=keep))


def _encoder_forward_upper(ws, ef, geom, ctx, time, time_offset, train, bn_stage, keep):
    """SA2, SA3 and the FC head of encoder_forward."""
    B = ctx.B
    st = current_stream()
    l1, l2 = geom.lv
    L = ef.layers
    s = ctx.sa[0]
    # ---- SA2: gather [feats | dxyz | pad] rows, three row-GEMMs, pool
    s2 = ctx.sa[1]
    lib.gaddpg_gather_rows(dp(s.out), 128, dp(l1.new_xyz), geom.npoint, dp(l2.new_xyz), geom.npoint, dp(l2.row_seg),
                           dp(l2.row_src), l2.cap, l2.M_dev, dp(s2.G), s2.G.shape[1], st)
    _mlp_first_forward(ws, L["sa1.0"], s2, l2.cap, l2.M_dev, l2.row_w, B * geom.npoint * l2.ns, train, bn_stage)
    _mlp_tail_forward(ws, [L["sa1.1"], L["sa1.2"]], s2, 1, l2.cap, l2.M_dev, l2.row_w, B * geom.npoint * l2.ns, train, bn_stage,
                      pool=NS(lv=l2, keep=keep))
    # ---- SA3: GroupAll over the 32 SA2 centroids (absolute xyz)
    s3 = ctx.sa[2]
    M3 = B * geom.npoint
    lib.gaddpg_gather_rows(dp(s2.out), 256, dp(l2.new_xyz), geom.npoint, None, geom.npoint, None, None, M3, None, dp(s3.G),
                           s3.G.shape[1], st)
    _mlp_first_forward(ws, L["sa2.0"], s3, M3, None, None, M3, train, bn_stage)
    _mlp_tail_forward(ws, [L["sa2.1"], L["sa2.2"]], s3, 1, M3, None, None, M3, train, bn_stage)
    lib.gaddpg_pool_fwd(dp(s3.Y[2]), 512, dp(s3.bn[2].scale), dp(s3.bn[2].shift), None, geom.npoint, B, dp(s3.out), dp(s3.arg), st)
    # ---- FC head: Linear + BN1d + ReLU twice
    f = ctx.fc
    F0, F1 = L["fc0"], L["fc1"]
    t0 = fwd_tail(ws, 1024, B, F0, f.bn[0], train, bn_stage)
    nt([nt_problem(op_plain(s3.out), F0.Wf, F0.Kp, f.Y[0], 1024, B, None, 1024, 512, bias=F0.bias,
                   stats=ws.stats if train else None, tail=t0)], OP_PLAIN, EPI_STORE)
    bn_fwd_after(t0, ws, 1024, B, F0, f.bn[0], train, bn_stage)
    t1 = fwd_tail(ws, 512, B, F1, f.bn[1], train, bn_stage)
    nt([nt_problem(op_bnrelu(f.Y[0], f.bn[0]), F1.Wf, F1.Kp, f.Y[1], 512, B, None, 512, 1024, bias=F1.bias,
                   stats=ws.stats if train else None, tail=t1)], OP_BNRELU, EPI_STORE)
    bn_fwd_after(t1, ws, 512, B, F1, f.bn[1], train, bn_stage)
    lib.gaddpg_feat_finish(dp(f.Y[1]), 512, dp(f.bn[1].scale), dp(f.bn[1].shift), dp(time), float(time_offset), B, dp(ctx.feat),
                           516, st)
    return ctx.feat


def _mlp_first_forward(ws, Lp, s, M_max, M_dev, rw, count, train, bn_stage=None):
    t0 = fwd_tail(ws, Lp.N, count, Lp, s.bn[0], train, bn_stage)
    nt([nt_problem(op_plain(s.G), Lp.Wf, Lp.Kp, s.Y[0], Lp.N, M_max, M_dev, Lp.N, Lp.Kp, stats=ws.stats if train else None,
                   srw=rw, tail=t0)], OP_PLAIN, EPI_STORE)
    bn_fwd_after(t0, ws, Lp.N, count, Lp, s.bn[0], train, bn_stage)


FUSED_POOL = False  # max-pool of SA1 / SA2 inside the epilogue of their last layer: correct and bit-identical, but measured slower


def _mlp_tail_forward(ws, layers, s, first, M_max, M_dev, rw, count, train, bn_stage=None, pool=None):
    """Layers ``first``.. of one SA level.  ``pool`` (lv = geometry level, keep): the level's max-pool follows the last
    layer; on the tcgen05 kernels it is fused into that layer's epilogue (per-segment max / min keys of the raw output,
    turned into s.out / s.arg by gaddpg_pool_keys_finalize once the batch statistics exist)."""
    for j, Lp in enumerate(layers):
        l = first + j
        tl = fwd_tail(ws, Lp.N, count, Lp, s.bn[l], train, bn_stage)
        kw = dict(stats=ws.stats if train else None, srw=rw, tail=tl)
        A = op_bnrelu(s.Y[l - 1], s.bn[l - 1])
        fused = False
        if pool is not None and j == len(layers) - 1 and FUSED_POOL:
            lv = pool.lv
            prob = nt_problem(A, Lp.Wf, Lp.Kp, s.Y[l], Lp.N, M_max, M_dev, Lp.N, Lp.Kp, pool_keys=ws.keys(lv.S, Lp.N),
                              pool_seg=lv.row_seg, pool_gamma=Lp.gamma, no_store=0 if pool.keep else 1, **kw)
            g = NTGroup()
            g.p[0] = prob
            fused = lib.gaddpg_gemm_nt_path(ctypes.byref(g), 1, OP_BNRELU, EPI_STORE) in (1, 2)
        if not fused:
            ws_ = (Lp.W_hi, Lp.W_lo) if (getattr(Lp, "W_hi", None) and Lp.Wf is Lp.W) else None   # SA1: TMA-fetched weight images
            prob = nt_problem(A, Lp.Wf, Lp.Kp, s.Y[l], Lp.N, M_max, M_dev, Lp.N, Lp.Kp, w_split=ws_, **kw)
        nt([prob], OP_BNRELU, EPI_STORE)
        bn_fwd_after(tl, ws, Lp.N, count, Lp, s.bn[l], train, bn_stage)
        if pool is not None and j == len(layers) - 1:
            lv, st = pool.lv, current_stream()
            if fused:
                lib.gaddpg_pool_keys_finalize(dp(ws.keys(lv.S, Lp.N)), lv.S, Lp.N, dp(Lp.gamma), dp(s.bn[l].scale), dp(s.bn[l].shift), dp(s.out),
                                              dp(s.arg), st)
            else:
                lib.gaddpg_pool_fwd(dp(s.Y[l]), Lp.N, dp(s.bn[l].scale), dp(s.bn[l].shift), dp(lv.seg_off), 0, lv.S, dp(s.out),
                                    dp(s.arg), st)


def entry_tail(ws, ef, ctx, sc, want_dw=True, accumulate=0):
    """BatchNorm-backward tail of the encoder's last FC BatchNorm, for the kernel that writes its (sum D, sum D*xhat) statistics:
    the EPI_DMASK epilogue at the end of policy_backward / critic_backward.  Pass it there as ``enc_tail`` and call
    encoder_backward(..., entry_done=(tail is not None)) with the same want_dw / accumulate."""
    return bwd_tail(ws, 512, ctx.B, ef.layers["fc1"], ctx.fc.bn[1], sc.bbfc[1], want_dw, accumulate)


def encoder_backward(ws, ef, ctx, sc, want_dw=True, want_dbc=False, accumulate=0, dfeat=None, part="all", entry_done=False):
    """Backward of ``encoder_forward``.  Entry: either ``dfeat`` (B, >=512) is given (gradient w.r.t. ctx.feat;
    masked here), or the caller already produced sc.Dfc[1] and its BN sums in ws.stats through an EPI_DMASK
    epilogue (the fused path).  Writes parameter gradients into the arena (``want_dw``) and, for the broadcast
    channels, sc.dbc (B, Cb) (``want_dbc``).  ``part``: "upper" stops after SA2 (FC head, SA3, SA2 and the scatter into
    SA1's pooled gradient: every parameter gradient except SA1's is complete), "sa1" runs the rest — the sharded agent
    all-reduces the upper gradients while the SA1 backward runs."""
    B = ctx.B
    st = current_stream()
    geom = ctx.geom
    l1, l2 = geom.lv
    L = ef.layers
    if part == "sa1":
        return _encoder_backward_sa1(ws, ef, ctx, sc, want_dw, want_dbc, accumulate)
    f = ctx.fc
    F0, F1 = L["fc0"], L["fc1"]
    if dfeat is not None:
        te = bwd_tail(ws, 512, B, F1, f.bn[1], sc.bbfc[1], want_dw, accumulate)
        lib.gaddpg_dmask_stats(dp(dfeat), dfeat.shape[1], dp(f.Y[1]), 512, B, dp(f.bn[1].scale), dp(f.bn[1].shift),
                               dp(f.bn[1].mean), dp(f.bn[1].rstd), dp(sc.Dfc[1]), dp(ws.stats), _tp(te), st)
        entry_done = te is not None
    s3 = ctx.sa[2]
    # ---- FC head (entry_done: the kernel that produced sc.Dfc[1] also finalized this BatchNorm's backward sums)
    if not entry_done:
        bn_bwd(ws, 512, B, F1, f.bn[1], sc.bbfc[1], want_dw, accumulate)
    dy1 = op_bnbwd(sc.Dfc[1], f.Y[1], f.bn[1], sc.bbfc[1])
    if want_dw:
        tn(ws, dy1, op_bnrelu(f.Y[0], f.bn[0]), OP_BNBWD, OP_BNRELU, B, None, 512, 1024, F1.dW, 1024, 512, 1024,
           dbias=F1.dbias, accumulate=accumulate)
    t0 = bwd_tail(ws, 1024, B, F0, f.bn[0], sc.bbfc[0], want_dw, accumulate)
    nt([nt_problem(dy1, F1.WT, 512, sc.Dfc[0], 1024, B, None, 1024, 512, stats=ws.stats, Yprev=f.Y[0], ldyp=1024,
                   pbn=f.bn[0], tail=t0)], OP_BNBWD, EPI_DMASK)
    if t0 is None:
        bn_bwd(ws, 1024, B, F0, f.bn[0], sc.bbfc[0], want_dw, accumulate)
    dy0 = op_bnbwd(sc.Dfc[0], f.Y[0], f.bn[0], sc.bbfc[0])
    if want_dw:
        tn(ws, dy0, op_plain(s3.out), OP_BNBWD, OP_PLAIN, B, None, 1024, 512, F0.dW, 512, 1024, 512, dbias=F0.dbias,
           accumulate=accumulate)
    nt([nt_problem(dy0, F0.WT, 1024, sc.dout[2], 512, B, None, 512, 1024)], OP_BNBWD, EPI_STORE)
    # ---- SA3
    M3 = B * geom.npoint
    _sa_backward(ws, [L["sa2.0"], L["sa2.1"], L["sa2.2"]], s3, sc, 2, sc.dout[2], 512, None, geom.npoint, M3, None, None, M3,
                 want_dw, accumulate, True)
    # dG3[:, :256] is the gradient of SA2's pooled output (identity rows)
    s2 = ctx.sa[1]
    _sa_backward(ws, [L["sa1.0"], L["sa1.1"], L["sa1.2"]], s2, sc, 1, sc.dG[2], sc.dG[2].shape[1], l2.row_seg, 0, l2.cap,
                 l2.M_dev, l2.row_w, B * geom.npoint * l2.ns, want_dw, accumulate, True)
    lib.gaddpg_scatter_rows(dp(sc.dG[1]), sc.dG[1].shape[1], 128, B, geom.npoint, geom.npoint, dp(l2.seg_off), dp(l2.row_src),
                            dp(sc.dout[0]), st)
    if part == "upper":
        return None
    return _encoder_backward_sa1(ws, ef, ctx, sc, want_dw, want_dbc, accumulate)


def _encoder_backward_sa1(ws, ef, ctx, sc, want_dw, want_dbc, accumulate):
    B = ctx.B
    st = current_stream()
    geom = ctx.geom
    l1, l2 = geom.lv
    L = ef.layers
    # ---- SA1: layers 2,1 generic; layer 0 custom
    s = ctx.sa[0]
    W0 = L["sa0.0"]
    count = B * geom.npoint * l1.ns
    _sa_backward(ws, [W0, L["sa0.1"], L["sa0.2"]], s, sc, 0, sc.dout[0], 128, l1.row_seg, 0, l1.cap, l1.M_dev, l1.row_w,
                 count, want_dw, accumulate, False)
    bb0 = sc.bb[0][0]
    lib.gaddpg_sa1_l1_bwd(dp(ctx.cloud), ctx.cloud.shape[1] * ctx.cloud.shape[2], ctx.cloud.shape[2], ctx.skip, ctx.Cp,
                          dp(ctx.bc), ctx.Cb, B, dp(l1.new_xyz), geom.npoint, dp(l1.seg_off), dp(l1.row_seg), dp(l1.row_src),
                          dp(l1.row_w), l1.cap, l1.M_dev, dp(sc.D[0][0]), dp(s.Y[0]), dp(bb0.g), dp(bb0.m1), dp(bb0.m2),
                          dp(s.bn[0].mean), dp(s.bn[0].rstd), dp(W0.W), W0.K, dp(W0.dW) if want_dw else None, accumulate,
                          dp(sc.dbc) if (want_dbc and ctx.Cb > 0) else None, dp(ws.sa1(B, l1.cap)), ws.sa1_ws_bytes, st)
    return sc.dbc.view(-1)[: B * ctx.Cb].view(B, ctx.Cb) if (want_dbc and ctx.Cb > 0) else None


def _sa_backward(ws, layers, s, sc, lvl, dOut, ld_dout, row_seg, fixed_len, M_max, M_dev, rw, count, want_dw, accumulate,
                 generic_l0):
    """Pool backward + the BN/conv layers of one SA level.  Always leaves D[lvl][0] and bb[lvl][0] (the BN-backward
    constants of layer 0) ready; with ``generic_l0`` it also runs layer 0's conv backward (dW, dG[lvl]) — SA1's
    first layer has its own kernel (gaddpg_sa1_l1_bwd)."""
    st = current_stream()
    D, bb = sc.D[lvl], sc.bb[lvl]
    C3 = layers[2].N
    # SA1 (423 k rows x 128 channels at cfg2): the dense pool gradient D[2] has one non-zero per (segment, channel), so it
    # is never written; its consumers (dX and dW of layer 2) rebuild it from the (S, C) tables (GADDPG_OP_BNBWD_POOL)
    sparse = (SPARSE_POOL and lvl == 0 and row_seg is not None and C3 == 128 and layers[2].Kp == 64
              and lib.gaddpg_get_tensor_core() >= 3)
    # every kernel that writes BatchNorm-backward statistics finalizes them in its tail (bwd_tail); tl = tail of layer l's sums
    tl = bwd_tail(ws, C3, count, layers[2], s.bn[2], bb[2], want_dw, accumulate)
    if sparse:
        lib.gaddpg_pool_bwd_sparse(dp(dOut), ld_dout, dp(s.out), dp(s.arg), dp(s.Y[2]), C3, s.out.shape[0], dp(s.bn[2].mean),
                                   dp(s.bn[2].rstd), dp(sc.E), dp(sc.mask), M_max, dp(ws.stats), _tp(tl), st)
    else:
        lib.gaddpg_pool_bwd(dp(dOut), ld_dout, dp(s.out), dp(s.arg), dp(s.Y[2]), C3, dp(row_seg), fixed_len, M_max, M_dev,
                            dp(s.bn[2].mean), dp(s.bn[2].rstd), dp(D[2]), dp(ws.stats), _tp(tl), st)
    for l in (2, 1):
        Lp = layers[l]
        if tl is None:
            bn_bwd(ws, Lp.N, count, Lp, s.bn[l], bb[l], want_dw, accumulate)
        if sparse and l == 2:
            dy, dmode = op_bnbwd_pool(sc.E, sc.mask, row_seg, s.Y[l], s.bn[l], bb[l], rw=rw), OP_BNBWD_POOL
        else:
            dy, dmode = op_bnbwd(D[l], s.Y[l], s.bn[l], bb[l], rw=rw), OP_BNBWD
        if want_dw:
            tn(ws, dy, op_bnrelu(s.Y[l - 1], s.bn[l - 1]), dmode, OP_BNRELU, M_max, M_dev, Lp.N, Lp.Kp, Lp.dW, Lp.K, Lp.N,
               Lp.K, accumulate=accumulate)
        tl = bwd_tail(ws, layers[l - 1].N, count, layers[l - 1], s.bn[l - 1], bb[l - 1], want_dw, accumulate)
        nt([nt_problem(dy, Lp.WT, Lp.N, D[l - 1], Lp.Kp, M_max, M_dev, Lp.Kp, Lp.N, stats=ws.stats, Yprev=s.Y[l - 1],
                       ldyp=Lp.Kp, pbn=s.bn[l - 1], tail=tl)], dmode, EPI_DMASK)
    L0 = layers[0]
    if tl is None:
        bn_bwd(ws, L0.N, count, L0, s.bn[0], bb[0], want_dw, accumulate)
    if generic_l0:
        dy = op_bnbwd(D[0], s.Y[0], s.bn[0], bb[0], rw=rw)
        if want_dw:
            tn(ws, dy, op_plain(s.G), OP_BNBWD, OP_PLAIN, M_max, M_dev, L0.N, L0.Kp, L0.dW, L0.K, L0.N, L0.K, rot=L0.rot,
               accumulate=accumulate)
        nt([nt_problem(dy, L0.WT, L0.N, sc.dG[lvl], L0.Kp, M_max, M_dev, L0.Kp, L0.N)], OP_BNBWD, EPI_STORE)


# ------------------------------------------------------------------------------------------------
# actor / critic heads
# ------------------------------------------------------------------------------------------------
H = 256          # hidden_size (experiments/config.py:74)
FEAT_LD = 516    # 512 features + time column, padded to a multiple of 4
QA_LD, QA_Q2, QA_AUX = 16, 4, 8   # critic output row: q1 @0, q2 @4, aux(7) @8 (16-byte aligned sub-blocks)


def _jobs_tensor(jobs, device):
    return torch.tensor(jobs, dtype=torch.int64, device=device)


class PolicyFlat:
    """GaussianPolicy (networks.py:303-351): 513 -> 256 -> 256 -> [mean(6) | extra_pred(E) | log_std(6)]."""

    def __init__(self, mod, device, with_opt=True, grad_pool=None):
        self.mod, self.E = mod, mod.extra_pred_dim
        E = self.E
        order = [("l1.W", mod.linear1.weight, False), ("l1.b", mod.linear1.bias, False),
                 ("l2.W", mod.linear2.weight, False), ("l2.b", mod.linear2.bias, False),
                 ("h.W", mod.mean.weight, False), ("h.W.e", mod.extra_pred.weight, True), ("h.W.s", mod.log_std_linear.weight, True),
                 ("h.b", mod.mean.bias, False), ("h.b.e", mod.extra_pred.bias, True), ("h.b.s", mod.log_std_linear.bias, True)]
        self.arena = A = nets.Arena(order, device, with_opt=with_opt, grad_pool=grad_pool)
        self.NH = 6 + E + 6
        self.NHp = _pad4(self.NH)
        self.Kin = mod.linear1.weight.shape[1]
        o = A.offsets

        def v(base, name, n, shape):
            return None if base is None else base[o[name]: o[name] + n].view(shape)

        self.W1, self.b1 = v(A.p, "l1.W", H * self.Kin, (H, self.Kin)), v(A.p, "l1.b", H, (H,))
        self.W2, self.b2 = v(A.p, "l2.W", H * H, (H, H)), v(A.p, "l2.b", H, (H,))
        self.Wh, self.bh = v(A.p, "h.W", self.NH * H, (self.NH, H)), v(A.p, "h.b", self.NH, (self.NH,))
        self.dW1, self.db1 = v(A.g, "l1.W", H * self.Kin, (H, self.Kin)), v(A.g, "l1.b", H, (H,))
        self.dW2, self.db2 = v(A.g, "l2.W", H * H, (H, H)), v(A.g, "l2.b", H, (H,))
        self.dWh, self.dbh = v(A.g, "h.W", self.NH * H, (self.NH, H)), v(A.g, "h.b", self.NH, (self.NH,))
        Kp = _pad4(self.Kin)
        self.derived = _f(device, H * Kp + Kp * H + H * H + H * self.NHp)
        d, off = self.derived, 0
        self.W1f = d[off: off + H * Kp].view(H, Kp); off += H * Kp
        self.W1T = d[off: off + Kp * H].view(Kp, H); off += Kp * H
        self.W2T = d[off: off + H * H].view(H, H); off += H * H
        self.WhT = d[off: off + H * self.NHp].view(H, self.NHp)
        self.jobs = _jobs_tensor([[self.W1.data_ptr(), H, self.Kin, 0, self.W1f.data_ptr(), Kp, self.W1T.data_ptr(), H],
                                  [self.W2.data_ptr(), H, H, 0, 0, H, self.W2T.data_ptr(), H],
                                  [self.Wh.data_ptr(), self.NH, H, 0, 0, H, self.WhT.data_ptr(), self.NHp]], device)
        self.refresh_derived()

    def refresh_derived(self):
        lib.gaddpg_wprep_batched(dp(self.jobs), self.jobs.shape[0], current_stream())

    def adam_ranges(self, use_aux):
        """(offset, count) ranges that receive gradients in update(): everything except log_std_linear, and
        extra_pred when policy_aux is off (their .grad stays None in the reference, so Adam skips them)."""
        o, s = self.arena.offsets, self.arena.sizes
        wend = (o["h.W.e"] + s["h.W.e"]) if use_aux else (o["h.W"] + s["h.W"])
        nb = 6 + (self.E if use_aux else 0)
        return [(0, wend), (o["h.b"], nb)], nb


def policy_ctx(B, pf, device):
    return NS(H1=_f(device, B, H), H2=_f(device, B, H), raw=_f(device, B, pf.NHp), pi=_f(device, B, 6),
              draw=_f(device, B, pf.NHp), dZ2=_f(device, B, H), dZ1=_f(device, B, H))


def policy_forward(pf, feat, pc, B):
    nt([nt_problem(op_plain(feat, FEAT_LD), pf.W1f, FEAT_LD, pc.H1, H, B, None, H, FEAT_LD, bias=pf.b1, relu=1)], OP_PLAIN, EPI_STORE)
    nt([nt_problem(op_plain(pc.H1), pf.W2, H, pc.H2, H, B, None, H, H, bias=pf.b2, relu=1)], OP_PLAIN, EPI_STORE)
    nt([nt_problem(op_plain(pc.H2), pf.Wh, H, pc.raw, pf.NHp, B, None, pf.NH, H, bias=pf.bh)], OP_PLAIN, EPI_STORE)
    return pc.raw


def policy_backward(ws, pf, feat, pc, B, n_grad, enc_ctx, sc, accumulate=0, enc_tail=None):
    """pc.draw (B, NHp): gradient w.r.t. the raw head output (columns >= n_grad are zero).  Ends in the encoder's
    FC head through the fused mask epilogue: sc.Dfc[1] + BN sums in ws.stats (entry state of encoder_backward);
    ``enc_tail`` (entry_tail()): that kernel also finalizes the sums."""
    tn(ws, op_plain(pc.draw), op_plain(pc.H2), OP_PLAIN, OP_PLAIN, B, None, pf.NHp, H, pf.dWh, H, n_grad, H, dbias=pf.dbh,
       accumulate=accumulate)
    nt([nt_problem(op_plain(pc.draw), pf.WhT, pf.NHp, pc.dZ2, H, B, None, H, pf.NHp, Yprev=pc.H2, ldyp=H)], OP_PLAIN, EPI_DMASK)
    tn(ws, op_plain(pc.dZ2), op_plain(pc.H1), OP_PLAIN, OP_PLAIN, B, None, H, H, pf.dW2, H, H, H, dbias=pf.db2, accumulate=accumulate)
    nt([nt_problem(op_plain(pc.dZ2), pf.W2T, H, pc.dZ1, H, B, None, H, H, Yprev=pc.H1, ldyp=H)], OP_PLAIN, EPI_DMASK)
    tn(ws, op_plain(pc.dZ1), op_plain(feat, FEAT_LD), OP_PLAIN, OP_PLAIN, B, None, H, FEAT_LD, pf.dW1, pf.Kin, H, pf.Kin,
       dbias=pf.db1, accumulate=accumulate)
    f = enc_ctx.fc
    nt([nt_problem(op_plain(pc.dZ1), pf.W1T, H, sc.Dfc[1], 512, B, None, 512, H, stats=ws.stats, Yprev=f.Y[1], ldyp=512,
                   pbn=f.bn[1], tail=enc_tail)], OP_PLAIN, EPI_DMASK)


class CriticFlat:
    """QNetwork (networks.py:253-300, num_actions=0): twin Q (linear1-3, linear4-6) + aux branch (linear7, 8,
    extra_pred) when extra_pred_dim > 0.  First layers are stacked into one (256*nb, 513) matrix."""

    def __init__(self, mod, device, with_opt=True, grad_pool=None):
        self.mod, self.E = mod, mod.extra_pred_dim
        self.nb = 3 if self.E > 0 else 2
        l1 = [mod.linear1, mod.linear4] + ([mod.linear7] if self.E else [])
        l2 = [mod.linear2, mod.linear5] + ([mod.linear8] if self.E else [])
        l3 = [mod.linear3, mod.linear6] + ([mod.extra_pred] if self.E else [])
        self.nout = [1, 1] + ([self.E] if self.E else [])
        self.cols = [0, QA_Q2, QA_AUX][: self.nb]
        order = []
        for i, m in enumerate(l1):
            order.append(("l1.W.%d" % i, m.weight, i > 0))
        for i, m in enumerate(l1):
            order.append(("l1.b.%d" % i, m.bias, i > 0))
        for i, m in enumerate(l2):
            order.append(("l2.W.%d" % i, m.weight, False))
            order.append(("l2.b.%d" % i, m.bias, False))
        for i, m in enumerate(l3):
            order.append(("l3.W.%d" % i, m.weight, False))
            order.append(("l3.b.%d" % i, m.bias, False))
        self.arena = A = nets.Arena(order, device, with_opt=with_opt, grad_pool=grad_pool)
        self.Kin = mod.linear1.weight.shape[1]
        o, nb = A.offsets, self.nb
        Kp = _pad4(self.Kin)

        def v(base, name, n, shape):
            return None if base is None else base[o[name]: o[name] + n].view(shape)

        self.W1, self.b1 = v(A.p, "l1.W.0", nb * H * self.Kin, (nb * H, self.Kin)), v(A.p, "l1.b.0", nb * H, (nb * H,))
        self.dW1, self.db1 = v(A.g, "l1.W.0", nb * H * self.Kin, (nb * H, self.Kin)), v(A.g, "l1.b.0", nb * H, (nb * H,))
        self.W2 = [v(A.p, "l2.W.%d" % i, H * H, (H, H)) for i in range(nb)]
        self.b2 = [v(A.p, "l2.b.%d" % i, H, (H,)) for i in range(nb)]
        self.dW2 = [v(A.g, "l2.W.%d" % i, H * H, (H, H)) for i in range(nb)]
        self.db2 = [v(A.g, "l2.b.%d" % i, H, (H,)) for i in range(nb)]
        self.W3 = [v(A.p, "l3.W.%d" % i, self.nout[i] * H, (self.nout[i], H)) for i in range(nb)]
        self.b3 = [v(A.p, "l3.b.%d" % i, self.nout[i], (self.nout[i],)) for i in range(nb)]
        self.dW3 = [v(A.g, "l3.W.%d" % i, self.nout[i] * H, (self.nout[i], H)) for i in range(nb)]
        self.db3 = [v(A.g, "l3.b.%d" % i, self.nout[i], (self.nout[i],)) for i in range(nb)]
        self.npad = [_pad4(n) for n in self.nout]
        size = nb * H * Kp + Kp * nb * H + nb * H * H + sum(H * p for p in self.npad)
        self.derived = d = _f(device, size)
        off = 0
        self.W1f = d[off: off + nb * H * Kp].view(nb * H, Kp); off += nb * H * Kp
        self.W1T = d[off: off + Kp * nb * H].view(Kp, nb * H); off += Kp * nb * H
        self.W2T, self.W3T = [], []
        for i in range(nb):
            self.W2T.append(d[off: off + H * H].view(H, H)); off += H * H
        for i in range(nb):
            self.W3T.append(d[off: off + H * self.npad[i]].view(H, self.npad[i])); off += H * self.npad[i]
        jobs = [[self.W1.data_ptr(), nb * H, self.Kin, 0, self.W1f.data_ptr(), Kp, self.W1T.data_ptr(), nb * H]]
        for i in range(nb):
            jobs.append([self.W2[i].data_ptr(), H, H, 0, 0, H, self.W2T[i].data_ptr(), H])
            jobs.append([self.W3[i].data_ptr(), self.nout[i], H, 0, 0, H, self.W3T[i].data_ptr(), self.npad[i]])
        self.jobs = _jobs_tensor(jobs, device)
        self.refresh_derived()

    def refresh_derived(self):
        lib.gaddpg_wprep_batched(dp(self.jobs), self.jobs.shape[0], current_stream())

    def tau_vectors(self, tau):
        """half_soft_update / half_hard_update masks (utils.py:757-770): Polyak on linear1-3 only; hard copy of
        linear4-6 every target_update_interval; the aux branch of the target is never touched."""
        A = self.arena
        soft = torch.zeros(A.n, dtype=torch.float32, device=A.p.device)
        hard = torch.zeros(A.n, dtype=torch.float32, device=A.p.device)
        for name in ("l1.W.0", "l1.b.0", "l2.W.0", "l2.b.0", "l3.W.0", "l3.b.0"):
            soft[A.offsets[name]: A.offsets[name] + A.sizes[name]] = tau
        for name in ("l1.W.1", "l1.b.1", "l2.W.1", "l2.b.1", "l3.W.1", "l3.b.1"):
            hard[A.offsets[name]: A.offsets[name] + A.sizes[name]] = 1.0
        return soft, hard


def critic_ctx(B, cf, device):
    n = cf.nb * H
    return NS(H1=_f(device, B, n), H2=_f(device, B, n), qa=_f(device, B, QA_LD), dqa=_f(device, B, QA_LD),
              dZ2=_f(device, B, n), dZ1=_f(device, B, n))


def _off(t, col):
    return t.data_ptr() + 4 * col


def critic_forward(cf, feat, cc, B, nb=None):
    nb = cf.nb if nb is None else nb
    n = cf.nb * H
    nt([nt_problem(op_plain(feat, FEAT_LD), cf.W1f, FEAT_LD, cc.H1, n, B, None, nb * H, FEAT_LD, bias=cf.b1, relu=1)],
       OP_PLAIN, EPI_STORE)
    nt([nt_problem(Operand(X=_off(cc.H1, H * i), ldx=n), cf.W2[i], H, _off(cc.H2, H * i), n, B, None, H, H, bias=cf.b2[i], relu=1)
        for i in range(nb)], OP_PLAIN, EPI_STORE)
    nt([nt_problem(Operand(X=_off(cc.H2, H * i), ldx=n), cf.W3[i], H, _off(cc.qa, cf.cols[i]), QA_LD, B, None, cf.nout[i], H,
                   bias=cf.b3[i]) for i in range(nb)], OP_PLAIN, EPI_STORE)
    return cc.qa


def critic_backward(ws, cf, feat, cc, B, nb, enc_ctx, sc, accumulate=0, enc_tail=None):
    """cc.dqa (B,16): gradient w.r.t. [q1 | q2 | aux_raw]; ``nb`` = branches that carry gradient (2 in the actor
    step, where the aux branch is not part of the loss).  Ends like policy_backward in sc.Dfc[1] + ws.stats."""
    n = cf.nb * H
    for i in range(nb):
        tn(ws, Operand(X=_off(cc.dqa, cf.cols[i]), ldx=QA_LD), Operand(X=_off(cc.H2, H * i), ldx=n), OP_PLAIN, OP_PLAIN, B, None,
           cf.npad[i], H, cf.dW3[i], H, cf.nout[i], H, dbias=cf.db3[i], accumulate=accumulate)
    nt([nt_problem(Operand(X=_off(cc.dqa, cf.cols[i]), ldx=QA_LD), cf.W3T[i], cf.npad[i], _off(cc.dZ2, H * i), n, B, None, H,
                   cf.npad[i], Yprev=_off(cc.H2, H * i), ldyp=n) for i in range(nb)], OP_PLAIN, EPI_DMASK)
    for i in range(nb):
        tn(ws, Operand(X=_off(cc.dZ2, H * i), ldx=n), Operand(X=_off(cc.H1, H * i), ldx=n), OP_PLAIN, OP_PLAIN, B, None, H, H,
           cf.dW2[i], H, H, H, dbias=cf.db2[i], accumulate=accumulate)
    nt([nt_problem(Operand(X=_off(cc.dZ2, H * i), ldx=n), cf.W2T[i], H, _off(cc.dZ1, H * i), n, B, None, H, H,
                   Yprev=_off(cc.H1, H * i), ldyp=n) for i in range(nb)], OP_PLAIN, EPI_DMASK)
    tn(ws, op_plain(cc.dZ1, n), op_plain(feat, FEAT_LD), OP_PLAIN, OP_PLAIN, B, None, nb * H, FEAT_LD, cf.dW1, cf.Kin, nb * H,
       cf.Kin, dbias=cf.db1, accumulate=accumulate)
    f = enc_ctx.fc
    nt([nt_problem(op_plain(cc.dZ1, n), cf.W1T, n, sc.Dfc[1], 512, B, None, 512, nb * H, stats=ws.stats, Yprev=f.Y[1], ldyp=512,
                   pbn=f.bn[1], tail=enc_tail)], OP_PLAIN, EPI_DMASK)
