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
from types import SimpleNamespace as NS

import numpy as np
import torch

from . import nets
from .capi import current_stream, lib
from .structs import (EPI_DMASK, EPI_STORE, OP_BNBWD, OP_BNRELU, OP_PLAIN, STAT_SLOTS, NTGroup, NTProblem, Operand,
                      TNProblem, check_sizes, dp, op_bnbwd, op_bnrelu, op_plain)

BN_EPS, BN_MOMENTUM = 1e-5, 0.1
_checked = False


def _init_once():
    global _checked
    if not _checked:
        check_sizes()
        scale = ((nets.ACTION_HIGH - nets.ACTION_LOW) / 2.0).astype(np.float32)
        bias = ((nets.ACTION_HIGH + nets.ACTION_LOW) / 2.0).astype(np.float32)
        cp = np.array([[0, 0, 0], [0, 0, 0], [0.053, -0.0, 0.075], [-0.053, 0, 0.075], [0.053, -0.0, 0.105],
                       [-0.053, 0, 0.105]], dtype=np.float32)
        a = np.pi / 2  # utils.py:826-829: float64 matmul with rotZ(pi/2)[:3,:3], then cast to float32
        rz = np.array([[np.cos(a), -np.sin(a), 0.0], [np.sin(a), np.cos(a), 0.0], [0.0, 0.0, 1.0]])
        cpz = np.ascontiguousarray(np.matmul(cp, rz).astype(np.float32))
        lib.gaddpg_heads_init(scale.ctypes.data, bias.ctypes.data, cpz.ctypes.data)
        _checked = True


def _f(device, *shape):
    return torch.zeros(*shape, dtype=torch.float32, device=device)


def _pad4(k):
    return (k + 3) // 4 * 4


class Workspace:
    """Scratch shared by all passes on one device (stream-ordered reuse)."""

    def __init__(self, device):
        _init_once()
        self.device = device
        self.stats = _f(device, STAT_SLOTS * 2 * 1024)
        self.tn_bytes = int(lib.gaddpg_gemm_tn_workspace_bytes())
        self.tn = _f(device, self.tn_bytes // 4)
        self.red = _f(device, 2048)
        self.sa1_ws_bytes = 0
        self.sa1_ws = None

    def sa1(self, B):
        need = (STAT_SLOTS * 64 * 16 + B * 64) * 4
        if self.sa1_ws is None or self.sa1_ws_bytes < need:
            self.sa1_ws = _f(self.device, need // 4)
            self.sa1_ws_bytes = need
        return self.sa1_ws


# ------------------------------------------------------------------------------------------------
# thin launch helpers
# ------------------------------------------------------------------------------------------------
def nt(problems, amode, emode):
    g = NTGroup()
    for i, p in enumerate(problems):
        g.p[i] = p
    lib.gaddpg_gemm_nt(ctypes.byref(g), len(problems), amode, emode, current_stream())


def nt_problem(A, Bw, ldb, C, ldc, M_max, M_dev, N, K, bias=None, relu=0, stats=None, srw=None, Yprev=None, ldyp=0,
               pbn=None):
    p = NTProblem(A=A, Bw=dp(Bw) if torch.is_tensor(Bw) else Bw, ldb=ldb, bias=dp(bias), C=dp(C) if torch.is_tensor(C) else C,
                  ldc=ldc, M_max=M_max, M_dev=M_dev, N=N, K=K, relu=relu, stats=dp(stats), srw=dp(srw),
                  Yprev=dp(Yprev) if torch.is_tensor(Yprev) else Yprev, ldyp=ldyp)
    if pbn is not None:
        p.psc, p.psh, p.pmean, p.prstd = dp(pbn.scale), dp(pbn.shift), dp(pbn.mean), dp(pbn.rstd)
    return p


def tn(ws, P, Q, pmode, qmode, M_max, M_dev, N, K, dW, ldd, Ntrue, Ktrue, rot=0, dbias=None, accumulate=0):
    prob = TNProblem(P=P, Q=Q, M_max=M_max, M_dev=M_dev, N=N, K=K)
    lib.gaddpg_gemm_tn(ctypes.byref(prob), pmode, qmode, dp(dW), ldd, Ntrue, Ktrue, rot, dp(dbias), accumulate,
                       dp(ws.tn), ws.tn_bytes, current_stream())


def bn_fwd(ws, C, count, bnp, st, train):
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

    def __init__(self, enc, device):
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
        self.arena = nets.Arena(order, device)
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
        self.refresh_derived()

    def refresh_derived(self):
        lib.gaddpg_wprep_batched(dp(self.jobs), self.jobs.shape[0], current_stream())


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
        self.dY1 = _f(device, caps[0], 64)
        self.dbc = _f(device, B, 8)


WIDTHS = [(64, 64, 128), (128, 128, 256), (256, 256, 512)]


# ------------------------------------------------------------------------------------------------
# encoder forward / backward
# ------------------------------------------------------------------------------------------------
def encoder_forward(ws, ef, geom, cloud, skip, Cp, bc, ctx, time=None, time_offset=0.0, train=True):
    """One pass of an encoder over ``cloud`` (B, C, skip+N).  Per-point input channels are cloud rows [0, Cp);
    ``bc`` (B, Cb) are per-sample constant channels appended after them (the action for the value encoder).
    Writes ctx.feat (B, 516) = [z(512) | time+offset | 0 0 0] and keeps the activations backward needs."""
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
    lib.gaddpg_sa1_l1_fwd(dp(cloud), C * Np, Np, skip, Cp, dp(bc), Cb, B, dp(l1.new_xyz), geom.npoint, dp(l1.row_seg),
                          dp(l1.row_src), dp(l1.row_w), l1.cap, l1.M_dev, dp(W0.W), W0.K, dp(ctx.bcbias), dp(s.Y[0]),
                          dp(ws.stats) if train else None, st)
    bn_fwd(ws, 64, B * geom.npoint * l1.ns, W0, s.bn[0], train)
    _mlp_tail_forward(ws, [L["sa0.1"], L["sa0.2"]], s, 1, l1.cap, l1.M_dev, l1.row_w, B * geom.npoint * l1.ns, train)
    lib.gaddpg_pool_fwd(dp(s.Y[2]), 128, dp(s.bn[2].scale), dp(s.bn[2].shift), dp(l1.seg_off), 0, l1.S, dp(s.out), dp(s.arg), st)
    # ---- SA2: gather [feats | dxyz | pad] rows, three row-GEMMs, pool
    s2 = ctx.sa[1]
    lib.gaddpg_gather_rows(dp(s.out), 128, dp(l1.new_xyz), geom.npoint, dp(l2.new_xyz), geom.npoint, dp(l2.row_seg),
                           dp(l2.row_src), l2.cap, l2.M_dev, dp(s2.G), s2.G.shape[1], st)
    _mlp_first_forward(ws, L["sa1.0"], s2, l2.cap, l2.M_dev, l2.row_w, B * geom.npoint * l2.ns, train)
    _mlp_tail_forward(ws, [L["sa1.1"], L["sa1.2"]], s2, 1, l2.cap, l2.M_dev, l2.row_w, B * geom.npoint * l2.ns, train)
    lib.gaddpg_pool_fwd(dp(s2.Y[2]), 256, dp(s2.bn[2].scale), dp(s2.bn[2].shift), dp(l2.seg_off), 0, l2.S, dp(s2.out),
                        dp(s2.arg), st)
    # ---- SA3: GroupAll over the 32 SA2 centroids (absolute xyz)
    s3 = ctx.sa[2]
    M3 = B * geom.npoint
    lib.gaddpg_gather_rows(dp(s2.out), 256, dp(l2.new_xyz), geom.npoint, None, geom.npoint, None, None, M3, None, dp(s3.G),
                           s3.G.shape[1], st)
    _mlp_first_forward(ws, L["sa2.0"], s3, M3, None, None, M3, train)
    _mlp_tail_forward(ws, [L["sa2.1"], L["sa2.2"]], s3, 1, M3, None, None, M3, train)
    lib.gaddpg_pool_fwd(dp(s3.Y[2]), 512, dp(s3.bn[2].scale), dp(s3.bn[2].shift), None, geom.npoint, B, dp(s3.out), dp(s3.arg), st)
    # ---- FC head: Linear + BN1d + ReLU twice
    f = ctx.fc
    F0, F1 = L["fc0"], L["fc1"]
    nt([nt_problem(op_plain(s3.out), F0.Wf, F0.Kp, f.Y[0], 1024, B, None, 1024, 512, bias=F0.bias,
                   stats=ws.stats if train else None)], OP_PLAIN, EPI_STORE)
    bn_fwd(ws, 1024, B, F0, f.bn[0], train)
    nt([nt_problem(op_bnrelu(f.Y[0], f.bn[0]), F1.Wf, F1.Kp, f.Y[1], 512, B, None, 512, 1024, bias=F1.bias,
                   stats=ws.stats if train else None)], OP_BNRELU, EPI_STORE)
    bn_fwd(ws, 512, B, F1, f.bn[1], train)
    lib.gaddpg_feat_finish(dp(f.Y[1]), 512, dp(f.bn[1].scale), dp(f.bn[1].shift), dp(time), float(time_offset), B, dp(ctx.feat),
                           516, st)
    return ctx.feat


def _mlp_first_forward(ws, Lp, s, M_max, M_dev, rw, count, train):
    nt([nt_problem(op_plain(s.G), Lp.Wf, Lp.Kp, s.Y[0], Lp.N, M_max, M_dev, Lp.N, Lp.Kp, stats=ws.stats if train else None,
                   srw=rw)], OP_PLAIN, EPI_STORE)
    bn_fwd(ws, Lp.N, count, Lp, s.bn[0], train)


def _mlp_tail_forward(ws, layers, s, first, M_max, M_dev, rw, count, train):
    for j, Lp in enumerate(layers):
        l = first + j
        nt([nt_problem(op_bnrelu(s.Y[l - 1], s.bn[l - 1]), Lp.Wf, Lp.Kp, s.Y[l], Lp.N, M_max, M_dev, Lp.N, Lp.Kp,
                       stats=ws.stats if train else None, srw=rw)], OP_BNRELU, EPI_STORE)
        bn_fwd(ws, Lp.N, count, Lp, s.bn[l], train)


def encoder_backward(ws, ef, ctx, sc, want_dw=True, want_dbc=False, accumulate=0, dfeat=None):
    """Backward of ``encoder_forward``.  Entry: either ``dfeat`` (B, >=512) is given (gradient w.r.t. ctx.feat;
    masked here), or the caller already produced sc.Dfc[1] and its BN sums in ws.stats through an EPI_DMASK
    epilogue (the fused path).  Writes parameter gradients into the arena (``want_dw``) and, for the broadcast
    channels, sc.dbc (B, Cb) (``want_dbc``)."""
    B = ctx.B
    st = current_stream()
    geom = ctx.geom
    l1, l2 = geom.lv
    L = ef.layers
    f = ctx.fc
    F0, F1 = L["fc0"], L["fc1"]
    if dfeat is not None:
        lib.gaddpg_dmask_stats(dp(dfeat), dfeat.shape[1], dp(f.Y[1]), 512, B, dp(f.bn[1].scale), dp(f.bn[1].shift),
                               dp(f.bn[1].mean), dp(f.bn[1].rstd), dp(sc.Dfc[1]), dp(ws.stats), st)
    s3 = ctx.sa[2]
    # ---- FC head
    bn_bwd(ws, 512, B, F1, f.bn[1], sc.bbfc[1], want_dw, accumulate)
    dy1 = op_bnbwd(sc.Dfc[1], f.Y[1], f.bn[1], sc.bbfc[1])
    if want_dw:
        tn(ws, dy1, op_bnrelu(f.Y[0], f.bn[0]), OP_BNBWD, OP_BNRELU, B, None, 512, 1024, F1.dW, 1024, 512, 1024,
           dbias=F1.dbias, accumulate=accumulate)
    nt([nt_problem(dy1, F1.WT, 512, sc.Dfc[0], 1024, B, None, 1024, 512, stats=ws.stats, Yprev=f.Y[0], ldyp=1024,
                   pbn=f.bn[0])], OP_BNBWD, EPI_DMASK)
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
                          dp(sc.dbc) if (want_dbc and ctx.Cb > 0) else None, dp(sc.dY1) if ctx.Cb > 0 else None,
                          dp(ws.sa1(B)), ws.sa1_ws_bytes, st)
    return sc.dbc.view(-1)[: B * ctx.Cb].view(B, ctx.Cb) if (want_dbc and ctx.Cb > 0) else None


def _sa_backward(ws, layers, s, sc, lvl, dOut, ld_dout, row_seg, fixed_len, M_max, M_dev, rw, count, want_dw, accumulate,
                 generic_l0):
    """Pool backward + the BN/conv layers of one SA level.  Always leaves D[lvl][0] and bb[lvl][0] (the BN-backward
    constants of layer 0) ready; with ``generic_l0`` it also runs layer 0's conv backward (dW, dG[lvl]) — SA1's
    first layer has its own kernel (gaddpg_sa1_l1_bwd)."""
    st = current_stream()
    D, bb = sc.D[lvl], sc.bb[lvl]
    C3 = layers[2].N
    lib.gaddpg_pool_bwd(dp(dOut), ld_dout, dp(s.out), dp(s.arg), dp(s.Y[2]), C3, dp(row_seg), fixed_len, M_max, M_dev,
                        dp(s.bn[2].mean), dp(s.bn[2].rstd), dp(D[2]), dp(ws.stats), st)
    for l in (2, 1):
        Lp = layers[l]
        bn_bwd(ws, Lp.N, count, Lp, s.bn[l], bb[l], want_dw, accumulate)
        dy = op_bnbwd(D[l], s.Y[l], s.bn[l], bb[l], rw=rw)
        if want_dw:
            tn(ws, dy, op_bnrelu(s.Y[l - 1], s.bn[l - 1]), OP_BNBWD, OP_BNRELU, M_max, M_dev, Lp.N, Lp.Kp, Lp.dW, Lp.K, Lp.N,
               Lp.K, accumulate=accumulate)
        nt([nt_problem(dy, Lp.WT, Lp.N, D[l - 1], Lp.Kp, M_max, M_dev, Lp.Kp, Lp.N, stats=ws.stats, Yprev=s.Y[l - 1],
                       ldyp=Lp.Kp, pbn=s.bn[l - 1])], OP_BNBWD, EPI_DMASK)
    L0 = layers[0]
    bn_bwd(ws, L0.N, count, L0, s.bn[0], bb[0], want_dw, accumulate)
    if generic_l0:
        dy = op_bnbwd(D[0], s.Y[0], s.bn[0], bb[0], rw=rw)
        if want_dw:
            tn(ws, dy, op_plain(s.G), OP_BNBWD, OP_PLAIN, M_max, M_dev, L0.N, L0.Kp, L0.dW, L0.K, L0.N, L0.K, rot=L0.rot,
               accumulate=accumulate)
        nt([nt_problem(dy, L0.WT, L0.N, sc.dG[lvl], L0.Kp, M_max, M_dev, L0.Kp, L0.N)], OP_BNBWD, EPI_STORE)
