"""GPU: the per-level C-ABI composites (gaddpg_sa_forward / gaddpg_sa_backward / gaddpg_adam_fused_step, csrc/composite.cu —
SURVEY.md §8(b)) drive set-abstraction levels WITHOUT the Python sequencing of engine.py: a generic level (SA2: gathered
rows in, pooled features out) and the first level (SA1: straight from the cloud, with the broadcast action channels),
forward (train and eval BatchNorm) and backward, bit-identical to the engine path that runs the same kernels, which in turn is
held to the CPU oracle by test_encoder_gpu.py."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _p(t):
    return None if t is None else t.data_ptr()


def _level(engine, ef, geom, lvl, ctx_in, B, device, S, M_max, cloud=None, bc=None, Cp=0):
    """gaddpg_sa_level for SA level ``lvl`` (0 = first level) with its OWN output / scratch buffers."""
    from gaddpg_b200.structs import SALevel

    f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=device)  # noqa: E731
    lv = geom.lv[lvl]
    d = SALevel()
    keep = []
    for l in range(3):
        L = ef.layers["sa%d.%d" % (lvl, l)]
        bufs = dict(scale=f(L.N), shift=f(L.N), mean=f(L.N), rstd=f(L.N), Y=f(M_max, L.N), D=f(M_max, L.N), bw_g=f(L.N), bw_m1=f(L.N),
                    bw_m2=f(L.N), dW=f(L.N, L.K), dgamma=f(L.N), dbeta=f(L.N), rm=L.rm.clone(), rv=L.rv.clone(), nbt=L.nbt.clone())
        keep.append(bufs)
        y = d.layer[l]
        raw0 = lvl == 0 and l == 0
        y.W, y.WT, y.N, y.K, y.Kp = _p(L.W if raw0 else L.Wf), _p(L.WT), L.N, L.K, (L.K if raw0 else L.Kp)
        y.gamma, y.beta, y.running_mean, y.running_var, y.num_batches_tracked = _p(L.gamma), _p(L.beta), _p(bufs["rm"]), _p(bufs["rv"]), _p(bufs["nbt"])
        for k in ("scale", "shift", "mean", "rstd", "Y", "D", "bw_g", "bw_m1", "bw_m2", "dW", "dgamma", "dbeta"):
            setattr(y, k, _p(bufs[k]))
    d.B, d.S, d.M_max, d.M_dev, d.count = B, S, M_max, lv.M_dev, float(B * geom.npoint * lv.ns)
    d.seg_off, d.row_seg, d.row_src, d.row_w, d.fixed_len = _p(lv.seg_off), _p(lv.row_seg), _p(lv.row_src), _p(lv.row_w), 0
    out, arg = f(S, ef.layers["sa%d.2" % lvl].N), torch.zeros(S, ef.layers["sa%d.2" % lvl].N, dtype=torch.int32, device=device)
    d.out, d.arg = _p(out), _p(arg)
    if lvl == 0:
        Bc, C, Np = cloud.shape
        d.cloud, d.cloud_stride_b, d.cloud_stride_c, d.skip, d.Cp = _p(cloud), C * Np, Np, 6, Cp
        d.bc, d.Cb, d.ctr, d.npoint = _p(bc), (0 if bc is None else bc.shape[1]), _p(lv.new_xyz), geom.npoint
    else:
        G = ctx_in.sa[lvl].G
        d.G, d.ldg, d.rot = _p(G), G.shape[1], ef.layers["sa%d.0" % lvl].rot
    return d, keep, out, arg


@pytest.mark.parametrize("lvl", [1, 0])
def test_sa_level_composites_equal_the_engine_path(cuda, lvl):
    from gaddpg_b200 import engine, synthetic
    from gaddpg_b200.capi import current_stream, lib
    from tests.test_encoder_gpu import _build

    B, N = 8, 512
    ora, mine, ef = _build(10, 41, cuda)
    batch = synthetic.make_batch(B, N, step=6)
    cloud = torch.from_numpy(batch["point_state_batch"]).to(cuda)
    bc = torch.from_numpy(batch["action_batch"]).to(cuda).contiguous()
    was = (engine.SPARSE_POOL, engine.FUSED_SA1)
    engine.SPARSE_POOL, engine.FUSED_SA1 = False, False       # the composite uses the dense pool backward of every level
    try:
        ws = engine.Workspace(cuda)
        geom = engine.Geometry(B, N, cuda).build(cloud, 6)
        caps = (geom.lv[0].cap, geom.lv[1].cap)
        ctx = engine.EncoderCtx(B, caps, engine.WIDTHS, cuda)
        sc = engine.BwdScratch(B, caps, engine.WIDTHS, cuda)
        running0 = ef.buffers.p.clone()
        engine.encoder_forward(ws, ef, geom, cloud, 6, 4, bc, ctx, train=True)
        R = torch.zeros(B, 516, device=cuda)
        R[:, :512] = torch.from_numpy(np.random.RandomState(3).randn(B, 512).astype(np.float32)).to(cuda)
        dbc_ref = engine.encoder_backward(ws, ef, ctx, sc, want_dw=True, want_dbc=True, dfeat=R)
        torch.cuda.synchronize()
        ref_running = ef.buffers.p.clone()
        ef.buffers.p.copy_(running0)                           # the composite gets the pre-pass running statistics (its own copies)
        S, M_max = geom.lv[lvl].S, geom.lv[lvl].cap
        d, keep, out, arg = _level(engine, ef, geom, lvl, ctx, B, cuda, S, M_max, cloud=cloud, bc=bc, Cp=4)
        for l in range(3):   # fresh running-stat copies hold the PRE-pass values now
            L = ef.layers["sa%d.%d" % (lvl, l)]
            keep[l]["rm"].copy_(L.rm), keep[l]["rv"].copy_(L.rv)
        nbytes = int(lib.gaddpg_sa_level_workspace_bytes(B, M_max))
        wsb = torch.zeros(nbytes // 4 + 4, dtype=torch.float32, device=cuda)
        lib.gaddpg_sa_forward(ctypes.byref(d), 1, wsb.data_ptr(), nbytes, current_stream())
        torch.cuda.synchronize()
        s = ctx.sa[lvl]
        M = int(geom.lv[lvl].seg_off[-1])
        assert torch.equal(out, s.out) and torch.equal(arg, s.arg), "pooled features / arg-max rows"
        for l in range(3):
            assert torch.equal(keep[l]["Y"][:M], s.Y[l][:M]) and torch.equal(keep[l]["scale"], s.bn[l].scale) and torch.equal(keep[l]["rstd"], s.bn[l].rstd)
        ef.buffers.p.copy_(ref_running)
        for l in range(3):   # running statistics after one training pass
            L = ef.layers["sa%d.%d" % (lvl, l)]
            assert torch.equal(keep[l]["rm"], L.rm) and torch.equal(keep[l]["rv"], L.rv), l
        # ---- backward from the same pooled gradient the engine used
        if lvl == 1:
            dOut, ld = sc.dG[2], sc.dG[2].shape[1]
            dG = torch.zeros_like(sc.dG[1])
            lib.gaddpg_sa_backward(ctypes.byref(d), dOut.data_ptr(), ld, 1, 0, dG.data_ptr(), dG.shape[1], None, wsb.data_ptr(), nbytes,
                                   current_stream())
            torch.cuda.synchronize()
            assert torch.equal(dG[:M], sc.dG[1][:M]), "gradient w.r.t. the gathered input rows"
        else:
            dbc = torch.zeros(B, 8, device=cuda)
            lib.gaddpg_sa_backward(ctypes.byref(d), sc.dout[0].data_ptr(), 128, 1, 0, None, 0, dbc.data_ptr(), wsb.data_ptr(), nbytes,
                                   current_stream())
            torch.cuda.synchronize()
            assert torch.equal(dbc.view(-1)[: B * 6].view(B, 6), dbc_ref), "gradient w.r.t. the broadcast (action) channels"
        for l in range(3):
            L = ef.layers["sa%d.%d" % (lvl, l)]
            assert torch.equal(keep[l]["dW"], L.dW) and torch.equal(keep[l]["dgamma"], L.dgamma) and torch.equal(keep[l]["dbeta"], L.dbeta), l
        # ---- eval mode: running statistics instead of batch statistics
        engine.encoder_forward(ws, ef, geom, cloud, 6, 4, bc, ctx, train=False)   # (refills the gathered input rows d.G points at)
        lib.gaddpg_sa_forward(ctypes.byref(d), 0, wsb.data_ptr(), nbytes, current_stream())
        torch.cuda.synchronize()
        assert torch.equal(out, ctx.sa[lvl].out)
    finally:
        engine.SPARSE_POOL, engine.FUSED_SA1 = was


def test_adam_fused_step_matches_torch_adam(cuda):
    """§8(b) adam_fused_step: L2 weight decay, eps 1e-5 (utils.py:969-970), clip coefficient, Polyak target in one pass."""
    from gaddpg_b200.capi import current_stream, lib

    n = 10_000
    rs = np.random.RandomState(0)
    p0, g0 = rs.randn(n).astype(np.float32), rs.randn(n).astype(np.float32)
    ref = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([ref], lr=3e-4, eps=1e-5, weight_decay=1e-5)
    p, g, m, v = (torch.from_numpy(x.copy()).to(cuda) for x in (p0, g0, np.zeros(n, np.float32), np.zeros(n, np.float32)))
    tgt = torch.from_numpy(p0.copy()).to(cuda)
    clip = torch.tensor([0.25], device=cuda)
    for step in range(1, 4):
        ref.grad = torch.from_numpy(g0 * 0.25)
        opt.step()
        lib.gaddpg_adam_fused_step(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 3e-4, 0.9, 0.999, 1e-5, 1e-5, step, 1.0,
                                   clip.data_ptr(), 0, tgt.data_ptr(), 1e-4, current_stream())
    torch.cuda.synchronize()
    assert float((p.cpu() - ref.detach()).abs().max()) < 2e-6
    assert float((tgt.cpu() - torch.from_numpy(p0)).abs().max()) < 1e-5 and not torch.equal(tgt.cpu(), torch.from_numpy(p0))
