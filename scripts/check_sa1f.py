"""Debug / A-B of the fused SA1 chain against the unfused kernels, phase by phase (GPU box).  env: B, N, CH, ACT"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaddpg_b200 import engine, synthetic
from gaddpg_b200.capi import lib, current_stream
from gaddpg_b200.structs import dp
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_encoder_gpu import _build

B, N, CH, ACT = int(os.environ.get("B", 8)), int(os.environ.get("N", 512)), int(os.environ.get("CH", 4)), int(os.environ.get("ACT", 1))
dev = torch.device("cuda:0")
Cb = (10 - CH) if ACT else 0
ora, mine, ef = _build(CH + Cb, 31, dev)
batch = synthetic.make_batch(B, N, step=2, channels=CH)
cloud = torch.from_numpy(batch["point_state_batch"]).to(dev)
bc = torch.from_numpy(batch["action_batch"][:, :Cb]).to(dev).contiguous() if ACT else None
rel = lambda a, b: float((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30))

def run(fused, keep):
    ws = engine.Workspace(dev)
    geom = engine.Geometry(B, N, dev).build(cloud, 6)
    ctx = engine.EncoderCtx(B, (geom.lv[0].cap, geom.lv[1].cap), engine.WIDTHS, dev)
    engine.FUSED_SA1 = fused
    feat = engine.encoder_forward(ws, ef, geom, cloud, 6, CH, bc, ctx, train=True, keep=keep).clone()
    torch.cuda.synchronize()
    return geom, ctx, feat

g0, c0, f0 = run(False, True)
M = int(g0.lv[0].seg_off[-1])
print("B=%d N=%d CH=%d Cb=%d  M=%d cap=%d grid=%d" % (B, N, CH, Cb, M, g0.lv[0].cap, lib.gaddpg_sa1_fused_grid(g0.lv[0].cap)), flush=True)
for keep in (True, False):
    g1, c1, f1 = run(True, keep)
    s0, s1 = c0.sa[0], c1.sa[0]
    for l in range(3):
        line = "keep=%d layer %d: scale %.2e shift %.2e mean %.2e rstd %.2e" % (keep, l, rel(s1.bn[l].scale, s0.bn[l].scale), rel(s1.bn[l].shift, s0.bn[l].shift),
                                                                               rel(s1.bn[l].mean, s0.bn[l].mean), rel(s1.bn[l].rstd, s0.bn[l].rstd))
        if keep:
            line += "  Y%d %.2e" % (l, rel(s1.Y[l][:M], s0.Y[l][:M]))
        print(line, flush=True)
    print("   pooled out %.2e   arg mismatch %.4f%%   feat %.2e" % (rel(s1.out, s0.out), 100 * float(((s1.arg != s0.arg) & (s0.out > 0)).float().mean()), rel(f1[:, :512], f0[:, :512])), flush=True)
# timing
for fused in (False, True):
    for keep in (True, False):
        ws = engine.Workspace(dev); geom = engine.Geometry(B, N, dev).build(cloud, 6)
        ctx = engine.EncoderCtx(B, (geom.lv[0].cap, geom.lv[1].cap), engine.WIDTHS, dev)
        Cbb = 0 if bc is None else bc.shape[1]
        ctx.bc, ctx.Cp, ctx.Cb, ctx.cloud, ctx.skip, ctx.geom = bc, CH, Cbb, cloud, 6, geom
        fn = engine._sa1_fused_forward if fused else engine._sa1_unfused_forward
        for _ in range(3):
            fn(ws, ef, geom, cloud, 6, CH, bc, Cbb, ctx, True, None, keep)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn(ws, ef, geom, cloud, 6, CH, bc, Cbb, ctx, True, None, keep)
        e1.record(); torch.cuda.synchronize()
        print("SA1 forward fused=%d keep=%d: %.1f us per pass" % (fused, keep, 100 * e0.elapsed_time(e1)), flush=True)
# ---- per-phase timing of the fused chain (stale BN constants are fine for timing)
ws = engine.Workspace(dev); geom = engine.Geometry(B, N, dev).build(cloud, 6)
ctx = engine.EncoderCtx(B, (geom.lv[0].cap, geom.lv[1].cap), engine.WIDTHS, dev)
l1 = geom.lv[0]; s = ctx.sa[0]; L = ef.layers; b = ws.sa1f(l1.S, l1.cap)
Cbb = 0 if bc is None else bc.shape[1]
engine._sa1_fused_forward(ws, ef, geom, cloud, 6, CH, bc, Cbb, ctx, True, None, True)
def phase(k, keep):
    lib.gaddpg_sa1_fused_fwd(k, dp(cloud), cloud.shape[1] * cloud.shape[2], cloud.shape[2], 6, CH, dp(bc), Cbb, dp(l1.new_xyz), 32, dp(l1.seg_off),
                             dp(l1.row_seg), dp(l1.row_src), dp(l1.row_w), l1.cap, l1.M_dev, dp(ef.sa1f_w), dp(s.bn[0].scale), dp(s.bn[0].shift),
                             dp(s.bn[1].scale), dp(s.bn[1].shift), dp(L["sa0.2"].gamma), dp(ws.stats), dp(s.Y[k - 1]) if keep else None,
                             dp(b.ext), dp(b.arg), dp(b.part_ext), dp(b.part_arg), dp(b.seg_part), current_stream())
def fin():
    lib.gaddpg_sa1_pool_finalize(dp(b.ext), dp(b.arg), dp(b.part_ext), dp(b.part_arg), dp(b.seg_part), dp(L["sa0.2"].gamma),
                                 dp(s.bn[2].scale), dp(s.bn[2].shift), l1.S, dp(s.out), dp(s.arg), current_stream())
def bnf():
    engine.bn_fwd(ws, 128, B * 32 * 64, L["sa0.2"], s.bn[2], True, None)
def tm(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n
for keep in (False, True):
    print("keep=%d: phase1 %.1f us  phase2 %.1f us  phase3 %.1f us" % (keep, tm(lambda: phase(1, keep)), tm(lambda: phase(2, keep)), tm(lambda: phase(3, keep))), flush=True)
print("pool finalize %.1f us   bn_finalize(128) %.1f us" % (tm(fin), tm(bnf)), flush=True)
