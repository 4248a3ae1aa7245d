"""Diagnostic (GPU box): encoder backward with the tcgen05 path vs the FFMA path vs the CPU oracle, per tensor."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from gaddpg_b200 import engine, synthetic
from gaddpg_b200.capi import lib
from tests.test_encoder_gpu import _build, _oracle_forward, _rel

cuda = torch.device("cuda")
for (B, N, with_action) in ((16, 256, False), (16, 256, True), (6, 1024, True)):
    in_features = 10 if with_action else 4
    batch = synthetic.make_batch(B, N, step=5)
    cloud = torch.from_numpy(batch["point_state_batch"]); action = torch.from_numpy(batch["action_batch"]) if with_action else None
    R = torch.from_numpy(np.random.RandomState(3).randn(B, 512).astype(np.float32))
    res = {}
    for tc in (0, 1):
        lib.gaddpg_set_tensor_core(tc)
        ora, mine, ef = _build(in_features, 11, cuda)
        if tc == 0:
            ora.train(); act_o = action.clone().requires_grad_(True) if with_action else None
            z_o = _oracle_forward(ora, cloud, act_o); (z_o * R).sum().backward()
            og = {k: p.grad.clone() for k, p in ora.named_parameters()}
        ws = engine.Workspace(cuda)
        geom = engine.Geometry(B, N, cuda).build(cloud.to(cuda), 6)
        caps = (geom.lv[0].cap, geom.lv[1].cap)
        ctx = engine.EncoderCtx(B, caps, engine.WIDTHS, cuda); sc = engine.BwdScratch(B, caps, engine.WIDTHS, cuda)
        bc = action.to(cuda).contiguous() if with_action else None
        feat = engine.encoder_forward(ws, ef, geom, cloud.to(cuda), 6, 4, bc, ctx, time=None, train=True)
        dfeat = torch.zeros(B, 516, device=cuda); dfeat[:, :512] = R.to(cuda)
        engine.encoder_backward(ws, ef, ctx, sc, want_dw=True, want_dbc=with_action, dfeat=dfeat)
        torch.cuda.synchronize()
        res[tc] = ({k: p.grad.clone().cpu() for k, p in mine.named_parameters()}, feat[:, :512].clone().cpu(), int(geom.lv[0].seg_off[-1]))
    print("== B=%d N=%d action=%s  M1=%d  fwd: ffma %.2e tc %.2e" % (B, N, with_action, res[0][2], _rel(res[0][1], z_o), _rel(res[1][1], z_o)))
    for k in og:
        e0, e1, e01 = _rel(res[0][0][k], og[k]), _rel(res[1][0][k], og[k]), _rel(res[1][0][k], res[0][0][k])
        if max(e0, e1) > 2e-4 and not k.endswith(("1.0.bias", "1.3.bias")):
            print("  %-24s ffma-vs-oracle %.2e   tc-vs-oracle %.2e   tc-vs-ffma %.2e" % (k, e0, e1, e01))
