"""Diagnostic (GPU box): per-tensor error of the CUDA encoder backward and of the fp32 oracle against a float64 gradient,
for the tensor-core (tc=3), mixed (tc=1) and FFMA (tc=0) paths.  env: B, N."""
import copy, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from gaddpg_b200 import engine, synthetic
from gaddpg_b200.capi import lib
from tests.test_encoder_gpu import _build, _oracle_forward
from tests.f64ref import f64_ops

B, N = int(os.environ.get("B", 16)), int(os.environ.get("N", 256))
cuda = torch.device("cuda:0")
torch.set_num_threads(os.cpu_count())
for with_action in (True,):
    in_features = 10 if with_action else 4
    ora, mine, ef = _build(in_features, 11, cuda)
    batch = synthetic.make_batch(B, N, step=5)
    cloud = torch.from_numpy(batch["point_state_batch"]); action = torch.from_numpy(batch["action_batch"])
    R = torch.from_numpy(np.random.RandomState(3).randn(B, 512).astype(np.float32))
    ora.train()
    o64 = copy.deepcopy(ora).double()
    act_o = action.clone().requires_grad_(True)
    z_o = _oracle_forward(ora, cloud, act_o); (z_o * R).sum().backward()
    a64 = action.double().clone().requires_grad_(True)
    with f64_ops():
        z64 = _oracle_forward(o64, cloud.double(), a64); (z64 * R.double()).sum().backward()
    po, p64 = dict(ora.named_parameters()), dict(o64.named_parameters())
    for tc in (3, 1, 0):
        lib.gaddpg_set_tensor_core(tc)
        ws = engine.Workspace(cuda)
        cl = cloud.to(cuda)
        geom = engine.Geometry(B, N, cuda).build(cl, 6)
        caps = (geom.lv[0].cap, geom.lv[1].cap)
        ctx = engine.EncoderCtx(B, caps, engine.WIDTHS, cuda); sc = engine.BwdScratch(B, caps, engine.WIDTHS, cuda)
        feat = engine.encoder_forward(ws, ef, geom, cl, 6, 4, action.to(cuda).contiguous(), ctx, train=True)
        dfeat = torch.zeros(B, 516, device=cuda); dfeat[:, :512] = R.to(cuda)
        dbc = engine.encoder_backward(ws, ef, ctx, sc, want_dw=True, want_dbc=True, dfeat=dfeat)
        torch.cuda.synchronize()
        print("== B=%d N=%d tc=%d  rows sa1=%d sa2=%d   z: cuda %.2e ora %.2e" % (B, N, tc, int(geom.lv[0].seg_off[-1]), int(geom.lv[1].seg_off[-1]),
              float((feat[:, :512].cpu().double() - z64).abs().max() / z64.abs().max()), float((z_o.double() - z64).abs().max() / z64.abs().max())))
        pm = dict(mine.named_parameters())
        for k in po:
            g64 = p64[k].grad; den = float(g64.norm()) + 1e-300
            ec = float((pm[k].grad.cpu().double() - g64).norm()) / den; eo = float((po[k].grad.double() - g64).norm()) / den
            mc = float((pm[k].grad.cpu().double() - g64).abs().max()) / (float(g64.abs().max()) + 1e-300)
            mo = float((po[k].grad.double() - g64).abs().max()) / (float(g64.abs().max()) + 1e-300)
            print("  %-28s L2: cuda %.2e ora %.2e | max: cuda %.2e ora %.2e  %s" % (k, ec, eo, mc, mo, "<<<" if ec > 10 * eo + 1e-5 else ""))
        den = float(a64.grad.norm())
        print("  d/d(action) L2: cuda %.2e ora %.2e" % (float((dbc.cpu().double() - a64.grad).norm()) / den, float((act_o.grad.double() - a64.grad).norm()) / den))
lib.gaddpg_set_tensor_core(3)
