"""GPU: the fused SA1 chain (csrc/sa1_fused.cu: gather -> conv0 -> BN+ReLU -> conv1 -> BN+ReLU -> conv2 -> statistics +
per-group extremes, three recompute phases, TMA-fed weights, TMA-stored activations in keep mode) against the unfused kernels
it replaces (sa1_l1_fwd + two gemm_nt + pool_fwd, themselves held to the CPU oracle by test_encoder_gpu.py), and one full DDPG
step through the fused chain against the CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _run(engine, ef, cloud, bc, Cp, fused, keep, train, device):
    B, C, Np = cloud.shape
    ws = engine.Workspace(device)
    geom = engine.Geometry(B, Np - 6, device).build(cloud, 6)
    ctx = engine.EncoderCtx(B, (geom.lv[0].cap, geom.lv[1].cap), engine.WIDTHS, device)
    was, engine.FUSED_SA1 = engine.FUSED_SA1, fused
    try:
        feat = engine.encoder_forward(ws, ef, geom, cloud, 6, Cp, bc, ctx, train=train, keep=keep).clone()
    finally:
        engine.FUSED_SA1 = was
    torch.cuda.synchronize()
    return ws, geom, ctx, feat


@pytest.mark.parametrize("B,N,channels,with_action", [(8, 512, 4, True), (8, 512, 4, False), (5, 700, 6, True), (1, 1024, 4, False),
                                                      (256, 4096, 6, True)])
def test_fused_sa1_matches_unfused_kernels(cuda, B, N, channels, with_action):
    from gaddpg_b200 import engine, synthetic
    from tests.test_encoder_gpu import _build

    Cb = (10 - channels) if with_action else 0
    ora, mine, ef = _build(channels + Cb, 31, cuda)
    assert ef.sa1f_ok
    batch = synthetic.make_batch(B, N, step=2, channels=channels)
    cloud = torch.from_numpy(batch["point_state_batch"]).to(cuda)
    bc = torch.from_numpy(batch["action_batch"][:, :Cb]).to(cuda).contiguous() if with_action else None
    for train in (True, False):
        for keep in ((False, True) if train else (False,)):
            _, g0, c0, f0 = _run(engine, ef, cloud, bc, channels, False, keep, train, cuda)
            _, g1, c1, f1 = _run(engine, ef, cloud, bc, channels, True, keep, train, cuda)
            M = int(g0.lv[0].seg_off[-1])
            s0, s1 = c0.sa[0], c1.sa[0]
            tag = (B, N, channels, with_action, train, keep)
            for l in range(3):
                for k in ("scale", "shift"):
                    assert _rel(getattr(s1.bn[l], k), getattr(s0.bn[l], k)) < 2e-5, (tag, l, k)
                if train:
                    assert _rel(s1.bn[l].mean, s0.bn[l].mean) < 2e-5 and _rel(s1.bn[l].rstd, s0.bn[l].rstd) < 2e-5, (tag, l)
            assert _rel(s1.out, s0.out) < 2e-5, tag                    # pooled SA1 features
            assert _rel(f1[:, :512], f0[:, :512]) < 5e-5, tag          # encoder output
            if keep:
                for l in range(3):
                    assert _rel(s1.Y[l][:M], s0.Y[l][:M]) < 2e-5, (tag, "Y%d" % l)
                # arg-max rows where the pooled value is positive (elsewhere ReLU zeroes the gradient and the row is immaterial:
                # pool_fwd reports the group's first row there, the fused chain the row of the pre-ReLU extreme): identical
                # except where two rows tie within rounding; where they differ the pooled values agree
                diff = (s1.arg != s0.arg) & (s0.out > 0)
                assert float(diff.float().mean()) < 2e-3, (tag, float(diff.float().mean()))
                y2 = s0.Y[2]
                ia, ib = s1.arg[diff].long(), s0.arg[diff].long()
                cols = diff.nonzero()[:, 1]
                if ia.numel():
                    va, vb = y2[ia, cols], y2[ib, cols]
                    assert float(((va - vb).abs() / (vb.abs() + 1e-6)).max()) < 1e-4, tag
                assert int(s1.arg.min()) >= 0 and int(s1.arg.max()) < M


def test_fused_sa1_backward_consumes_tma_stored_activations(cuda):
    """keep mode: the backward kernels read the Y0/Y1/Y2 the fused chain stored through TMA; parameter gradients and the
    action gradient equal those obtained from the unfused forward (same kernels downstream) to rounding."""
    from gaddpg_b200 import engine, synthetic
    from tests.test_encoder_gpu import _build

    B, N = 16, 1024
    ora, mine, ef = _build(10, 33, cuda)
    batch = synthetic.make_batch(B, N, step=4)
    cloud = torch.from_numpy(batch["point_state_batch"]).to(cuda)
    bc = torch.from_numpy(batch["action_batch"]).to(cuda).contiguous()
    R = torch.zeros(B, 516, device=cuda)
    R[:, :512] = torch.from_numpy(np.random.RandomState(3).randn(B, 512).astype(np.float32)).to(cuda)
    grads = {}
    for fused in (False, True):
        ws, geom, ctx, _ = _run(engine, ef, cloud, bc, 4, fused, True, True, cuda)
        sc = engine.BwdScratch(B, (geom.lv[0].cap, geom.lv[1].cap), engine.WIDTHS, cuda)
        dbc = engine.encoder_backward(ws, ef, ctx, sc, want_dw=True, want_dbc=True, dfeat=R)
        torch.cuda.synchronize()
        grads[fused] = ({k: p.grad.clone() for k, p in mine.named_parameters()}, dbc.clone())
    for k in grads[False][0]:
        if k.endswith(("1.0.bias", "1.3.bias")):
            continue
        assert _rel(grads[True][0][k], grads[False][0][k]) < 5e-2, k      # a few ReLU / max-pool routings may differ (1e-6 forward noise)
    num = sum(float(((grads[True][0][k] - grads[False][0][k]) ** 2).sum()) for k in grads[False][0] if not k.endswith(("1.0.bias", "1.3.bias")))
    den = sum(float((grads[False][0][k] ** 2).sum()) for k in grads[False][0] if not k.endswith(("1.0.bias", "1.3.bias")))
    assert (num / den) ** 0.5 < 2e-3
    assert _rel(grads[True][1], grads[False][1]) < 3e-2      # B x Cb sums over every point: routing flips reach all of them


def test_ddpg_step_through_the_fused_chain_matches_oracle(cuda):
    """One DDPG update with every SA1 forward (F1..F4 of an odd step, keep and no-keep) on the fused chain: all 11 scalars
    within 1e-4 of the CPU oracle, like the default path."""
    from gaddpg_b200 import agent as ag, engine, synthetic
    from gaddpg_b200.config import LOSS_KEYS
    from oracle.ddpg_cpu import OracleAgent

    B, N = 8, 512
    ora = OracleAgent("DDPG", seed=123456)
    was, engine.FUSED_SA1 = engine.FUSED_SA1, True
    try:
        mine = ag.make_agent("DDPG", seed=123456)
        batch = synthetic.make_batch(B, N, step=0)
        u = np.random.RandomState(1).rand(B, 6).astype(np.float32)
        o = ora.update_parameters(batch, noise_u=u)
        m = mine.update_parameters(batch, 1, 0, noise_u=u)
    finally:
        engine.FUSED_SA1 = was
    for k in LOSS_KEYS:
        assert (np.isnan(m[k]) and np.isnan(o[k])) or abs(m[k] - o[k]) <= 1e-6 + 1e-4 * abs(o[k]), (k, m[k], o[k])
