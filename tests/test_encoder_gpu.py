"""GPU parity of the fused CUDA encoder (forward, train-mode BatchNorm statistics, eval mode, and the
hand-written backward) against the CPU oracle's PointNet++ encoder under autograd.
Tolerances: 1e-4 relative to the tensor's max magnitude (north_star: fp32 outputs within 1e-4)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _build(in_features, seed, device):
    from gaddpg_b200 import engine, nets
    from oracle import nets_cpu

    torch.manual_seed(seed)
    ora = nets_cpu.make_encoder(in_features)
    # non-trivial BN affine parameters so gamma/beta gradients and the sign-dependent pooling are exercised
    with torch.no_grad():
        for m in ora.modules():
            if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                m.weight.uniform_(-1.0, 1.5)
                m.bias.uniform_(-0.3, 0.3)
    mine = nets.make_encoder_params(in_features)
    mine.load_state_dict(ora.state_dict())
    ef = engine.EncoderFlat(mine, device)
    return ora, mine, ef


def _oracle_forward(ora, cloud, action):
    from oracle.nets_cpu import PointFeature

    x = cloud[..., 6:]
    if action is not None:
        x = torch.cat((x, action.unsqueeze(2).expand(-1, -1, x.shape[2])), 1)
    x = x.contiguous()
    xyz = x.transpose(1, -1)[..., :3].contiguous()
    return PointFeature.encode(ora, xyz, x)


@pytest.mark.parametrize("B,N,with_action,tc", [(16, 256, False, 3), (16, 256, True, 3), (6, 1024, True, 3), (16, 256, True, 1),
                                                  (16, 256, True, 0), (4, 512, False, 0)])
def test_encoder_forward_backward_matches_oracle(cuda, B, N, with_action, tc):
    """tc=3: every row-GEMM (NT and TN) with N,K >= 32 runs on the tcgen05 3xTF32 kernels; tc=1: only the wide SA1 layers
    (resident-weight kernel, row count >= 8192); tc=0 forces the FP32 FFMA kernels everywhere.  Batches of >= 6 keep the BatchNorm1d head well enough conditioned that the two FP32
    evaluations take the same side of (almost) every ReLU / max-pool kink; gradients are still compared with a
    kink-tolerant pair of bounds: tight on the typical tensor, loose on the worst one (DESIGN.md "Parity")."""
    from gaddpg_b200 import engine, synthetic
    from gaddpg_b200.capi import lib

    lib.gaddpg_set_tensor_core(tc)

    in_features = 10 if with_action else 4
    ora, mine, ef = _build(in_features, 11, cuda)
    batch = synthetic.make_batch(B, N, step=5)
    cloud = torch.from_numpy(batch["point_state_batch"])
    action = torch.from_numpy(batch["action_batch"]) if with_action else None
    time = torch.from_numpy(batch["time_batch"])
    R = torch.from_numpy(np.random.RandomState(3).randn(B, 512).astype(np.float32))

    # ---- oracle (CPU, autograd)
    ora.train()
    act_o = action.clone().requires_grad_(True) if with_action else None
    z_o = _oracle_forward(ora, cloud, act_o)
    (z_o * R).sum().backward()

    # ---- CUDA
    ws = engine.Workspace(cuda)
    geom = engine.Geometry(B, N, cuda).build(cloud.to(cuda), 6)
    caps = (geom.lv[0].cap, geom.lv[1].cap)
    ctx = engine.EncoderCtx(B, caps, engine.WIDTHS, cuda)
    sc = engine.BwdScratch(B, caps, engine.WIDTHS, cuda)
    cl = cloud.to(cuda)
    bc = action.to(cuda).contiguous() if with_action else None
    feat = engine.encoder_forward(ws, ef, geom, cl, 6, 4, bc, ctx, time=time.to(cuda), time_offset=-1.0, train=True)
    torch.cuda.synchronize()
    assert _rel(feat[:, :512], z_o) < 1e-4
    assert torch.equal(feat[:, 512].cpu(), time - 1.0) and float(feat[:, 513:].abs().max()) == 0.0
    # running statistics / counters updated exactly like torch's BatchNorm in train mode
    sd_o, sd_m = ora.state_dict(), mine.state_dict()
    for k in sd_o:
        if "running" in k:
            assert _rel(sd_m[k], sd_o[k]) < 1e-4, k
        if "num_batches_tracked" in k:
            assert int(sd_m[k]) == int(sd_o[k]) == 1, k
    dfeat = torch.zeros(B, 516, device=cuda)
    dfeat[:, :512] = R.to(cuda)
    dbc = engine.encoder_backward(ws, ef, ctx, sc, want_dw=True, want_dbc=with_action, dfeat=dfeat)
    torch.cuda.synchronize()
    worst = {}
    for (k, po), (_, pm) in zip(ora.named_parameters(), mine.named_parameters()):
        worst[k] = _rel(pm.grad, po.grad)
    # Linear biases in front of BatchNorm1d have an exactly-zero true gradient (pure rounding noise on both sides)
    errs = {k: v for k, v in worst.items() if not k.endswith(("1.0.bias", "1.3.bias"))}
    vals = sorted(errs.values())
    # typical tensor: rounding level; worst tensor: a few ReLU / max-pool kink flips.  The 3xTF32 path carries ~3x the
    # forward rounding error of FFMA (1e-5 instead of 5e-6 at z), so it flips a few more elements — every flip moves one
    # full gradient element, which shows as 1e-3..5e-2 on the tensors below it (tests/diag/diag_tc_encoder.py)
    assert vals[len(vals) // 2] < (2e-3 if tc else 2e-4), errs
    assert vals[-1] < (1e-1 if tc else 2e-2), errs
    for k in ("1.0.bias", "1.3.bias"):
        pm = dict(mine.named_parameters())[k]
        assert float(pm.grad.abs().max()) < 1e-4 * float(R.abs().max()) * B
    if with_action:
        assert _rel(dbc, act_o.grad) < (5e-2 if tc else 2e-2)
    # ---- float64 referee for the kink-tolerant bounds above: over all gradient elements the CUDA backward is as close to
    # the float64 gradient as the fp32 oracle's autograd is (tests/f64ref.py)
    import copy

    from tests.f64ref import f64_ops, referee_elems

    o64 = copy.deepcopy(ora).double()
    o64.zero_grad()
    for m_ in o64.modules():   # the forward above already updated the running statistics once; irrelevant in train mode
        if isinstance(m_, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            m_.reset_running_stats()
    act64 = action.double().clone().requires_grad_(True) if with_action else None
    with f64_ops():
        z64 = _oracle_forward(o64, cloud.double(), act64)
        (z64 * R.double()).sum().backward()
    assert _rel(feat[:, :512], z64) < 1e-4
    keys = [k for k, _ in ora.named_parameters() if not k.endswith(("1.0.bias", "1.3.bias"))]
    po, pm, p64 = dict(ora.named_parameters()), dict(mine.named_parameters()), dict(o64.named_parameters())
    rep = referee_elems("encoder parameter gradients (tc=%d)" % tc, [pm[k].grad.cpu().numpy() for k in keys],
                        [po[k].grad.numpy() for k in keys], [p64[k].grad.numpy() for k in keys])
    print("encoder grads vs float64, element errors / tensor scale (cuda, oracle32):", rep)
    if with_action:   # 6 x B numbers: every one of them sums over all points, so routing flips show in all — quantiles only
        rep = referee_elems("d/d(action) (tc=%d)" % tc, [dbc.cpu().numpy()], [act_o.grad.numpy()], [act64.grad.numpy()],
                            floors=(2e-5, 1e-4), big=5e-2, frac_slack=0.05)
        print("d/d(action) vs float64 (cuda, oracle32):", rep)

    # ---- eval mode (running statistics), as select_action uses it
    ora.eval()
    with torch.no_grad():
        z_e = _oracle_forward(ora, cloud, action)
    feat_e = engine.encoder_forward(ws, ef, geom, cl, 6, 4, bc, ctx, time=None, train=False)
    torch.cuda.synchronize()
    assert _rel(feat_e[:, :512], z_e) < 1e-4
    for k in sd_o:
        if "num_batches_tracked" in k:
            assert int(mine.state_dict()[k]) == 1, k
    lib.gaddpg_set_tensor_core(3)


def test_duplicate_folding_matches_dense_semantics(cuda):
    """A sparse cloud (few neighbours per ball) forces heavy padding: the folded rows + multiplicities must
    reproduce BatchNorm over the full (B, C, npoint, nsample) tensor the oracle materialises."""
    from gaddpg_b200 import engine, synthetic

    B, N = 8, 300
    ora, mine, ef = _build(4, 5, cuda)
    rs = np.random.RandomState(0)
    cloud = torch.from_numpy(synthetic.make_batch(B, N, step=1)["point_state_batch"]).clone()
    cloud[:, :3, 6:] = torch.from_numpy(rs.uniform(0.05, 0.6, (B, 3, N)).astype(np.float32))  # ~1-3 hits per ball
    ora.train()
    z_o = _oracle_forward(ora, cloud, None)
    ws = engine.Workspace(cuda)
    geom = engine.Geometry(B, N, cuda).build(cloud.to(cuda), 6)
    M1 = int(geom.lv[0].seg_off[-1])
    assert M1 < B * 32 * 8  # really folded
    ctx = engine.EncoderCtx(B, (geom.lv[0].cap, geom.lv[1].cap), engine.WIDTHS, cuda)
    feat = engine.encoder_forward(ws, ef, geom, cloud.to(cuda), 6, 4, None, ctx, train=True)
    assert _rel(feat[:, :512], z_o) < 1e-4


@pytest.mark.parametrize("B,N", [(24, 8192), (512, 2048), (1024, 1024)])
def test_encoder_forward_cfg5_corners(cuda, B, N):
    """BASELINE config 5's corners (N up to 8192 points, B up to 1024): value-encoder forward (cloud + action channels,
    train-mode BatchNorm) on the production kernels — the 8192-point FPS variant (xyz staged in 96 KB of shared memory),
    row counts up to ~1.7 M at SA1 — against the CPU oracle: features within 1e-4, running statistics within 1e-4,
    FPS / ball-query indices of the whole batch bit-exact."""
    import os

    from gaddpg_b200 import engine, synthetic
    from oracle.pointnet2_ops_cpu import pointnet2_utils as U

    torch.set_num_threads(os.cpu_count())
    ora, mine, ef = _build(10, 21, cuda)
    batch = synthetic.make_batch(B, N, step=3)
    cloud = torch.from_numpy(batch["point_state_batch"])
    action = torch.from_numpy(batch["action_batch"])
    ora.train()
    with torch.no_grad():
        z_o = _oracle_forward(ora, cloud, action)
    ws = engine.Workspace(cuda)
    cl = cloud.to(cuda)
    geom = engine.Geometry(B, N, cuda).build(cl, 6)
    ctx = engine.EncoderCtx(B, (geom.lv[0].cap, geom.lv[1].cap), engine.WIDTHS, cuda)
    feat = engine.encoder_forward(ws, ef, geom, cl, 6, 4, action.to(cuda).contiguous(), ctx, train=True)
    torch.cuda.synchronize()
    assert _rel(feat[:, :512], z_o) < 1e-4, (B, N)
    sd_o, sd_m = ora.state_dict(), mine.state_dict()
    for k in sd_o:
        if "running" in k:
            assert _rel(sd_m[k], sd_o[k]) < 1e-4, k
    xyz = cloud[:, :3, 6:].transpose(1, 2).contiguous()
    fps = U.fps_raw(xyz, 32)
    assert torch.equal(geom.lv[0].fps_idx.cpu(), fps), "FPS indices"
    ctr = torch.gather(xyz, 1, fps.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    assert torch.equal(geom.lv[0].bq_idx.cpu(), U.ball_query_raw(0.02, 64, xyz, ctr)), "SA1 ball-query indices"
    fps2 = U.fps_raw(ctr, 32)
    assert torch.equal(geom.lv[1].fps_idx.cpu(), fps2), "SA2 FPS indices"
    c2 = torch.gather(ctr, 1, fps2.long().unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    assert torch.equal(geom.lv[1].bq_idx.cpu(), U.ball_query_raw(0.04, 128, ctr, c2)), "SA2 ball-query indices"
