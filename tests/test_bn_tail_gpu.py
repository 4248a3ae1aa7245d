"""GPU: BatchNorm finalize fused into the kernel that produced the statistics (gaddpg_bn_tail, csrc/bn_tail.cuh) against the
separate finalize launches (gaddpg_bn_finalize_fwd / _bwd) — same arithmetic (FP64 slot sums in slot order), so the encoder
output, the running statistics and every parameter gradient must agree to rounding of the last FP32 operation (1e-6), for the
tcgen05 kernels (tc=3) and the FFMA fallbacks (tc=0: the library runs the tail as its own launch behind the product).
Replaces the batch-statistics half of torch.nn.BatchNorm2d / BatchNorm1d behind upstream build_shared_mlp and
/root/reference/core/networks.py:84-91."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(cuda, tail, tc, B, N):
    from gaddpg_b200 import engine, nets, synthetic
    from gaddpg_b200.capi import lib

    lib.gaddpg_set_tensor_core(tc)
    prev = engine.FUSED_BN_TAIL
    engine.FUSED_BN_TAIL = tail
    try:
        torch.manual_seed(5)
        mine = nets.make_encoder_params(10)
        with torch.no_grad():
            for m in mine.modules():
                if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
                    m.weight.uniform_(-1.0, 1.5)
                    m.bias.uniform_(-0.3, 0.3)
        ef = engine.EncoderFlat(mine, cuda)
        batch = synthetic.make_batch(B, N, step=9)
        cloud = torch.from_numpy(batch["point_state_batch"]).to(cuda)
        bc = torch.from_numpy(batch["action_batch"]).to(cuda).contiguous()
        R = torch.from_numpy(np.random.RandomState(4).randn(B, 516).astype(np.float32)).to(cuda)
        ws = engine.Workspace(cuda)
        geom = engine.Geometry(B, N, cuda).build(cloud, 6)
        caps = (geom.lv[0].cap, geom.lv[1].cap)
        ctx = engine.EncoderCtx(B, caps, engine.WIDTHS, cuda)
        sc = engine.BwdScratch(B, caps, engine.WIDTHS, cuda)
        n0 = lib.gaddpg_launch_count()
        for _ in range(2):   # twice: the ticket word must be back at zero after every launch
            feat = engine.encoder_forward(ws, ef, geom, cloud, 6, 4, bc, ctx, time=None, train=True).clone()
            dbc = engine.encoder_backward(ws, ef, ctx, sc, want_dw=True, want_dbc=True, dfeat=R).clone()
        torch.cuda.synchronize()
        launches = lib.gaddpg_launch_count() - n0
        grads = {k: p.grad.detach().clone() for k, p in mine.named_parameters() if p.grad is not None}
        bufs = {k: v.detach().clone() for k, v in mine.state_dict().items() if "running" in k or "num_batches" in k}
        return feat, dbc, grads, bufs, launches
    finally:
        engine.FUSED_BN_TAIL = prev
        lib.gaddpg_set_tensor_core(3)


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


@pytest.mark.parametrize("tc,B,N", [(3, 16, 1024), (3, 6, 256), (0, 6, 256)])
def test_fused_tail_equals_separate_finalize(cuda, tc, B, N):
    f0, d0, g0, b0, n0 = _run(cuda, False, tc, B, N)
    f1, d1, g1, b1, n1 = _run(cuda, True, tc, B, N)
    assert n1 < n0 or tc == 0          # fewer launches with the tails fused (the FFMA fallback launches the tail separately)
    assert _rel(f1, f0) < 2e-6 and _rel(d1, d0) < 1e-5
    for k in b0:
        if "num_batches" in k:
            assert int(b1[k]) == int(b0[k]) == 2, k
        else:
            assert _rel(b1[k], b0[k]) < 2e-6, k
    assert set(g0) == set(g1)
    for k in g0:
        assert _rel(g1[k], g0[k]) < 2e-5, (k, _rel(g1[k], g0[k]))
