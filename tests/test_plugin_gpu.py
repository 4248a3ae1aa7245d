"""GPU: the drop-in plugin surface.  ``PointNetFeatureB200`` under autograd (the path that lets the reference's own
Agent/DDPG classes run unmodified) against the oracle's PointNetFeature restatement: same state_dict
keys, same outputs, gradients w.r.t. the parameters and w.r.t. the action that ``concat_state_action_channelwise``
broadcasts into the cloud (utils.py:291-297); and the ``pointnet2_ops``-style ops module."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _grads_close(named_o, named_m):
    """Kink-tolerant gradient comparison: two FP32 evaluations route a few ReLU / max-pool elements
    differently, which moves whole tensors by up to a few percent; the concatenated gradient must still agree in L2."""
    num = den = 0.0
    worst = 0.0
    for (k, po), (_, pm) in zip(named_o, named_m):
        if k.endswith(("1.0.bias", "1.3.bias")):   # exact-zero true gradient (bias in front of BatchNorm1d)
            continue
        a, b = pm.grad.detach().cpu().double(), po.grad.detach().cpu().double()
        num += float(((a - b) ** 2).sum())
        den += float((b ** 2).sum())
        worst = max(worst, float((a - b).abs().max() / (b.abs().max() + 1e-30)))
    l2 = (num / den) ** 0.5
    assert l2 < 5e-2 and worst < 0.25, (l2, worst)
    return l2


def _rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def test_plugin_forward_backward_like_reference_extractor(cuda):
    from gaddpg_b200 import synthetic
    from gaddpg_b200.networks import PointNetFeatureB200
    from oracle.nets_cpu import PointFeature

    B, N = 8, 512
    torch.manual_seed(3)
    ora = PointFeature(extra_latent=1, action_concat=True)
    torch.manual_seed(3)
    net = PointNetFeatureB200(input_dim=5, extra_latent=1, action_concat=True)  # kwargs as utils.py:190-198 passes them
    assert list(net.state_dict().keys()) == list(ora.state_dict().keys())
    for k, v in ora.state_dict().items():
        assert torch.equal(net.state_dict()[k], v), k                              # same constructors, same seed
    net = torch.nn.DataParallel(net).to("cuda")                                   # utils.py:202-204
    batch = synthetic.make_batch(B, N, step=2)
    cloud = torch.from_numpy(batch["point_state_batch"])
    action = torch.from_numpy(batch["action_batch"])
    R = torch.from_numpy(np.random.RandomState(0).randn(B, 512).astype(np.float32))

    # policy path
    ora.train(), net.train()
    z_o = ora(cloud, value=False)
    z_m, passthrough = net(cloud.to(cuda), feature_2=False)
    assert passthrough.shape == cloud.shape and _rel(z_m, z_o) < 1e-4
    (z_o * R).sum().backward()
    (z_m * R.to(cuda)).sum().backward()
    _grads_close(ora.encoder.named_parameters(), net.module.encoder.named_parameters())

    # value path: the agent concatenates the (grad-carrying) action over the points, ddpg.py:45-46
    a_o = action.clone().requires_grad_(True)
    a_m = action.clone().to(cuda).requires_grad_(True)
    pc_o = torch.cat((cloud, a_o.unsqueeze(2).expand(-1, -1, cloud.shape[2])), 1)
    pc_m = torch.cat((cloud.to(cuda), a_m.unsqueeze(2).expand(-1, -1, cloud.shape[2])), 1)
    v_o = ora(pc_o, value=True)
    v_m, _ = net(pc_m, feature_2=True)
    assert _rel(v_m, v_o) < 1e-4
    (v_o * R).sum().backward()
    (v_m * R.to(cuda)).sum().backward()
    assert _rel(a_m.grad, a_o.grad) < 0.25
    _grads_close(ora.value_encoder.named_parameters(), net.module.value_encoder.named_parameters())
    # float64 referee for the two kink-tolerant bounds above (B = 8: BatchNorm1d over 8 samples amplifies every flip):
    # the autograd plugin's gradients are as close to the float64 gradients as the fp32 oracle's are
    import copy

    from tests.f64ref import f64_ops, referee_elems

    o64 = copy.deepcopy(ora).double()
    o64.zero_grad()
    a64 = action.double().clone().requires_grad_(True)
    with f64_ops():
        v64 = o64(torch.cat((cloud.double(), a64.unsqueeze(2).expand(-1, -1, cloud.shape[2])), 1), value=True)
        (v64 * R.double()).sum().backward()
    rep = referee_elems("plugin d/d(action)", [a_m.grad.cpu().numpy()], [a_o.grad.numpy()], [a64.grad.numpy()],
                        floors=(2e-5, 1e-4), big=5e-2, frac_slack=0.05)
    print("plugin d/d(action) vs float64 (cuda, oracle32):", rep)
    keys = [k for k, _ in ora.value_encoder.named_parameters() if not k.endswith(("1.0.bias", "1.3.bias"))]
    po, pm, p64 = dict(ora.value_encoder.named_parameters()), dict(net.module.value_encoder.named_parameters()), dict(o64.value_encoder.named_parameters())
    rep = referee_elems("plugin value-encoder gradients", [pm[k].grad.cpu().numpy() for k in keys], [po[k].grad.numpy() for k in keys],
                        [p64[k].grad.numpy() for k in keys])
    print("plugin value-encoder grads vs float64 (cuda, oracle32):", rep)

    # a torch optimiser may step the parameters between calls (the reference's Adam instances do): views stay valid
    opt = torch.optim.Adam(net.module.encoder.parameters(), lr=1e-3)
    opt.step()
    z2, _ = net(cloud.to(cuda), feature_2=False)
    assert torch.isfinite(z2).all() and not torch.equal(z2, z_m)

    # eval mode = running statistics
    ora.eval(), net.eval()
    with torch.no_grad():
        assert _rel(net(pc_m, feature_2=True)[0], ora(pc_o, value=True)) < 1e-4


def test_ops_module_regularize_pc_point_count_usage(cuda):
    """core/utils.py:795-796: gather_operation(pc^T, furthest_point_sample(pc[..., :3], npoints))."""
    from gaddpg_b200 import ops
    from oracle.pointnet2_ops_cpu import pointnet2_utils as U

    pc = torch.from_numpy(np.random.RandomState(1).uniform(-0.2, 0.5, (1, 3000, 4)).astype(np.float32))
    want = U.gather_operation(pc.transpose(1, 2).contiguous(), U.furthest_point_sample(pc[..., :3].contiguous(), 1024))
    got = ops.gather_operation(pc.to(cuda).transpose(1, 2).contiguous(), ops.furthest_point_sample(pc.to(cuda)[..., :3].contiguous(), 1024))
    assert torch.equal(got.cpu(), want)
