"""CPU: the torch-format optimiser state (de)serialisation of ga-ddpg_b200/checkpoint.py against a real
``torch.optim.Adam`` (layout, values, skipped parameters), and — when the reference tree is present — the oracle's
``save_model`` / ``load_model`` interchanging files with the UNMODIFIED reference (agent.py:282-431) in both
directions.  The fused agents are checked against the oracle's files on the GPU (tests/test_checkpoint_gpu.py)."""
import pytest
import torch
from torch import nn

from gaddpg_b200 import checkpoint, nets


def _toy():
    torch.manual_seed(0)
    mod = nn.Sequential(nn.Linear(5, 7), nn.ReLU(), nn.Linear(7, 3), nn.Linear(3, 2))  # last layer never gets a gradient
    order = [("%d.%s" % (i, n), p, False) for i, l in enumerate(mod) for n, p in l.named_parameters()]
    arena = nets.Arena(order, "cpu")
    return mod, arena


def _adam_steps(mod, arena, steps, lr=3e-4, eps=1e-5, wd=1e-5):
    """A real torch Adam run; afterwards its moments are mirrored into the arena (as the fused kernel would hold them)."""
    opt = torch.optim.Adam(mod.parameters(), lr=lr, eps=eps, weight_decay=wd)
    for s in range(steps):
        for p in mod.parameters():
            p.grad = None
        x = torch.randn(4, 5)
        mod[2](torch.relu(mod[0](x))).square().mean().backward()
        opt.step()
    mo = checkpoint.arena_moments(arena)
    for i, p in enumerate(mod.parameters()):
        st = opt.state.get(p)
        if st:
            m, v = mo(i, p)
            m.copy_(st["exp_avg"].flatten()), v.copy_(st["exp_avg_sq"].flatten())
    return opt


def _stepped_ranges(mod, arena):
    base = arena.p.data_ptr()
    return [((p.data_ptr() - base) // 4, p.numel()) for p in list(mod.parameters())[:4]]


def test_adam_state_dict_matches_torch_layout_and_values():
    mod, arena = _toy()
    opt = _adam_steps(mod, arena, 3)
    want = opt.state_dict()
    got = checkpoint.adam_state_dict(list(mod.parameters()), checkpoint.arena_moments(arena, _stepped_ranges(mod, arena)), 3,
                                     3e-4, 1e-5, 1e-5)
    assert sorted(got["state"].keys()) == sorted(want["state"].keys()) == [0, 1, 2, 3]   # params 4, 5 never stepped
    for i in want["state"]:
        assert set(got["state"][i]) == set(want["state"][i])
        assert float(got["state"][i]["step"]) == float(want["state"][i]["step"]) == 3.0
        assert torch.equal(got["state"][i]["exp_avg"], want["state"][i]["exp_avg"])
        assert torch.equal(got["state"][i]["exp_avg_sq"], want["state"][i]["exp_avg_sq"])
    gw, gg = want["param_groups"][0], got["param_groups"][0]
    for k in gw:
        assert gg[k] == gw[k], k
    # torch accepts it
    opt2 = torch.optim.Adam(mod.parameters(), lr=1.0)
    opt2.load_state_dict(got)
    assert opt2.param_groups[0]["lr"] == 3e-4 and opt2.param_groups[0]["eps"] == 1e-5


def test_fresh_optimizer_has_empty_state():
    mod, arena = _toy()
    got = checkpoint.adam_state_dict(list(mod.parameters()), checkpoint.arena_moments(arena), 0, 1e-3, 1e-8, 0.0)
    assert got["state"] == {} and got["param_groups"][0]["params"] == list(range(6))


@pytest.mark.parametrize("int_step", [False, True])
def test_load_adam_state_dict_scatters_into_arena(int_step):
    mod, arena = _toy()
    opt = _adam_steps(mod, arena, 2)
    sd = opt.state_dict()
    if int_step:  # torch < 1.12 stored a python int
        for st in sd["state"].values():
            st["step"] = int(float(st["step"]))
    sd["param_groups"][0]["lr"] = 7.5e-5
    mod2, arena2 = _toy()
    arena2.m.fill_(9.0), arena2.v.fill_(9.0)
    step, lr = checkpoint.load_adam_state_dict(sd, list(mod2.parameters()), checkpoint.arena_moments(arena2))
    assert step == 2 and lr == 7.5e-5
    mo, mo2 = checkpoint.arena_moments(arena), checkpoint.arena_moments(arena2)
    for i, (p, q) in enumerate(zip(mod.parameters(), mod2.parameters())):
        (m, v), (m2, v2) = mo(i, p), mo2(i, q)
        if i < 4:
            assert torch.equal(m, m2) and torch.equal(v, v2)
        else:  # no state in the file: zero moments, what torch would lazily create
            assert float(m2.abs().sum()) == 0.0 and float(v2.abs().sum()) == 0.0


def test_load_rejects_mismatched_state():
    mod, arena = _toy()
    sd = _adam_steps(mod, arena, 1).state_dict()
    other = nn.Sequential(nn.Linear(5, 8), nn.ReLU(), nn.Linear(8, 3), nn.Linear(3, 2))
    arena_o = nets.Arena([("%d.%s" % (i, n), p, False) for i, l in enumerate(other) for n, p in l.named_parameters()], "cpu")
    with pytest.raises(ValueError):
        checkpoint.load_adam_state_dict(sd, list(other.parameters()), checkpoint.arena_moments(arena_o))
    with pytest.raises(ValueError):
        checkpoint.load_adam_state_dict(sd, list(mod.parameters())[:3], checkpoint.arena_moments(arena))


def test_paths_follow_reference_naming():
    p = checkpoint.paths("out", "DDPG", "PandaYCBEnv", "latest")
    assert p["actor"] == "out/DDPG_actor_PandaYCBEnv_latest" and p["critic"] == "out/DDPG_critic_PandaYCBEnv_latest"
    assert p["state_feat"] == "out/DDPG_state_feat_PandaYCBEnv_latest"


def test_oracle_checkpoints_interchange_with_reference_when_present():
    from oracle import refstack

    if not refstack.available():
        pytest.skip("/root/reference not present (GPU box)")
    from oracle import make_golden

    make_golden.checkpoint_roundtrip("DDPG")


def test_checkpoint_load_is_restricted_by_default(tmp_path):
    """Files in the reference's layout (tensors, numbers, the scheduler's Counter) load with the restricted unpickler;
    a file that needs arbitrary pickle is refused unless the caller declares it trusted."""
    import collections

    from torch.optim.lr_scheduler import MultiStepLR

    mod, arena = _toy()
    opt = _adam_steps(mod, arena, 1)
    sch = MultiStepLR(opt, milestones=[5, 9], gamma=0.5)
    good = str(tmp_path / "good")
    checkpoint.save({"net": mod.state_dict(), "opt": opt.state_dict(), "sch": sch.state_dict(), "step": 3}, good)
    d = checkpoint.load(good)
    assert d["step"] == 3 and isinstance(d["sch"]["milestones"], collections.Counter) and torch.equal(d["net"]["0.weight"], mod[0].weight)

    class Evil:
        def __reduce__(self):
            return (print, ("arbitrary code ran",))

    bad = str(tmp_path / "bad")
    torch.save({"net": Evil()}, bad)
    with pytest.raises(RuntimeError, match="trusted"):
        checkpoint.load(bad)
    assert "net" in checkpoint.load(bad, trusted=True)
