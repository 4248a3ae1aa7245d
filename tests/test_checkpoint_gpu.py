"""GPU: reference-format checkpoints of the fused agents (agent.py:282-431).

The oracle's ``save_model`` / ``load_model`` interchange files with the unmodified reference in both directions
(pinned in tests/test_checkpoint_cpu.py / oracle/make_golden.py::checkpoint_roundtrip), so:
  oracle file -> DDPGB200.load_model -> one more step on both sides: all scalars within 1e-4 (identical weights,
  Adam moments, step counters and learning rates must all have travelled through the file);
  DDPGB200.save_model -> oracle.load_model (torch's own ``Adam.load_state_dict`` / ``MultiStepLR.load_state_dict``
  accept the files) -> exact equality of every tensor with the fused arenas."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _close(a, b, rtol=1e-4, atol=1e-6):
    return (np.isnan(a) and np.isnan(b)) or abs(a - b) <= atol + rtol * abs(b)


@pytest.mark.parametrize("policy", ["DDPG", "BC"])
def test_load_reference_format_checkpoint_then_step(cuda, tmp_path, policy):
    from gaddpg_b200 import agent as ag, synthetic
    from gaddpg_b200.config import LOSS_KEYS
    from oracle.ddpg_cpu import OracleAgent

    B, N = 8, 512
    ora = OracleAgent(policy, seed=123456)
    rs = np.random.RandomState(3)
    for step in range(3):           # odd, even, odd: every optimiser has state, Adam step counters = 3
        ora.update_parameters(synthetic.make_batch(B, N, step=step), noise_u=rs.rand(B, 6).astype(np.float32))
        ora.step_scheduler()
    ora.save_model(ora.update_step, str(tmp_path), surfix="t")
    ora.load_model(str(tmp_path), surfix="t")          # the reference hard-updates the targets on load: same on both sides
    mine = ag.make_agent(policy, seed=99)              # different initial weights: everything must come from the files
    assert mine.load_model(str(tmp_path), surfix="t") == ora.update_step == mine.update_step
    so, sm = ora.state_dicts(), mine.state_dicts()
    for net in so:
        for k, v in so[net].items():
            assert torch.equal(sm[net][k].cpu(), v), (net, k)
    want_steps = dict(policy=3, enc=3, critic=3, venc=3) if policy == "DDPG" else dict(policy=3, enc=3, critic=0, venc=0)
    assert mine.opt_steps == want_steps, mine.opt_steps
    assert abs(mine.get_lr()["policy_lr"] - ora.policy_sched.get_last_lr()[0]) < 1e-12
    for step in (3, 4):
        batch = synthetic.make_batch(B, N, step=step)
        u = rs.rand(B, 6).astype(np.float32)
        o = ora.update_parameters(batch, noise_u=u)
        m = mine.update_parameters(batch, mine.update_step, 0, noise_u=u)
        if step == 3:   # identical state on both sides: the 1e-4 bar (later steps free-run, see test_agent_gpu.py)
            for k in LOSS_KEYS:
                tol = {"actor_critic_loss": 3e-3, "critic_grad": 3e-2, "policy_param": 5e-4, "critic_param": 5e-4}.get(k, 1e-4)
                assert _close(m[k], o[k], rtol=tol), (policy, k, m[k], o[k])
            # the first Adam step after loading used the restored moments: heads stay within a fraction of an lr unit
            for net in ("policy",) + (("critic",) if policy == "DDPG" else ()):
                d = max(float((mine.state_dicts()[net][k].cpu() - v).abs().max()) for k, v in ora.state_dicts()[net].items())
                assert d < 6e-4, (net, d)


def test_saved_checkpoint_loads_into_torch_optimizers(cuda, tmp_path):
    from gaddpg_b200 import agent as ag, checkpoint, synthetic
    from oracle.ddpg_cpu import OracleAgent

    B, N = 8, 512
    mine = ag.make_agent("DDPG", seed=123456)
    rs = np.random.RandomState(5)
    for step in range(2):
        mine.update_parameters(synthetic.make_batch(B, N, step=step), mine.update_step, 0, noise_u=rs.rand(B, 6).astype(np.float32))
        mine.step_scheduler(mine.update_step)
    files = mine.save_model(mine.update_step, str(tmp_path), surfix="m")
    # file / key layout of agent.py:312-352
    a, c, f = (checkpoint.load(files[k]) for k in ("actor", "critic", "state_feat"))
    assert set(a) == {"net", "opt", "sch"} and set(c) == {"net", "opt", "sch"}
    assert set(f) == {"net", "opt", "encoder_opt", "sch", "encoder_sch", "val_encoder_opt", "val_encoder_sch", "step"}
    assert all(k.startswith("module.") for k in f["net"]) and f["step"] == 3
    assert f["opt"]["state"] == {}                                # whole-extractor optimiser is never stepped
    assert f["val_encoder_sch"]["last_epoch"] == 0 and f["encoder_sch"]["last_epoch"] == 2   # never-stepped scheduler
    n_pol = len(list(mine.policy.parameters()))
    assert sorted(a["opt"]["state"]) == list(range(n_pol - 2))    # log_std_linear never receives a gradient
    ora = OracleAgent("DDPG", seed=7)
    assert ora.load_model(str(tmp_path), surfix="m") == 3         # torch's own load_state_dict accepts every dict
    for opt, mod, arena in ((ora.policy_opt, ora.policy, mine.pf.arena), (ora.critic_opt, ora.critic, mine.cf.arena),
                            (ora.enc_opt, ora.feat.encoder, mine.ef_p.arena), (ora.venc_opt, ora.feat.value_encoder, mine.ef_v.arena)):
        mmod = {id(ora.policy): mine.policy, id(ora.critic): mine.critic, id(ora.feat.encoder): mine._extractor.encoder,
                id(ora.feat.value_encoder): mine._extractor.value_encoder}[id(mod)]
        mo = checkpoint.arena_moments(arena)
        for i, (po, pm) in enumerate(zip(mod.parameters(), mmod.parameters())):
            assert torch.equal(po.detach(), pm.detach().cpu())
            st = opt.state.get(po)
            if st:
                m, v = mo(i, pm)
                assert float(st["step"]) == 2.0
                assert torch.equal(st["exp_avg"].flatten(), m.cpu()) and torch.equal(st["exp_avg_sq"].flatten(), v.cpu())
    assert ora.policy_opt.param_groups[0]["eps"] == 1e-5 and ora.policy_opt.param_groups[0]["weight_decay"] == 1e-5
    # and back into a second fused agent: bit-identical continuation
    twin = ag.make_agent("DDPG", seed=11)
    twin.load_model(str(tmp_path), surfix="m")
    mine.load_model(str(tmp_path), surfix="m")    # hard-updates the targets, as the reference's load does
    batch, u = synthetic.make_batch(B, N, step=7), rs.rand(B, 6).astype(np.float32)
    r1 = mine.update_parameters(batch, mine.update_step, 0, noise_u=u)
    r2 = twin.update_parameters(batch, twin.update_step, 0, noise_u=u)
    assert r1 == r2 or all((np.isnan(r1[k]) and np.isnan(r2[k])) or r1[k] == r2[k] for k in r1)
