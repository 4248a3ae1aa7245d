"""Diagnostic (GPU): after each teacher-forced step of test_steps_from_synchronised_state, the max |cuda - oracle| of the
feature-extractor tensors split into the null-gradient tensors (Linear biases in front of BatchNorm1d: '1.0.bias', '1.3.bias')
and everything else — the data behind the two bounds in tests/test_agent_gpu.py."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from tests.test_agent_gpu import _sync_from_oracle
from gaddpg_b200 import agent as ag, synthetic
from oracle.ddpg_cpu import OracleAgent

for policy, over in (("DDPG", {}), ("DDPG", dict(policy_aux=False, critic_aux=False)), ("BC", {})):
    B, N = 8, 512
    ora = OracleAgent(policy, seed=123456, **over)
    mine = ag.make_agent(policy, seed=123456, **over)
    rs = np.random.RandomState(9)
    for step in range(4):
        _sync_from_oracle(mine, ora)
        batch = synthetic.make_batch(B, N, step=step)
        u = rs.rand(B, 6).astype(np.float32)
        ora.update_parameters(batch, noise_u=u); ora.step_scheduler()
        mine.update_parameters(batch, mine.update_step, 0, noise_u=u); mine.step_scheduler(mine.update_step)
        sa, so = mine.state_dicts()["state_feat"], ora.state_dicts()["state_feat"]
        null = other = run = 0.0
        worst_other = None
        for k in so:
            if "num_batches" in k:
                continue
            d = float((sa[k].detach().cpu().double() - so[k].detach().cpu().double()).abs().max())
            if k.endswith(("1.0.bias", "1.3.bias")):
                null = max(null, d)
            elif "running" in k:
                run = max(run, d)
            else:
                if d > other:
                    other, worst_other = d, k
        print("%s %s step %d: null %.2e  other params %.2e (%s)  running stats %.2e" % (policy, "aux" if not over else "noaux", step, null, other, worst_other, run))
