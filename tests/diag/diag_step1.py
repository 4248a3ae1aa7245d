"""Diagnostic (GPU box): after ONE DDPG step from identical weights, compare gradients and parameter deltas
per tensor between the fused agent and the CPU oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from gaddpg_b200 import agent as ag, synthetic
from oracle.ddpg_cpu import OracleAgent

B, N = int(os.environ.get("B", 8)), int(os.environ.get("N", 512))
ora = OracleAgent("DDPG", seed=123456)
mine = ag.make_agent("DDPG", seed=123456)
mine.use_graph = False
init = {k: v.clone() for k, v in ora.feat.state_dict().items()}
batch = synthetic.make_batch(B, N, step=0)
u = np.random.RandomState(1).rand(B, 6).astype(np.float32)
o = ora.update_parameters(batch, noise_u=u)
m = mine.update_parameters(batch, 1, 0, noise_u=u)
print({k: (round(m[k], 6), round(o[k], 6)) for k in o})
for which, omod, mmod in (("policy_enc", ora.feat.encoder, mine._extractor.encoder), ("value_enc", ora.feat.value_encoder, mine._extractor.value_encoder)):
    print("==", which)
    for (k, po), (_, pm) in zip(omod.named_parameters(), mmod.named_parameters()):
        go, gm = po.grad.double(), pm.grad.cpu().double()
        gmax = go.abs().max().item()
        gerr = (go - gm).abs().max().item()
        dp_ = (po.detach().double() - pm.detach().cpu().double()).abs()
        flips = ((torch.sign(go) != torch.sign(gm)) & (go != 0)).double().mean().item()
        # gradient magnitude (relative to tensor max) at the entries whose parameters drifted the most
        idx = dp_.flatten().argsort(descending=True)[:5]
        rel_at = (go.flatten()[idx].abs() / (gmax + 1e-30)).tolist()
        print("%-28s gmax %.3e  gerr/gmax %.2e  signflip %.4f  max|dp| %.2e  |g|/gmax at worst dp: %s" % (
            k, gmax, gerr / (gmax + 1e-30), flips, dp_.max().item(), ["%.1e" % r for r in rel_at]))
