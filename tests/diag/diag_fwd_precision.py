"""Diagnostic (GPU box): forward accuracy of the fp32 CPU oracle and of the CUDA encoder against a float64 evaluation of the
same network on the same inputs (value encoder, DDPG init weights, B/N from env)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import copy
import numpy as np, torch
from gaddpg_b200 import agent as ag, synthetic, engine
from oracle.ddpg_cpu import OracleAgent
import oracle.pointnet2_ops_cpu.pointnet2_utils as U

B, N = int(os.environ.get("B", 8)), int(os.environ.get("N", 512))
ora = OracleAgent("DDPG", seed=123456)
mine = ag.make_agent("DDPG", seed=123456)
batch = synthetic.make_batch(B, N, step=0)
mine.prepare_data(batch, np.zeros((B, 6), np.float32))
d = ora._load(batch)
o64 = copy.deepcopy(ora); o64.feat.double()
_fps, _bq, _grp, _gat = U.fps_raw, U.ball_query_raw, U.grouping_operation, U.gather_operation
def grouping64(features, idx):
    B_, C_, N_ = features.shape
    ii = idx.long().view(B_, 1, -1).expand(-1, C_, -1)
    return torch.gather(features, 2, ii).view(B_, C_, idx.shape[1], idx.shape[2])
def gather64(features, idx):
    return torch.gather(features, 2, idx.long().unsqueeze(1).expand(-1, features.shape[1], -1))
U.grouping_operation, U.gather_operation = grouping64, gather64
U.furthest_point_sample = lambda xyz, n: _fps(xyz.float(), n)
U.ball_query = lambda r, ns, xyz, new: _bq(r, ns, xyz.float().contiguous(), new.float().contiguous())
def hooks(agent, store):
    hs = []
    for i in range(3):
        sq = agent.feat.value_encoder[0][i].mlps[0]
        for l in range(3):
            hs.append(sq[3 * l].register_forward_hook(lambda m, a, out, k="sa%d.%d" % (i, l): store.__setitem__(k, out.detach())))
    fc = agent.feat.value_encoder[1]
    for j, li in enumerate((0, 3)):
        hs.append(fc[li].register_forward_hook(lambda m, a, out, k="fc%d" % j: store.__setitem__(k, out.detach())))
    return hs
y32, y64 = {}, {}
hooks(ora, y32); hooks(o64, y64)
ora.feat.train(); o64.feat.train()
with torch.no_grad():
    pc = torch.cat((d["cloud"], d["action"].unsqueeze(2).expand(-1, -1, d["cloud"].shape[2])), 1)
    z32 = ora.feat(pc, value=True)
    z64 = o64.feat(pc.double(), value=True)
mine.geom_s.build(mine.cloud, mine.skip)
ctx = mine.ctx_v1
feat = engine.encoder_forward(mine.ws, mine.ef_v, mine.geom_s, mine.cloud, mine.skip, mine.Cp_value, mine.v.action, ctx, time=mine.v.time, train=True)
torch.cuda.synchronize()
rel = lambda a, b: float((a.double().cpu() - b).abs().max() / b.abs().max())
print("B=%d N=%d" % (B, N))
for k in sorted(y32):
    line = "%-6s fp32-oracle vs fp64: %.2e" % (k, rel(y32[k], y64[k]))
    if k.startswith("sa2."):
        l = int(k[-1]); Yo = y64[k].squeeze(2).transpose(1, 2).reshape(B * 32, -1)
        line += "   CUDA vs fp64: %.2e" % rel(ctx.sa[2].Y[l], Yo)
    if k.startswith("fc"):
        line += "   CUDA vs fp64: %.2e" % rel(ctx.fc.Y[int(k[-1])], y64[k])
    print(line)
print("z      fp32-oracle vs fp64: %.2e   CUDA vs fp64: %.2e" % (rel(z32, z64), rel(feat[:, :512], z64)))
