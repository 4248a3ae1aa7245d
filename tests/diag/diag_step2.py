"""Diagnostic (GPU box): even step (actor-critic branch) from synchronised state: compare gradients per tensor."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from gaddpg_b200 import agent as ag, synthetic
from oracle.ddpg_cpu import OracleAgent
from tests.test_agent_gpu import _sync_from_oracle

B, N = int(os.environ.get("B", 8)), int(os.environ.get("N", 512))
ora = OracleAgent("DDPG", seed=123456)
mine = ag.make_agent("DDPG", seed=123456)
mine.use_graph = False
rs = np.random.RandomState(9)
for step in range(2):
    _sync_from_oracle(mine, ora)
    batch = synthetic.make_batch(B, N, step=step)
    u = rs.rand(B, 6).astype(np.float32)
    o = ora.update_parameters(batch, noise_u=u)
    m = mine.update_parameters(batch, mine.update_step, 0, noise_u=u)
    ora.step_scheduler(); mine.step_scheduler()
    print(step, {k: (round(m[k], 6), round(o[k], 6)) for k in o if o[k]})
print("pi rel", float((mine.pc.pi.cpu() - ora.last["pi"]).abs().max() / ora.last["pi"].abs().max()))
for which, omod, mmod in (("policy", ora.policy, mine.policy), ("policy_enc", ora.feat.encoder, mine._extractor.encoder), ("critic", ora.critic, mine.critic)):
    print("==", which)
    for (k, po), (_, pm) in zip(omod.named_parameters(), mmod.named_parameters()):
        if po.grad is None:
            print("%-28s oracle grad None; mine |g|max %.2e" % (k, pm.grad.abs().max().item())); continue
        go, gm = po.grad.double(), pm.grad.cpu().double()
        gmax = go.abs().max().item()
        gerr = (go - gm).abs().max().item()
        dp_ = (po.detach().double() - pm.detach().cpu().double()).abs()
        print("%-28s gmax %.3e  gerr/gmax %.2e  max|dp| %.2e" % (k, gmax, gerr / (gmax + 1e-30), dp_.max().item()))

# ---- is the even-step actor gradient mismatch explained by the (legitimately different) null-direction Adam updates of
# the value encoder between the critic step and the actor forward?  Re-evaluate the ORACLE's actor gradient with the
# fused agent's post-update value encoder + critic and the synchronised pre-step policy side.
import copy
from oracle import losses_cpu as L
ora = OracleAgent("DDPG", seed=123456)
mine = ag.make_agent("DDPG", seed=123456)
mine.use_graph = False
rs = np.random.RandomState(9)
for step in range(2):
    _sync_from_oracle(mine, ora)
    pre = copy.deepcopy(ora.state_dicts())
    batch = synthetic.make_batch(B, N, step=step)
    u = rs.rand(B, 6).astype(np.float32)
    o = ora.update_parameters(batch, noise_u=u)
    m = mine.update_parameters(batch, mine.update_step, 0, noise_u=u)
    ora.step_scheduler(); mine.step_scheduler()
post = mine.state_dicts()
hyb = copy.deepcopy(pre)
for k in post["state_feat"]:
    if "value_encoder" in k:
        hyb["state_feat"][k] = post["state_feat"][k].cpu().clone()
hyb["critic"] = {k: v.cpu().clone() for k, v in post["critic"].items()}
o2 = OracleAgent("DDPG", seed=1)
o2.load_state_dicts(hyb)
o2.update_step = 2
o2.feat.train(); o2.policy.train(); o2.critic.train()
d = o2._load(batch)
mix = o2._mix_policy_ratio()
pf = o2.features(d["cloud"], d["time"], value=False)
pi, _, _, aux = o2.policy.sample(pf)
pi.retain_grad()
vpf = o2.features(d["cloud"], d["time"], pi, value=True)
vpf.retain_grad()
q1p, q2p, _ = o2.critic(vpf)
sel = ~d["expert_reward_mask"]
ac = -mix * torch.min(q1p.squeeze()[sel], q2p.squeeze()[sel]).mean()
gm = d["reward_mask"]; em = d["expert_mask"]
loss = L.pose_bc_loss(pi[em], d["expert_action"][em]) * (1 - mix) + L.goal_pred_loss(aux[gm, :7], d["goal"][gm, :7]) + ac
g_ac = torch.autograd.grad(ac, pi, retain_graph=True)[0]
print("dpi_ac oracle vs mine:\n", g_ac[:3], "\n", mine.dpi_ac[:3].cpu())
print("rel err dpi_ac", float((g_ac - mine.dpi_ac.cpu()).abs().max() / g_ac.abs().max()))
loss.backward()
print("== hybrid oracle (fused agent's post-update value encoder/critic) vs fused agent: policy grads; ac", float(ac), m["actor_critic_loss"])
for (k, po), (_, pm) in zip(o2.policy.named_parameters(), mine.policy.named_parameters()):
    if po.grad is None: continue
    go, gmn = po.grad.double(), pm.grad.cpu().double()
    print("%-28s gmax %.3e  gerr/gmax %.2e" % (k, go.abs().max().item(), (go - gmn).abs().max().item() / go.abs().max().item()))
for (k, po), (_, pm) in list(zip(o2.feat.encoder.named_parameters(), mine._extractor.encoder.named_parameters()))[:6]:
    go, gmn = po.grad.double(), pm.grad.cpu().double()
    print("%-28s gmax %.3e  gerr/gmax %.2e" % (k, go.abs().max().item(), (go - gmn).abs().max().item() / go.abs().max().item()))

# ---- isolate: feed the oracle's d(ac)/d(value_pi_feat) into the CUDA encoder backward (unit-test entry path)
from gaddpg_b200 import engine
dfeat_o = vpf.grad.clone()  # (B,513) gradient of the total loss w.r.t. value_pi_feat == gradient of ac (only ac uses it)
ws, ef, geom = mine.ws, mine.ef_v, mine.geom_s
ctx = mine.ctx_v5
# re-run F5 forward with the fused agent's current weights and ITS pi (matches oracle pi to ~1e-6)
feat = engine.encoder_forward(ws, ef, geom, mine.cloud, mine.skip, mine.Cp_value, mine.pc.pi, ctx, time=mine.v.time, train=True)
print("F5 feat rel err", float((feat[:, :513].cpu() - vpf.detach()).abs().max() / vpf.abs().max()))
dfe = torch.zeros(B, 516, device="cuda"); dfe[:, :513] = dfeat_o.cuda()
for want_dw in (True, False):
    dbc = engine.encoder_backward(ws, ef, ctx, mine.sc, want_dw=want_dw, want_dbc=True, accumulate=0, dfeat=dfe)
    print("want_dw", want_dw, "rel err dbc (oracle dfeat -> CUDA encoder bwd)", float((g_ac - dbc.cpu()).abs().max() / g_ac.abs().max()))
# and the critic part: CUDA critic backward from the oracle's dq
qa = engine.critic_forward(mine.cf, feat, mine.cc5, B, nb=2)
print("q1 rel err", float((qa[:, 0].cpu() - q1p.detach().squeeze()).abs().max() / q1p.abs().max()))
print("g_ac[0]", g_ac[0].tolist()); print("dbc[0]", dbc[0].cpu().tolist()); print("dpi_ac[0]", mine.dpi_ac[0].cpu().tolist())
print("dfeat_o stats", float(dfeat_o.abs().max()), float(dfeat_o[:, 512].abs().max()), dfeat_o.shape)
# oracle: gradient w.r.t. action of sum(vpf * dfeat_o) recomputed from scratch (sanity of the entry point)
a2 = pi.detach().clone().requires_grad_(True)
v2 = o2.features(d["cloud"], d["time"], a2, value=True)
g2 = torch.autograd.grad((v2 * dfeat_o).sum(), a2)[0]
print("oracle recomputed via dfeat entry: rel to g_ac", float((g2 - g_ac).abs().max() / g_ac.abs().max()))
print("CUDA dbc vs g2", float((g2 - dbc.cpu()).abs().max() / g2.abs().max()))

# ---- float64 ground truth of d(ac)/d(pi) with the same weights
import oracle.pointnet2_ops_cpu.pointnet2_utils as U64
o3 = OracleAgent("DDPG", seed=1)
o3.load_state_dicts(hyb)
o3.feat.train(); o3.policy.train(); o3.critic.train()
# run the value path in float64: torch modules .double(); the C index ops need float32 xyz -> compute indices in f32 (same as everyone), features in f64
o3.feat.double(); o3.critic.double()
_orig_fps, _orig_bq = U64.fps_raw, U64.ball_query_raw
cl64 = d["cloud"].double()
a64 = pi.detach().double().clone().requires_grad_(True)
# patch the ctypes-backed ops to accept float64 by casting the geometry to float32 and the gathered features stay float64 through torch indexing
def grouping64(features, idx):
    B_, C_, N_ = features.shape
    ii = idx.long().view(B_, 1, -1).expand(-1, C_, -1)
    return torch.gather(features, 2, ii).view(B_, C_, idx.shape[1], idx.shape[2])
def gather64(features, idx):
    ii = idx.long().unsqueeze(1).expand(-1, features.shape[1], -1)
    return torch.gather(features, 2, ii)
U64.grouping_operation = grouping64
U64.gather_operation = gather64
U64.furthest_point_sample = lambda xyz, n: _orig_fps(xyz.float(), n)
U64.ball_query = lambda r, ns, xyz, new: _orig_bq(r, ns, xyz.float().contiguous(), new.float().contiguous())
pc64 = torch.cat((cl64, a64.unsqueeze(2).expand(-1, -1, cl64.shape[2])), 1)
z64 = o3.feat(pc64, value=True)
v64 = torch.cat((z64, d["time"].double()[:, None]), 1)
q1_, q2_, _ = o3.critic(v64)
ac64 = -mix * torch.min(q1_.squeeze()[sel], q2_.squeeze()[sel]).mean()
g64 = torch.autograd.grad(ac64, a64)[0]
den = g64.abs().max()
print("fp64 truth: ac", float(ac64))
print("oracle fp32 vs fp64:", float((g_ac.double() - g64).abs().max() / den))
print("CUDA   fp32 vs fp64:", float((mine.dpi_ac.cpu().double() - g64).abs().max() / den))

# ---- where does the per-sample systematic error enter?  Compare per-sample column sums of dL/d(conv output) for the
# three SA1 layers: fp64 oracle (hooks) vs values reconstructed from the CUDA buffers of the last encoder_backward.
grads64 = {}
def mk(name):
    def hook(mod, inp, out):
        out.retain_grad(); grads64[name] = out
    return hook
hs = []
seq = o3.feat.value_encoder[0][0].mlps[0]
for l in range(3):
    hs.append(seq[3 * l].register_forward_hook(mk("sa1.%d" % l)))
a64 = pi.detach().double().clone().requires_grad_(True)
pc64 = torch.cat((cl64, a64.unsqueeze(2).expand(-1, -1, cl64.shape[2])), 1)
z64 = o3.feat(pc64, value=True)
v64 = torch.cat((z64, d["time"].double()[:, None]), 1)
q1_, q2_, _ = o3.critic(v64)
ac64 = -mix * torch.min(q1_.squeeze()[sel], q2_.squeeze()[sel]).mean()
ac64.backward()
# CUDA side: rerun fused F5 + its backward exactly as the agent does
feat = engine.encoder_forward(ws, ef, geom, mine.cloud, mine.skip, mine.Cp_value, mine.pc.pi, ctx, time=mine.v.time, train=True)
qa = engine.critic_forward(mine.cf, feat, mine.cc5, B, nb=2)
from gaddpg_b200.capi import lib, current_stream
from gaddpg_b200.engine import dp, QA_LD, QA_Q2
lib.gaddpg_actor_critic_loss(dp(qa), QA_LD, QA_Q2, dp(mine.v.ret), dp(mine.v.expert_flag), float(mix), B, 1.0, QA_LD, dp(mine.cc5.dqa), mine.out.data_ptr() + 4 * 6, current_stream())
engine.critic_backward(ws, mine.cf, feat, mine.cc5, B, 2, ctx, mine.sc, accumulate=1)
dbc = engine.encoder_backward(ws, ef, ctx, mine.sc, want_dw=False, want_dbc=True)
torch.cuda.synchronize()
l1 = geom.lv[0]
M = int(l1.seg_off[-1]); seg_off = l1.seg_off.cpu().numpy(); rw = l1.row_w[:M].double()
s = ctx.sa[0]
for l in range(3):
    D, Y, bn, bb = mine.sc.D[0][l][:M].double(), s.Y[l][:M].double(), s.bn[l], mine.sc.bb[0][l]
    dY = bb.g.double() * (D - rw[:, None] * (bb.m1.double() + (Y - bn.mean.double()) * bn.rstd.double() * bb.m2.double()))
    mine_sum = torch.stack([dY[seg_off[b * 32]: seg_off[(b + 1) * 32]].sum(0) for b in range(B)]).cpu()
    G = grads64["sa1.%d" % l].grad  # (B, C, 32, 64)
    tru = G.sum((2, 3))
    print("SA1 layer %d per-sample column sums: rel err %.3e   (|truth|max %.3e, sum over batch truth %.2e mine %.2e)" % (
        l, float((mine_sum - tru).abs().max() / tru.abs().max()), float(tru.abs().max()), float(tru.sum(0).abs().max()), float(mine_sum.sum(0).abs().max())))
    # same comparison with the CUDA D/Y but float64 statistics (isolates the precision of m1/m2/mean/rstd)
    n = float(B * 32 * 64)
    mu = (rw[:, None] * Y).sum(0) / n; var = (rw[:, None] * Y * Y).sum(0) / n - mu * mu; rstd = 1 / torch.sqrt(var + 1e-5)
    xh = (Y - mu) * rstd; m1 = D.sum(0) / n; m2 = (D * xh).sum(0) / n
    gam = ef.layers["sa0.%d" % l].gamma.double()
    dY2 = gam * rstd * (D - rw[:, None] * (m1 + xh * m2))
    s2 = torch.stack([dY2[seg_off[b * 32]: seg_off[(b + 1) * 32]].sum(0) for b in range(B)]).cpu()
    print("      with float64 statistics on the same D,Y: rel err %.3e" % float((s2 - tru).abs().max() / tru.abs().max()))
    print("      stat errors: mean %.2e rstd %.2e m1 %.2e m2 %.2e (relative to max)" % (
        float((mu - bn.mean.double()).abs().max() / mu.abs().max()), float((rstd - bn.rstd.double()).abs().max() / rstd.abs().max()),
        float((m1 - bb.m1.double()).abs().max() / m1.abs().max()), float((m2 - bb.m2.double()).abs().max() / m2.abs().max())))

# ---- bisect on the dense inter-level gradients (no folding involved): d/d(SA1 out) (B,128,32), d/d(SA2 out) (B,256,32), d/d(SA3 out) (B,512)
for h in hs: h.remove()
outs = {}
def mk2(name):
    def hook(mod, inp, out):
        out[1].retain_grad(); outs[name] = out[1]
    return hook
hs = [o3.feat.value_encoder[0][i].register_forward_hook(mk2("sa%d" % i)) for i in range(3)]
a64 = pi.detach().double().clone().requires_grad_(True)
pc64 = torch.cat((cl64, a64.unsqueeze(2).expand(-1, -1, cl64.shape[2])), 1)
z64 = o3.feat(pc64, value=True); z64.retain_grad()
v64 = torch.cat((z64, d["time"].double()[:, None]), 1)
q1_, q2_, _ = o3.critic(v64)
ac64 = -mix * torch.min(q1_.squeeze()[sel], q2_.squeeze()[sel]).mean()
ac64.backward()
def rel(a, b): return float((a.double().cpu() - b).abs().max() / b.abs().max())
print("d/dz (masked D of fc1 cannot be compared directly); d/d(SA3 out):", rel(mine.sc.dout[2], outs["sa2"].grad.squeeze(-1)))
print("d/d(SA2 out):", rel(mine.sc.dG[2][:, :256].reshape(B, 32, 256).transpose(1, 2), outs["sa1"].grad))
print("d/d(SA1 out):", rel(mine.sc.dout[0].reshape(B, 32, 128).transpose(1, 2), outs["sa0"].grad))
print("forward SA1 out:", rel(ctx.sa[0].out.reshape(B, 32, 128).transpose(1, 2), outs["sa0"].detach()), " SA2 out:", rel(ctx.sa[1].out.reshape(B, 32, 256).transpose(1, 2), outs["sa1"].detach()))

# ---- derived weight layouts consistent with the parameters?
def check_derived(ef, name):
    bad = []
    for key, L in ef.layers.items():
        W = L.W.detach()
        Wr = torch.roll(W, shifts=-L.rot, dims=1) if L.rot else W
        Wp = torch.zeros(L.N, L.Kp, device=W.device); Wp[:, :L.K] = Wr
        e1 = float((L.Wf[:, :L.Kp] - Wp).abs().max()) if L.Wf.shape[1] == L.Kp else float((L.Wf - W).abs().max())
        e2 = float((L.WT - Wp.t()).abs().max())
        if e1 > 0 or e2 > 0: bad.append((key, e1, e2))
    print("derived check", name, "BAD:" if bad else "ok", bad)
check_derived(mine.ef_v, "value enc"); check_derived(mine.ef_p, "policy enc")

# ---- F5 backward with want_dw=True: per-tensor value-encoder grads vs the fp64 oracle (ac loss only)
mine.ef_v.arena.g.zero_()
feat = engine.encoder_forward(ws, ef, geom, mine.cloud, mine.skip, mine.Cp_value, mine.pc.pi, ctx, time=mine.v.time, train=True)
qa = engine.critic_forward(mine.cf, feat, mine.cc5, B, nb=2)
lib.gaddpg_actor_critic_loss(dp(qa), QA_LD, QA_Q2, dp(mine.v.ret), dp(mine.v.expert_flag), float(mix), B, 1.0, QA_LD, dp(mine.cc5.dqa), mine.out.data_ptr() + 4 * 6, current_stream())
engine.critic_backward(ws, mine.cf, feat, mine.cc5, B, 2, ctx, mine.sc, accumulate=1)
engine.encoder_backward(ws, ef, ctx, mine.sc, want_dw=True, want_dbc=True)
torch.cuda.synchronize()
for (k, po), (_, pm) in zip(o3.feat.value_encoder.named_parameters(), mine._extractor.value_encoder.named_parameters()):
    go, gmn = po.grad.double(), pm.grad.cpu().double()
    print("%-24s gmax %.3e gerr/gmax %.2e" % (k, go.abs().max().item(), (go - gmn).abs().max().item() / (go.abs().max().item() + 1e-30)))

# ---- SA3 elementwise: dL/d(conv_l output) rows (B*32, C) CUDA-reconstructed vs fp64 oracle
for h in hs: h.remove()
o3.feat.zero_grad(); o3.critic.zero_grad()
grads64 = {}
seq3 = o3.feat.value_encoder[0][2].mlps[0]
hs = [seq3[3 * l].register_forward_hook(mk("sa3.%d" % l)) for l in range(3)]
a64 = pi.detach().double().clone().requires_grad_(True)
pc64 = torch.cat((cl64, a64.unsqueeze(2).expand(-1, -1, cl64.shape[2])), 1)
z64 = o3.feat(pc64, value=True)
v64 = torch.cat((z64, d["time"].double()[:, None]), 1)
q1_, q2_, _ = o3.critic(v64)
ac64 = -mix * torch.min(q1_.squeeze()[sel], q2_.squeeze()[sel]).mean()
ac64.backward()
s3 = ctx.sa[2]
for l in (2, 1, 0):
    D, Y, bn, bb = mine.sc.D[2][l].double(), s3.Y[l].double(), s3.bn[l], mine.sc.bb[2][l]
    dY = bb.g.double() * (D - (bb.m1.double() + (Y - bn.mean.double()) * bn.rstd.double() * bb.m2.double()))
    G = grads64["sa3.%d" % l].grad.squeeze(2).transpose(1, 2).reshape(B * 32, -1)
    Yo = grads64["sa3.%d" % l].detach().squeeze(2).transpose(1, 2).reshape(B * 32, -1)
    print("SA3 layer %d: dY rel err %.3e   Y rel err %.3e   D nonzero frac %.3f" % (l, rel(dY, G), rel(Y, Yo), float((D != 0).double().mean())))
    n = float(B * 32)
    gam = ef.layers["sa2.%d" % l].gamma.double()
    zo = (Yo.cuda() - Yo.cuda().mean(0)) / torch.sqrt(Yo.cuda().var(0, unbiased=False) + 1e-5) * gam + ef.layers["sa2.%d" % l].beta.double()
    zm = Y * bn.scale.double() + bn.shift.double()
    flips = ((zo > 0) != (zm > 0)).sum().item()
    print("      relu-mask flips vs fp64: %d of %d; min |z| at flips %s" % (flips, zo.numel(), zo[(zo > 0) != (zm > 0)].abs().tolist()[:5]))

# ---- how accurate is the fp32 ORACLE's forward against fp64, at the same places?
y32 = {}
def mk3(name):
    def hook(mod, inp, out): y32[name] = out.detach()
    return hook
h3 = []
for i in range(3):
    sq = o2.feat.value_encoder[0][i].mlps[0]
    h3 += [sq[3 * l].register_forward_hook(mk3("sa%d.%d" % (i, l))) for l in range(3)]
y64 = {}
def mk4(name):
    def hook(mod, inp, out): y64[name] = out.detach()
    return hook
for i in range(3):
    sq = o3.feat.value_encoder[0][i].mlps[0]
    h3 += [sq[3 * l].register_forward_hook(mk4("sa%d.%d" % (i, l))) for l in range(3)]
with torch.no_grad():
    z32 = o2.features(d["cloud"], d["time"], pi.detach(), value=True)
    z64b = o3.feat(torch.cat((cl64, pi.detach().double().unsqueeze(2).expand(-1, -1, cl64.shape[2])), 1), value=True)
for k in sorted(y32):
    print("fp32 oracle vs fp64  conv out %s: rel err %.2e" % (k, float((y32[k].double() - y64[k]).abs().max() / y64[k].abs().max())))
print("fp32 oracle z vs fp64:", float((z32[:, :512].double() - z64b).abs().max() / z64b.abs().max()), "  CUDA z vs fp64:", float((feat[:, :512].cpu().double() - z64b).abs().max() / z64b.abs().max()))
# CUDA conv outputs for SA3 (dense rows) vs fp64
for l in range(3):
    Yo = y64["sa2.%d" % l].squeeze(2).transpose(1, 2).reshape(B * 32, -1)
    print("CUDA vs fp64 conv out sa2.%d: %.2e" % (l, rel(ctx.sa[2].Y[l], Yo)))
