"""GPU: the sample-sharded (N>1) path of the fused agent with two ranks on ONE device over gloo — same code path as the
NCCL run (one contiguous gradient range per optimiser phase, reduced in two asynchronous pieces around the SA1 backward,
between the captured phases), cheap enough for every round.

SURVEY.md §8(e): "multi-GPU parity test = average of the single-GPU shard runs".  Every rank therefore also runs, in the
same process, two single-GPU agents S0 / S1 on shard 0 / shard 1 in LOCKSTEP — the same phases, with the all-reduce
replaced by the plain average of the two agents' gradient pools ((g0 + g1) * 0.5: the arithmetic gloo's SUM + scale
performs).  The sharded agent on rank r must then equal S_r BIT FOR BIT after every step: gradients of both phases,
all parameters, BatchNorm statistics, Adam moments and the 11 returned scalars.  At the first step the phase-1 gradients
of S0 / S1 are by construction those of two independent single-GPU runs, so "sharded gradient = mean of the shard
gradients" is checked literally as well.  Runs with CUDA graphs on (no deadlock with the collective's own stream) and
in both reduce modes (split around the SA1 backward / one blocking call per phase)."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = """
import sys, os, numpy as np, torch
sys.path.insert(0, %r)
os.environ["GADDPG_NO_REBUILD"] = "1"
import torch.distributed as dist
from gaddpg_b200.dist import World
from gaddpg_b200 import agent as ag, synthetic
torch.cuda.set_device(0)
w = World(backend="gloo")
SPLIT = bool(int(os.environ["SPLIT_REDUCE"]))
B = 8

def shard(full, u, r):
    lo, hi = r * B, (r + 1) * B
    return {k: (v[lo:hi] if hasattr(v, "shape") and v.shape[:1] == (2 * B,) else v) for k, v in full.items()}, u[lo:hi]

def lockstep(agents, shards):
    # the single-GPU agents' own phases (agent.DDPGB200.update_parameters), all-reduce := average of the two pools
    def avg(pool_name):
        p0, p1 = (getattr(a, pool_name).used for a in agents)
        m = (p0 + p1) * 0.5
        p0.copy_(m); p1.copy_(m)
    for a, (b, u) in zip(agents, shards):
        a._begin_step()
        a.prepare_data(b, u, after_clouds=a._geometry)
    a0 = agents[0]
    even = (a0.update_step %% a0.policy_update_gap) == 0
    hard = (a0.update_step %% a0.target_update_interval) == 0
    sig = (a0._mix_idx(),)
    g1 = []
    for a in agents:
        a._set_dyn(("critic", "venc", "policy", "enc"))
        a._phase1(sig)
        g1.append(a.gpool_c.used.clone())
    avg("gpool_c")
    for a in agents:
        a._run(("p2", even) + sig, lambda a=a: a._phase2(even))
    avg("gpool_a")
    out = []
    for a in agents:
        a._run(("p3", hard) + sig, lambda a=a: a._phase3(hard))
        a.update_step += 1
        out.append(a._finish())
    return out, g1

mine = ag.make_agent("DDPG", seed=123456, device="cuda:0", world=w)
mine.split_reduce = SPLIT
singles = [ag.make_agent("DDPG", seed=123456, device="cuda:0") for _ in range(2)]
for s in singles:
    s.use_graph = False
indep = ag.make_agent("DDPG", seed=123456, device="cuda:0")     # a plain single-GPU run on THIS rank's shard, step 1 only
for step in range(5):
    full = synthetic.make_batch(2 * B, 512, step=step)
    u = np.random.RandomState(step).rand(2 * B, 6).astype(np.float32)
    mine_b, mine_u = shard(full, u, w.rank)
    out = mine.update_parameters(mine_b, mine.update_step, 0, noise_u=mine_u)
    mine.step_scheduler(mine.update_step)
    assert all(np.isfinite(v) for v in out.values()), out
    ref_out, g1 = lockstep(singles, [shard(full, u, 0), shard(full, u, 1)])
    for s in singles:
        s.step_scheduler(s.update_step)
    ref = singles[w.rank]
    assert out == ref_out[w.rank], (step, out, ref_out[w.rank])
    for name in ("gpool_c", "gpool_a"):
        assert torch.equal(getattr(mine, name).used, getattr(ref, name).used), (step, name, "sharded gradients != mean of the shard gradients")
    for a, b in ((mine.ef_p, ref.ef_p), (mine.ef_v, ref.ef_v), (mine.pf, ref.pf), (mine.cf, ref.cf)):
        assert torch.equal(a.arena.p, b.arena.p) and torch.equal(a.arena.m, b.arena.m) and torch.equal(a.arena.v, b.arena.v), step
    for a, b in ((mine.ef_p, ref.ef_p), (mine.ef_v, ref.ef_v)):
        assert torch.equal(a.buffers.p, b.buffers.p) and torch.equal(a.nbt, b.nbt), (step, "BatchNorm running statistics")
    assert torch.equal(mine.pft.arena.p, ref.pft.arena.p) and torch.equal(mine.cft.arena.p, ref.cft.arena.p), step
    if step == 0:
        # literally: a stand-alone single-GPU update on this rank's shard produces the phase-1 gradients that were averaged
        indep.use_graph = False
        indep._begin_step(); indep.prepare_data(mine_b, mine_u, after_clouds=indep._geometry)
        indep._set_dyn(("critic", "venc", "policy", "enc")); indep._phase1((indep._mix_idx(),))
        assert torch.equal(indep.gpool_c.used, g1[w.rank]), "lockstep agent is not a plain single-GPU evaluation"
        both = [torch.zeros_like(g1[0]) for _ in range(2)]
        dist.all_gather(both, indep.gpool_c.used)
        # value-encoder range (never clipped in place): sharded gradient == mean over the ranks' independent gradients
        n = mine.ef_v.arena.n
        assert torch.equal(ref.gpool_c.used[:n], ((both[0] + both[1]) * 0.5)[:n])
    # replicas identical: compare arena checksums across ranks
    cs = torch.stack([a.p.double().sum() for a in (mine.ef_p.arena, mine.ef_v.arena, mine.pf.arena, mine.cf.arena)]).cpu()
    other = [torch.zeros_like(cs) for _ in range(2)]
    dist.all_gather(other, cs)
    assert torch.equal(other[0], other[1]), (step, other)
w.barrier(); w.close()
print("ok", w.rank)
"""


@pytest.mark.parametrize("split", [1, 0])
def test_two_ranks_equal_mean_of_single_gpu_shards(tmp_path, cuda, split):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent(WORKER % ROOT))
    port = str(29633 + split)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=port, SPLIT_REDUCE=str(split))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", port, str(script)], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("ok") == 2
