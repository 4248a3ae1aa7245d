"""CPU, world_size 2 over gloo: the sharding helper and the mean all-reduce that the N>1 path uses
(the CUDA agents call exactly World.all_reduce_mean on their gradient arenas)."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(textwrap.dedent("""
        import sys, torch
        sys.path.insert(0, %r)
        from gaddpg_b200.dist import World
        w = World(backend="gloo")
        lo, hi = w.shard(8)
        assert (lo, hi) == (4 * w.rank, 4 * w.rank + 4)
        # average of per-rank mean gradients == gradient of the global mean loss (equal shard sizes)
        torch.manual_seed(0)
        x = torch.randn(8, 5)
        g_local = x[lo:hi].mean(0)
        w.all_reduce_mean(g_local)
        assert torch.allclose(g_local, x.mean(0), atol=1e-6)
        t = torch.tensor([float(w.rank + 1)])
        assert float(w.all_reduce_max(t)) == 2.0
        w.barrier(); w.close()
        print("ok", w.rank)
    """ % ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29611")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                          "127.0.0.1", "--master-port", "29611", str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2
