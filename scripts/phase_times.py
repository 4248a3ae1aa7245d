"""Where the update step's time goes, per graph segment (GADDPG_WHOLE_GRAPH=0 layout): CUDA events around every segment replay,
averaged over even / odd steps.  Segments on the side stream (target chain, next-cloud geometry) are timed on their own stream."""
import os, sys
os.environ["GADDPG_WHOLE_GRAPH"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import collections
import torch
from gaddpg_b200 import agent as A, synthetic

torch.cuda.set_device(0)
B, N = int(os.environ.get("B", 256)), 4096
ag = A.make_agent("DDPG", seed=123456, extra_latent=3, policy_aux=False, critic_aux=False)
batches = [synthetic.make_batch(B, N, step=s, channels=6) for s in range(2)]
dev = [{k: (torch.from_numpy(v).cuda() if hasattr(v, "dtype") and v.dtype.kind == "f" else v) for k, v in b.items()} for b in batches]
rec = collections.defaultdict(list)
orig = ag._run
def timed(key, fn, outer=False):
    s = torch.cuda.current_stream()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s); orig(key, fn, outer); b.record(s)
    rec[key[0] + ("/even" if (len(key) > 1 and key[1] is True and key[0].startswith("p2")) else "")].append((a, b))
ag._run = timed
for i in range(12):
    if i == 6:
        rec.clear()
    t0 = torch.cuda.Event(enable_timing=True); t0.record()
    ag.update_parameters(dev[i % 2]); ag.step_scheduler()
    t1 = torch.cuda.Event(enable_timing=True); t1.record()
    rec["whole step (" + ("even" if (ag.update_step - 1) % 2 == 0 else "odd") + ")"].append((t0, t1))
torch.cuda.synchronize()
for k, v in rec.items():
    ts = [a.elapsed_time(b) for a, b in v]
    print("%-22s n=%2d  mean %.3f ms" % (k, len(ts), sum(ts) / len(ts)))
