"""select_action latency (B = 1, eval-mode policy path, CUDA graph) with the unfused SA1 kernels vs the fused TMA-fed chain."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gaddpg_b200 import agent as ag, engine, synthetic

for N in (1024, 4096):
    cloud = synthetic.make_batch(1, N, step=3)["point_state_batch"][0]
    for fused in (False, True):
        engine.FUSED_SA1 = fused
        a = ag.make_agent("DDPG", seed=1)
        for _ in range(5):
            r = a.select_action([[cloud, None]], remain_timestep=5, eps=np.zeros((1, 6), np.float32))
        t0 = time.perf_counter()
        for _ in range(200):
            a.select_action([[cloud, None]], remain_timestep=5, eps=np.zeros((1, 6), np.float32))
        dt = (time.perf_counter() - t0) / 200
        # device time of the graph alone
        st = a._act_states[("act", 1, 4, N + 6)]
        g = a._graphs[st.key]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(50): g.replay()
        e1.record(); torch.cuda.synchronize()
        print("N=%d fused=%d: %.1f us per action (host wall), graph alone %.1f us, action %s" % (N, fused, 1e6 * dt, 20 * e0.elapsed_time(e1), np.round(r[0][:3], 5)), flush=True)
