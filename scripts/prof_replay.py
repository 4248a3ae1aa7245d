"""Target for `ncu -k regex:replay_gather` : a few cfg2-sized minibatch gathers from a device-resident replay store
(B = 256 clouds of 6 x 4102 floats out of 2048 stored transitions; store + outputs exceed the 126 MB L2)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gaddpg_b200.replay_memory import ReplayMemoryB200  # noqa: E402

cap, N, B, C = 2048, 4096, 256, 6
mem = ReplayMemoryB200(cap, uniform_num_pts=N, channels=C)
mem.point_state.copy_(torch.randn(cap, C, N + 6, device="cuda"))
mem.episode_map[:] = np.minimum((np.arange(cap) // 16) * 16 + 15, cap - 1)
mem.timestep[:] = np.arange(cap) % 16 + 1
mem.cur_idx, mem.is_full = 0, True
mem._mark(0, cap)
np.random.seed(0)
for _ in range(6):
    mem.sample(B).materialise()
torch.cuda.synchronize()
print("ok")
