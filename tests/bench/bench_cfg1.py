"""BASELINE config 1 — the reference's own CPU-runnable case: BC, batch 8, 512-point clouds, one update step.
Times BCB200.update_parameters (pinned host batches through the public API, CUDA graphs on) and, beside it, the oracle
port of the reference's BC step on the host cores.  Lives under tests/ because it imports the oracle.

    python tests/bench/bench_cfg1.py
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gaddpg_b200 import agent as ag, synthetic  # noqa: E402
from oracle.ddpg_cpu import OracleAgent  # noqa: E402

B, N, STEPS = 8, 512, 200
mine = ag.make_agent("BC", seed=123456)
batches = []
for i in range(4):
    b = synthetic.make_batch(B, N, step=i)
    batches.append({k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).pin_memory() for k, v in b.items()
                    if k not in ("grasp_sample_batch", "batch_idx")})
for i in range(10):
    mine.update_parameters(batches[i % 4], mine.update_step, 0)
    mine.step_scheduler(mine.update_step)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(STEPS):
    mine.update_parameters(batches[i % 4], mine.update_step, 0)
    mine.step_scheduler(mine.update_step)
torch.cuda.synchronize()
gpu = STEPS / (time.perf_counter() - t0)
torch.set_num_threads(os.cpu_count())
ora = OracleAgent("BC", seed=123456)
raw = [synthetic.make_batch(B, N, step=i) for i in range(4)]
ora.update_parameters(raw[0])
t0 = time.perf_counter()
for i in range(20):
    ora.update_parameters(raw[i % 4])
    ora.step_scheduler()
cpu = 20 / (time.perf_counter() - t0)
print(json.dumps(dict(workload="cfg1: BC update, B=8, N=512 points x 4 channels, one step", gpu_steps_per_s=gpu, ms_per_step=1e3 / gpu,
                      cpu_oracle_steps_per_s=cpu, cpu_cores=os.cpu_count(), timing="host wall clock, synchronous public API")))
