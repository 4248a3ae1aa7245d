#!/usr/bin/env python
"""The offline training loop of the reference (core/train_test_offline.py:107-161, ``train_off_policy``) on the B200
path, end to end and without the reference tree: device-resident replay -> fused DDPG / BC update -> schedulers ->
reference-format checkpoints -> ``select_action``.

    python examples/offline_loop.py --policy DDPG --updates 40 --batch 64 --points 1024 --out /tmp/gaddpg_demo

The released replay data cannot be downloaded here, so the buffer is filled with synthetic rollouts
(``gaddpg_b200.synthetic.make_episode``); with the real data, replace that block by ``memory.load(data_dir)`` — the
``.npz`` layout is the reference's.  Every call below is the one ``train_off_policy`` makes on the reference classes.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--policy", default="DDPG", choices=["DDPG", "BC"])
    ap.add_argument("--updates", type=int, default=40)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--points", type=int, default=1024, help="uniform_num_pts of the replay clouds (reference default 1024)")
    ap.add_argument("--episodes", type=int, default=40)
    ap.add_argument("--out", default="/tmp/gaddpg_b200_demo")
    ap.add_argument("--synchronous", action="store_true",
                    help="the reference's strictly sequential order (sample, update, read the scalars) instead of the double-buffered "
                         "feed loop (feed.FeedLoop: step i+1 is staged and its geometry built while step i computes)")
    args = ap.parse_args()

    from gaddpg_b200 import agent as ag, synthetic
    from gaddpg_b200.config import LOSS_KEYS
    from gaddpg_b200.replay_memory import ReplayMemoryB200

    np.random.seed(233)                                                     # core/train_test_offline.py:33
    agent = ag.make_agent(args.policy, seed=123456)
    memory = ReplayMemoryB200(args.episodes * 30 + 1, uniform_num_pts=args.points, RL=args.policy == "DDPG")
    for e in range(args.episodes):                                          # stands in for memory.load(cfg.RL_SAVE_DATA_ROOT_DIR, ...)
        memory.add_episode(synthetic.make_episode(int(np.random.randint(8, 30)), args.points, seed=e, success=e % 4 != 0))
    print("replay: %d transitions in HBM (%.1f MB)" % (len(memory), memory.point_state[: len(memory)].numel() * 4 / 1e6))

    losses = {k: [] for k in LOSS_KEYS}
    t0 = time.time()
    if args.synchronous:
        for i in range(args.updates):
            batch_data = memory.sample(batch_size=args.batch)                   # train_test_offline.py:120
            loss = agent.update_parameters(batch_data, agent.update_step, i)    # :123
            agent.step_scheduler(agent.update_step)                             # :129
            for k, v in loss.items():
                losses[k].append(v)
    else:
        # the same three calls per update, issued by feed.FeedLoop one step ahead of the device (trainer.py:202-293 overlaps the next
        # minibatch with the current update through ray; here it is done in-process) — identical losses, bit for bit
        from gaddpg_b200.feed import FeedLoop

        for loss in FeedLoop(agent, memory, args.batch).train_iter(args.updates):
            for k, v in loss.items():
                losses[k].append(v)
    dt = time.time() - t0
    print("%d updates of %d samples in %.2f s (%.1f updates/s, first calls include CUDA-graph capture)" % (args.updates, args.batch, dt, args.updates / dt))
    for k, v in losses.items():
        if np.nanmean(v) != 0:
            print("  %-24s first %.5f  last %.5f" % (k, v[0], v[-1]))
    print("lr:", agent.get_lr())

    files = agent.save_model(agent.update_step, output_dir=args.out)        # :133 — the reference's three files
    print("saved:", ", ".join(sorted(os.path.basename(f) for f in files.values() if os.path.exists(f))))
    twin = ag.make_agent(args.policy, seed=1)
    assert twin.load_model(args.out) == agent.update_step

    state = memory.sample(1)["point_state_batch"][0].cpu().numpy()          # one observation, as the rollout loop passes it
    action, _, _, aux = agent.select_action([[state, None]], remain_timestep=10)   # :232
    action2, _, _, _ = twin.select_action([[state, None]], remain_timestep=10)
    assert np.array_equal(action, action2), "the reloaded agent must act identically"
    print("action:", np.round(action, 4), " aux:", np.round(aux, 3))


if __name__ == "__main__":
    main()
