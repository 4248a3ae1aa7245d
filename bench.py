#!/usr/bin/env python
"""bench.py — offline DDPG update steps/s on B200 (BASELINE.json metric) + roofline + CPU baseline.

    python bench.py --gpus 1 --steps K --warmup W            # fused CUDA agent (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --steps K --warmup W     # the reference algorithm's CPU path (oracle port), host cores

One "step" = one ``agent.update_parameters`` (ddpg.py:146-185) + ``step_scheduler`` on one synthetic replay
minibatch.  Workload at N=1: BASELINE config 2 — DDPG, B=256, 4096-point x 6-channel clouds (extra_latent=3),
policy_aux = critic_aux = False; under N ranks every rank takes its own 256-sample shard (weak scaling) with the
two per-step gradient all-reduces over NCCL.

JSON keys beyond the base contract: ``roofline`` (dominant entry point of the step, timed live with CUDA events in
an eager profiled pass after the timed region), ``cpu_baseline`` (oracle port on the host cores, bounded sample),
``kernels`` (top entry points with their share of the step), ``clocks``.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "offline DDPG update steps/sec (B=256, 4096 pts)"
UNIT = "update steps/s (256-sample minibatch per GPU)"
FP32_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12  # 148 SMs x 128 FFMA lanes x 2 flop x max SM clock = 74.5


def workload(args):
    name = "cfg3" if getattr(args, "aux", False) else "cfg2"
    return dict(workload="%s: offline TD3/DDPG update, B=%d per GPU, N=%d points x 6 channels (extra_latent=3), aux heads %s"
                % (name, args.batch, args.points, "on (goal-aux + grasp-aux losses)" if name == "cfg3" else "off"),
                B=args.batch, N=args.points, channels=6, policy_update_gap=2,
                l2="per-step working set (activations of 5 encoder passes, ~GBs) >> 126 MB L2 and 4 distinct batches are cycled: no flush needed")


AGENT_KW = dict(extra_latent=3, policy_aux=False, critic_aux=False)
# dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured launches of that kernel) from the
# `ncu --set full` capture summarised in profiles/ (same workload: cfg2); keyed like profile()'s kernel families
NCU_TRAFFIC = {   # profiles/r1_ncu_full_v2.md (SA1 shapes, M = 423 k rows): bytes per launch, by shape label
    "gemm_nt:tc:": {"nt[64x64,a1,e0]": 166.1e6, "nt[128x64,a1,e0]": 264.7e6, "nt[64x64,a2,e1]": 409.0e6, "nt[64x128,a3,e1]": 430.6e6},
    "gemm_tn:": {"tn[128x64,p3,q1]": 343.4e6, "tn[64x64,p2,q1]": 331.5e6},
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
               0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while not self.stop_flag and self.nv is not None:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        s = sorted(self.samples)
        return dict(sm_mhz=s[len(s) // 2] if s else None, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(s))


def make_batches(B, N, nb, pinned=True):
    import numpy as np
    import torch

    from gaddpg_b200 import synthetic

    out = []
    for i in range(nb):
        b = synthetic.make_batch(B, N, step=i, channels=6)
        t = {}
        for k, v in b.items():
            if k in ("grasp_sample_batch", "batch_idx"):
                continue
            x = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
            t[k] = x.pin_memory() if pinned else x
        t["noise_u"] = torch.rand(B, 6).pin_memory() if pinned else torch.rand(B, 6)
        out.append(t)
    return out


REF_MAX_TIMED, REF_MAX_WARMUP = 4, 1   # a real B=256 CPU step takes ~10 s on the box's 16 cores: cap the reference arm


def _oracle_steps(args, warmup, steps):
    """Real full-size steps of the oracle port (pinned bit-exactly to the unmodified reference, oracle/make_golden.py): the
    SAME minibatch size, cloud size and channel count as the GPU arm — nothing is scaled.  -> (seconds, cores, torch)"""
    import torch

    from gaddpg_b200 import synthetic
    from oracle.ddpg_cpu import OracleAgent

    cores = os.cpu_count()
    torch.set_num_threads(cores)
    agent = OracleAgent("DDPG", seed=123456, **AGENT_KW)
    batches = [synthetic.make_batch(args.ref_batch, args.points, step=i, channels=6) for i in range(2)]
    for i in range(warmup):                       # update_step 1 (odd): no actor-critic branch
        agent.update_parameters(batches[i % 2])
        agent.step_scheduler()
    t0 = time.perf_counter()
    for i in range(steps):                        # alternating even / odd update steps, starting with an even one
        agent.update_parameters(batches[(i + warmup) % 2])
        agent.step_scheduler()
    return time.perf_counter() - t0, cores, torch.__version__


def run_reference(args):
    """The reference algorithm's CPU implementation of the step (oracle port of core/ddpg.py + networks + pointnet2_ops,
    pinned to the unmodified reference by oracle/make_golden.py) on all host threads at the REAL workload size.  A full
    B=256 step costs ~10 s of CPU, so the arm times at most REF_MAX_TIMED steps after at most REF_MAX_WARMUP warm-up
    steps whatever --steps / --warmup say, and prints the counts it actually ran."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warmup, steps = min(args.warmup, REF_MAX_WARMUP), max(1, min(args.steps, REF_MAX_TIMED))
    dt, cores, tver = _oracle_steps(args, warmup, steps)
    scale = args.ref_batch / float(args.batch)
    value = steps / dt * scale
    cfg = workload(args)
    extrap = args.ref_batch != args.batch
    if extrap:  # only on explicit request (--ref-batch): say so and do not pose as the same configuration
        cfg.update(B=args.ref_batch, extrapolated=True)
    sample = "%d real DDPG update steps (alternating even / odd) on a %d-sample minibatch (N=%d, 6 ch) after %d warm-up; %.1f s of CPU work%s" % (
        steps, args.ref_batch, args.points, warmup, dt, "; value scaled by %d/%d (EXTRAPOLATED)" % (args.ref_batch, args.batch) if extrap else "; nothing scaled")
    print(json.dumps(dict(metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=steps, warmup=warmup,
                          requested=dict(steps=args.steps, warmup=args.warmup), ms_per_step=1e3 * dt / steps,
                          higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", config=cfg,
                          impl="reference", extrapolated=extrap,
                          cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample, torch=tver),
                          e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))))


def cpu_baseline(args):
    """1 warm-up + 2 timed (one even, one odd) REAL B=256 steps of the oracle port: ~30 s on the box's host cores."""
    n = args.cpu_steps
    dt, cores, tver = _oracle_steps(args, 1, n)
    extrap = args.ref_batch != args.batch
    return dict(value=n / dt * (args.ref_batch / float(args.batch)), unit=UNIT, cores=cores, kind="port", torch=tver, extrapolated=extrap,
                sample="oracle port, %d timed real DDPG steps (alternating even / odd) on a %d-sample minibatch (N=%d, 6 ch) after 1 warm-up; "
                       "%.1f s of CPU work%s" % (n, args.ref_batch, args.points, dt, "; scaled by %d/%d (EXTRAPOLATED)" % (args.ref_batch, args.batch) if extrap else "; nothing scaled"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--points", type=int, default=4096)
    ap.add_argument("--ref-batch", type=int, default=0, help="minibatch of the CPU arm (default: --batch, i.e. the real workload)")
    ap.add_argument("--cpu-steps", type=int, default=2, help="timed CPU-baseline steps (about 10 s each at B=256 on 16 cores)")
    ap.add_argument("--no-profile", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--reduce-mode", default="single", choices=["split", "single"],
                    help="N>1: 'split' all-reduces everything but SA1's gradients asynchronously behind the SA1 backward (2 ranges per "
                         "phase, one hidden); 'single' is one blocking all-reduce per optimiser phase")
    ap.add_argument("--no-extra", action="store_true", help="skip the e2e_f64 and dense-worst-case legs")
    ap.add_argument("--no-replay", action="store_true", help="skip the device-resident replay leg (SURVEY.md §8 row f1)")
    ap.add_argument("--aux", action="store_true", help="BASELINE config 3: both auxiliary losses on (use with --batch 512); "
                    "the default run is config 2, the one the metric is quoted on")
    args = ap.parse_args()
    if args.aux:
        AGENT_KW.update(policy_aux=True, critic_aux=True)
    if args.ref_batch <= 0:
        args.ref_batch = args.batch
    global UNIT
    UNIT = "update steps/s (%d-sample minibatch per GPU)" % args.batch
    if args.impl == "reference":
        return run_reference(args)

    import torch

    from gaddpg_b200 import agent as ag
    from gaddpg_b200.capi import lib
    from gaddpg_b200.dist import World

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product has no CPU fallback (use --impl reference for the CPU arm)")
    world = World("nccl") if int(os.environ.get("WORLD_SIZE", "1")) > 1 else None
    rank = world.rank if world else 0
    local = world.local_rank if world else 0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    W = max(args.warmup, 3)
    K = args.steps
    torch.manual_seed(1000 + rank)
    agent = ag.make_agent("DDPG", seed=123456, device=dev, world=world, **AGENT_KW)
    agent.use_graph = not args.no_graph
    agent.split_reduce = args.reduce_mode == "split"
    agent.static_device_batches = True   # the device-resident minibatches of the `value` leg are generated once, before any timing
    nb = 4
    host = make_batches(args.batch, args.points, nb)
    devb = [{k: v.to(dev) for k, v in b.items()} for b in host]

    def sync_all():
        if world:
            world.barrier()
        torch.cuda.synchronize()

    def run(batches, n, timed, pipelined=False):
        """``pipelined``: the double-buffered feed loop of feed.FeedLoop (step i+1 is enqueued before step i's scalars are read;
        every step's 11 scalars are still read back inside the timed region) — removes the ~70 us per step the GPU otherwise
        idles between two synchronous update_parameters calls."""
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pending = None
        for i in range(n):
            b = batches[i % nb]
            h = agent.update_parameters(b, agent.update_step, 0, noise_u=b["noise_u"], defer=pipelined)
            agent.step_scheduler(agent.update_step)
            if pipelined:
                if pending is not None:
                    pending.result()
                pending = h
        if pending is not None:
            pending.result()
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        if world and timed:
            t = torch.tensor([ms], device=dev)
            world.all_reduce_max(t)
            ms = float(t)
        return ms

    # ---- launches per step (eager, one odd + one even step) — also the first warm-up
    agent.use_graph = False
    l0 = lib.gaddpg_launch_count()
    run(devb, 2, False)
    launches_per_2 = lib.gaddpg_launch_count() - l0
    agent.use_graph = not args.no_graph
    run(devb, max(W, 4), False)  # graph warm-up + capture for both parities

    # ---- value: inputs resident in HBM
    clk = ClockSampler(local)
    clk.start()
    ms_dev_sync = run(devb, K, True)
    ms_dev = run(devb, K, True, pipelined=True)
    # ---- e2e: pinned host buffers through the public API (H2D of the batch + D2H of the scalars inside the timed region)
    run(host, 2, False)
    ms_e2e_sync = run(host, K, True)
    run(host, 2, False, pipelined=True)
    ms_e2e = run(host, K, True, pipelined=True)
    clk.stop_flag = True
    clk.join(timeout=2)
    n_gpus = world.size if world else 1
    value = n_gpus * K / (ms_dev / 1e3)
    e2e = n_gpus * K / (ms_e2e / 1e3)

    out = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=n_gpus, steps=K, warmup=W, ms_per_step=ms_dev / K, higher_is_better=True,
               scaling="weak", vs_baseline=None, dtype="f32", data="synthetic", config=workload(args),
               e2e=dict(value=e2e, unit=UNIT, ms_per_step=ms_e2e / K, h2d_bytes_per_step=agent.h2d_bytes(), d2h_bytes_per_step=64),
               gpu_launches=int(launches_per_2 * K / 2), clocks=clk.result(), graph=bool(agent.use_graph),
               feed=dict(value="pipelined (feed.FeedLoop order: step i+1 enqueued before step i's scalars are read back; all reads inside "
                               "the timed region); the staging copies and the xyz-only geometry kernels (FPS, ball query, row tables) of "
                               "minibatch i+1 run on prep streams beside step i (per-slot input buffers)", value_synchronous=n_gpus * K / (ms_dev_sync / 1e3), ms_per_step_synchronous=ms_dev_sync / K,
                         e2e="pipelined through the same public API (update_parameters(batch, defer=True) = feed.FeedLoop): the pinned-host -> "
                             "device copy of minibatch i+1 runs on a copy stream while step i computes; every step copies its own inputs "
                             "and reads its own 64 bytes of scalars inside the timed region",
                         e2e_synchronous=n_gpus * K / (ms_e2e_sync / 1e3), e2e_ms_per_step_synchronous=ms_e2e_sync / K),
               collectives=(dict(mode=args.reduce_mode, calls_per_step=4 if args.reduce_mode == "split" else 2,
                                 note="one contiguous gradient range per optimiser phase (value encoder + critic | policy encoder + "
                                      "policy); 'split' reduces all but the SA1 block asynchronously behind the SA1 backward")
                            if world else None))

    if not args.no_extra:
        # ---- e2e_f64: the dict the reference's BaseMemory.sample returns (replay_memory.py:166-176,376): float64 ndarray
        # clouds, ndarray fields — the host f64->f32 conversion of 2 x B x C x (N+6) x 8 bytes is inside the timed region
        import numpy as np

        from gaddpg_b200 import synthetic

        host64 = []
        for i in range(nb):
            b = synthetic.make_batch(args.batch, args.points, step=i, channels=6, dtype=np.float64)
            b["noise_u"] = host[i]["noise_u"]
            host64.append(b)
        run(host64, 2, False)
        k64 = max(4, K // 2)
        ms64 = run(host64, k64, True)          # synchronous: the host float64 -> float32 conversion sits on the critical path
        out["e2e_f64"] = dict(value=n_gpus * k64 / (ms64 / 1e3), unit=UNIT, ms_per_step=ms64 / k64, steps=k64,
                              host_bytes_converted_per_step=2 * host64[0]["point_state_batch"].nbytes,
                              h2d_bytes_per_step=agent.h2d_bytes(), d2h_bytes_per_step=64,
                              note="update_parameters(dict of float64 ndarrays, exactly what BaseMemory.sample returns); the reference "
                                   "pays the same conversion in torch.cuda.FloatTensor(v) (agent.py:221-222)")
        # the same float64 dicts through feed.FeedLoop: a worker thread converts / pins minibatch i+1 while step i runs
        from gaddpg_b200.feed import FeedLoop

        class _Cycle:
            def __init__(self, batches):
                self.b, self.i = batches, 0

            def sample(self, B):
                self.i += 1
                return self.b[(self.i - 1) % len(self.b)]

        loop = FeedLoop(agent, _Cycle(host64), args.batch)
        loop.train_iter(4)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loop.train_iter(k64)
        e1.record()
        sync_all()
        msp = e0.elapsed_time(e1)
        if world:
            t = torch.tensor([msp], device=dev)
            world.all_reduce_max(t)
            msp = float(t)
        out["e2e_f64"].update(pipelined_value=n_gpus * k64 / (msp / 1e3), pipelined_ms_per_step=msp / k64,
                              pipelined_note="feed.FeedLoop(agent, memory).train_iter: prefetch thread (float64 -> pinned float32) + copy "
                                             "stream + deferred result read-back; same minibatches, same results")
        del host64
        # ---- dense worst case: clouds that defeat duplicate folding (every SA1 ball holds >= 64 distinct points, all 32 SA2
        # centroids lie within one radius): M1 = B*32*64, M2 = B*32*32 live rows — the upper bound of the data-dependent cost
        dense = []
        for i in range(2):
            b = synthetic.make_batch(args.batch, args.points, step=100 + i, channels=6, compact=True)
            t = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(dev) for k, v in b.items()
                 if k not in ("grasp_sample_batch", "batch_idx")}
            t["noise_u"] = devb[i]["noise_u"]
            dense.append(t)
        dense = dense * 2
        run(dense, 4, False)
        msd = run(dense, K, True)
        out["dense_worst_case"] = dict(value=n_gpus * K / (msd / 1e3), unit=UNIT, ms_per_step=msd / K,
                                       live_rows={"sa1": int(agent.geom_s.lv[0].seg_off[-1]), "sa2": int(agent.geom_s.lv[1].seg_off[-1])},
                                       dense_rows={"sa1": agent.B * 32 * 64, "sa2": agent.B * 32 * 32},
                                       note="2 cm-cube objects: no duplicate rows to fold at SA1, all 32 x 32 centroid pairs live at SA2")
        del dense
        run(devb, 2, False)   # back to the metric's clouds for the legs below
    if not args.no_replay:
        # every rank runs its own device-resident store (the N>1 number is the honest end-to-end figure for a device store)
        out["replay"] = replay_leg(agent, devb, args, run, nb, n_gpus, rank == 0 and n_gpus == 1)
    if rank == 0 and n_gpus == 1 and not args.no_replay:
        # row f2: select_action latency (B = 1, eval-mode policy path; H2D of the cloud + one graph replay + D2H of 80 bytes)
        cloud1 = host[0]["point_state_batch"][0].numpy()
        for _ in range(4):
            agent.select_action([[cloud1, None]], remain_timestep=5)
        t0 = time.perf_counter()
        for _ in range(50):
            agent.select_action([[cloud1, None]], remain_timestep=5)
        out["select_action"] = dict(ms_per_action=1e3 * (time.perf_counter() - t0) / 50, points=args.points,
                                    timing="host wall clock over 50 synchronous calls (each ends with a device->host read)")
    if rank == 0 and not args.no_profile:
        out.update(profile(agent, devb, args))
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args)
    if world:
        world.barrier()
    if rank == 0:
        print(json.dumps(out))
    if world:
        # The captured whole-step graphs hold NCCL kernels; destroying the process group while they exist blocked the 8-rank run
        # at exit (the JSON line was out, the workers never returned).  Drop the graphs, drain the device, and leave without the
        # collective teardown: every rank has passed the barrier above, nothing is in flight.
        agent._graphs.clear()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def replay_leg(agent, devb, args, run, nb, n_gpus=1, gather_bench=True):
    """Row f1: the same update fed from the device-resident replay buffer (ReplayMemoryB200.sample = host index draw +
    one gather launch) instead of pinned host batches, and the gather kernel alone against the HBM roofline
    (algorithmic bytes = 2 clouds x B rows read once + written once, plus the 128-byte records)."""
    import numpy as np
    import torch

    from gaddpg_b200 import replay_memory as rm

    B, N = args.batch, args.points
    clouds = torch.cat([b[k] for b in devb for k in ("point_state_batch", "next_point_state_batch")])
    cap, C = clouds.shape[0], clouds.shape[1]
    mem = rm.ReplayMemoryB200(cap, uniform_num_pts=N, channels=C, device=clouds.device)
    mem.point_state.copy_(clouds)
    del clouds
    rs = np.random.RandomState(0)
    cat = lambda k: torch.cat([b[k].reshape(B, -1) for b in devb] * 2).cpu().numpy()  # noqa: E731
    mem.action[:], mem.expert_action[:], mem.goal[:] = cat("action_batch"), cat("expert_action_batch"), cat("goal_batch")
    for name, key in (("reward", "reward_batch"), ("returns", "return_batch"), ("terminal", "mask_batch"),
                      ("expert_flags", "expert_flag_batch"), ("perturb_flags", "perturb_flag_batch")):
        getattr(mem, name)[:] = cat(key)[:, 0]
    L = 16                                                      # synthetic episodes of 16 transitions
    mem.timestep[:] = np.arange(cap) % L + 1
    mem.episode_map[:] = np.minimum((np.arange(cap) // L) * L + L - 1, cap - 1)
    mem.cur_idx, mem.is_full = 0, True
    mem._mark(0, cap)
    np.random.seed(0)

    class Feed:                                                 # run() indexes batches[i % nb]: draw a fresh minibatch each time
        def __getitem__(self, i):
            d = mem.sample(B)
            d["noise_u"] = devb[i]["noise_u"]
            return d

    run(Feed(), 4, False)
    ms = run(Feed(), args.steps, True, pipelined=True)
    if not gather_bench:
        return dict(value=n_gpus * args.steps / (ms / 1e3), unit=UNIT, ms_per_step=ms / args.steps, store_transitions=cap,
                    note="every rank feeds update_parameters from its own ReplayMemoryB200 store in HBM (max over ranks, aggregate)")
    # the gather alone: 20 launches with distinct pre-uploaded random index sets captured in one CUDA graph (no host time
    # between launches), CUDA events around the replay; store (~200 MB) + outputs (~100 MB) exceed the 126 MB L2
    from gaddpg_b200.capi import current_stream, lib
    o = mem._buffers(B)
    idx = [torch.from_numpy(mem.draw_indices(B).astype(np.int32)).to(o["idx"].device) for _ in range(20)]

    def gathers():
        for t in idx:
            lib.gaddpg_replay_gather(mem.point_state.data_ptr(), C * (N + 6), mem.records.data_ptr(), rm.REC_W, rm.C_TIMESTEP,
                                     mem.episode_map_dev.data_ptr(), cap, t.data_ptr(), B, o["state"].data_ptr(), o["next"].data_ptr(),
                                     o["rec"].data_ptr(), o["inc"].data_ptr(), None, None, current_stream())

    gathers()
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        gathers()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / len(idx)
    row_bytes = 4 * C * (N + 6)
    alg = 2 * B * row_bytes * 2 + B * 128 * 3
    pk = peaks()
    gbs = alg / (us * 1e-6) / 1e9
    return dict(value=args.steps / (ms / 1e3), unit=UNIT, ms_per_step=ms / args.steps, store_transitions=cap,
                gather=dict(us_per_minibatch=us, algorithmic_bytes=alg, achieved=gbs, peak=pk["hbm"], unit="GB/s",
                            frac=gbs / pk["hbm"], bound="hbm", peak_source=pk["src"], includes="cloud gather kernel + record gather kernel (2 launches per minibatch)",
                            traffic=49.0e6, traffic_source="ncu --set full, profiles/r1_ncu_full_v3.md: 43.2 MB read + 5.8 MB written per "
                            "launch at this size; the gathered clouds stay in L2 for the kernels that consume them next"),
                note="update_parameters(ReplayMemoryB200.sample(B)): minibatch assembled on the GPU from a float32 store in HBM; "
                     "no cloud bytes cross PCIe")


def profile(agent, devb, args):
    """Roofline leg: two eager steps (even + odd) with every C-ABI call bracketed by CUDA events on its stream."""
    import torch

    from gaddpg_b200.profiler import KernelProfile

    pk = peaks()
    agent.use_graph = False
    agent.overlap = False  # one stream: per-call CUDA-event times must not overlap each other
    agent.world = None  # rank 0 profiles alone: no collectives in this leg (the other ranks wait at the barrier)
    prof = KernelProfile()
    with prof:
        for i in range(2):
            b = devb[i % len(devb)]
            agent.update_parameters(b, agent.update_step, 0, noise_u=b["noise_u"])
    torch.cuda.synchronize()
    mres = {}
    for g in list(agent._geom_s_s) + list(agent._geom_n_s):   # every staging slot: the two profiled steps use two of them
        for l in g.lv:
            mres[l.M_dev] = int(l.seg_off[-1])
    tab = prof.table(mres)
    total = sum(r["ms"] for r in tab.values())
    rows = sorted(tab.items(), key=lambda kv: -kv[1]["ms"])
    kernels = [dict(entry=k, ms_per_step=r["ms"] / 2, share=r["ms"] / total, calls_per_step=r["calls"] / 2,
                    tflops=(r["flops"] / (r["ms"] * 1e-3) / 1e12) if r["ms"] > 0 else 0.0,
                    gbs=(r["bytes"] / (r["ms"] * 1e-3) / 1e9) if r["ms"] > 0 else 0.0) for k, r in rows[:60]]
    # group the row-GEMM launches by the kernel that ran them (engine.nt asks the library which path a call takes)
    FAM = {"gemm_nt:tc:": "tc_gemm_nt_kernel (tcgen05 3xTF32 row-GEMM, SA1 shared-MLP layers fwd + dX, M ~ 423k rows)",
           "gemm_nt:kc:": "tc_nt_kc_kernel (K-chunked tcgen05 row-GEMM, SA2/SA3 layers)",
           "gemm_nt:sk:": "skinny_nt_kernel (cp.async ring + mma.sync 3xTF32, FC head + actor/critic layers, M = B)",
           "gemm_nt:ffma:": "gemm_nt_kernel (FP32 FFMA row-GEMM fallback)",
           "gemm_tn:": "tc_tn_kernel + tn_reduce (tcgen05 weight gradients)"}
    fam = {k: dict(ms=0.0, flops=0.0, bytes=0.0, calls=0, shapes={}) for k in FAM}
    for k, r in tab.items():
        for f in FAM:
            if k.startswith(f):
                for kk in ("ms", "flops", "bytes"):
                    fam[f][kk] += r[kk]
                fam[f]["calls"] += r["calls"]
                fam[f]["shapes"][k[len(f):]] = dict(us_per_launch=1e3 * r["ms"] / r["calls"], launches_per_step=r["calls"] / 2,
                                                    gbs=r["bytes"] / (r["ms"] * 1e-3) / 1e9, tflops=r["flops"] / (r["ms"] * 1e-3) / 1e12)
    # The dominant KERNEL: every gemm_nt path is exactly one kernel per call, so its event time is that kernel's time.  A
    # gemm_tn call is a composite of 2-3 kernels (tc_tn_kernel split + fixed-order tn_reduce [+ bias_reduce]); in the ncu
    # launch list (profiles/r1_launches_v3_summary.md) tc_tn_kernel alone is 16.6 % of the GPU time against 21.1 % for
    # tc_gemm_nt_kernel, so the composite is reported beside the roofline ("roofline_tn_composite"), not as "the kernel".
    dom = max((f for f in fam if f != "gemm_tn:"), key=lambda f: fam[f]["ms"])
    d = fam[dom]
    tn = fam["gemm_tn:"]
    tn_gbs = tn["bytes"] / (tn["ms"] * 1e-3) / 1e9 if tn["ms"] > 0 else 0.0
    tn_roof = dict(kernel=FAM["gemm_tn:"], bound="hbm", achieved=tn_gbs, peak=pk["hbm"], unit="GB/s", frac=tn_gbs / pk["hbm"],
                   share_of_step=tn["ms"] / total, launches_per_step=tn["calls"] / 2,
                   note="composite of 2-3 kernels per call; the SA1 shapes (M ~ 423k rows) carry the bytes, the M = B layers are "
                        "latency-bound launches with ~no bytes", shapes=tn["shapes"])
    gbs = d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0
    tfl = d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0
    hbm_bound = gbs / pk["hbm"] >= 3.0 * tfl / pk["tensor_sustained"]   # 3xTF32: three tensor passes per algorithmic flop
    # measured DRAM traffic per launch: launch-weighted mean over the shapes of this family that were captured under ncu
    tr = NCU_TRAFFIC.get(dom, {})
    tw = [(tr[k], v["launches_per_step"]) for k, v in d["shapes"].items() if k in tr]
    traffic = sum(b * w for b, w in tw) / sum(w for _, w in tw) if tw else None
    alg_same = sum(v["gbs"] * 1e9 * v["us_per_launch"] * 1e-6 * v["launches_per_step"] for k, v in d["shapes"].items() if k in tr)
    alg_same = alg_same / sum(w for _, w in tw) if tw else None
    if hbm_bound:
        roof = dict(kernel=FAM[dom], bound="hbm", achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"],
                    traffic=traffic, traffic_algorithmic_same_launches=alg_same, peak_source=pk["src"] + ", copy bandwidth",
                    algorithmic_tflops=tfl, tensor_peak_tflops=pk["tensor_sustained"])
    else:
        roof = dict(kernel=FAM[dom], bound="tensor", achieved=tfl, peak=pk["tensor_sustained"], unit="TFLOP/s",
                    frac=tfl / pk["tensor_sustained"], traffic=traffic, peak_source=pk["src"] + ", sustained bf16",
                    algorithmic_gbs=gbs, hbm_peak_gbs=pk["hbm"])
    roof.update(share_of_step=d["ms"] / total, avg_launch_ms=d["ms"] / max(d["calls"], 1), launches_per_step=d["calls"] / 2,
                algorithmic_bytes_per_launch=d["bytes"] / max(d["calls"], 1), shapes=d["shapes"],
                note="achieved = algorithmic bytes (A row [+Y row for BN-backward, + mask-source row] read once, C row written "
                     "once, weights once; DESIGN.md section 4) of all launches of this kernel in one even + one odd step / their "
                     "CUDA-event time in a single-stream eager pass; 'traffic' = dram bytes per launch from the ncu --set full "
                     "capture under profiles/r1_ncu_full_v2.md, launch-weighted over the SA1 shapes of this kernel (null if none was captured); traffic_algorithmic_same_launches = the algorithmic bytes of those same launches")
    families = {FAM[f].split(" ")[0]: dict(ms_per_step=fam[f]["ms"] / 2, share=fam[f]["ms"] / total,
                                           gbs=fam[f]["bytes"] / (fam[f]["ms"] * 1e-3) / 1e9 if fam[f]["ms"] > 0 else 0.0,
                                           tflops=fam[f]["flops"] / (fam[f]["ms"] * 1e-3) / 1e12 if fam[f]["ms"] > 0 else 0.0)
                for f in FAM}
    rows_live = {("state" if g is agent.geom_s else "next") + ".sa%d" % (i + 1): int(l.seg_off[-1]) for g in (agent.geom_s, agent.geom_n)
                 for i, l in enumerate(g.lv)}
    return dict(roofline=roof, roofline_tn_composite=tn_roof, kernel_families=families, kernels=kernels, eager_ms_per_step=total / 2, folded_rows=rows_live,
                dense_rows={"sa1": agent.B * 32 * 64, "sa2": agent.B * 32 * 128})


if __name__ == "__main__":
    main()
