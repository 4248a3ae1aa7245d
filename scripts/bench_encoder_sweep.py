"""BASELINE.json config 5 — PointNet++ encoder-only sweep on one B200: N in {1024, 2048, 4096, 8192} points, B in {64 ... 1024}.

For every (N, B): the value encoder (10 input channels: xyz + mask + 6 broadcast action channels) forward in training
mode (batch statistics, activations kept) and forward + backward (dX and all weight gradients), CUDA-event timed over
`--iters` repetitions after warm-up, geometry (FPS + ball query + row tables) timed separately.  Prints one JSON line per
point and a markdown table; `clouds/s` counts encoder passes over B clouds.  The +6 hand columns are included so that the
network sees exactly N points (networks.py:234-235)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from gaddpg_b200 import engine, nets, synthetic


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, nargs="+", default=[1024, 2048, 4096, 8192])
    ap.add_argument("--batches", type=int, nargs="+", default=[64, 128, 256, 512, 1024])
    ap.add_argument("--iters", type=int, default=10)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(123456)
    enc = nets.make_encoder_params(10)
    ef = engine.EncoderFlat(enc, dev)
    ws = engine.Workspace(dev)
    rows = []
    for N in args.points:
        for B in args.batches:
            batch = synthetic.make_batch(B, N, step=0)
            cloud = torch.from_numpy(np.ascontiguousarray(batch["point_state_batch"], dtype=np.float32)).to(dev)
            action = torch.from_numpy(batch["action_batch"]).to(dev)
            tvec = torch.from_numpy(batch["time_batch"]).to(dev).float()
            geom = engine.Geometry(B, N, dev)
            caps = (geom.lv[0].cap, geom.lv[1].cap)
            ctx = engine.EncoderCtx(B, caps, engine.WIDTHS, dev)
            sc = engine.BwdScratch(B, caps, engine.WIDTHS, dev)
            dfeat = torch.randn(B, 516, device=dev)
            t_geo = timeit(lambda: geom.build(cloud, 6), args.iters)

            def fwd():
                engine.encoder_forward(ws, ef, geom, cloud, 6, 4, action, ctx, time=tvec, train=True)

            def fwdbwd():
                fwd()
                engine.encoder_backward(ws, ef, ctx, sc, want_dw=True, want_dbc=True, dfeat=dfeat)

            t_f = timeit(fwd, args.iters)
            t_fb = timeit(fwdbwd, args.iters)
            r = dict(N=N, B=B, rows_sa1=int(geom.lv[0].seg_off[-1]), geometry_ms=t_geo, fwd_ms=t_f, fwd_bwd_ms=t_fb,
                     fwd_clouds_per_s=B / t_f * 1e3, fwd_bwd_clouds_per_s=B / t_fb * 1e3)
            rows.append(r)
            print(json.dumps(r), flush=True)
            del ctx, sc, geom, cloud
            torch.cuda.empty_cache()
    print("\n| N | B | SA1 rows | geometry ms | fwd ms | fwd+bwd ms | fwd clouds/s | fwd+bwd clouds/s |")
    print("|---:|---:|---:|---:|---:|---:|---:|---:|")
    for r in rows:
        print("| %d | %d | %d | %.3f | %.3f | %.3f | %.0f | %.0f |" % (r["N"], r["B"], r["rows_sa1"], r["geometry_ms"], r["fwd_ms"],
                                                                 r["fwd_bwd_ms"], r["fwd_clouds_per_s"], r["fwd_bwd_clouds_per_s"]))


if __name__ == "__main__":
    main()
