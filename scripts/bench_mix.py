"""Achievable HBM throughput of plain streaming kernels for the read:write mixes of the SA1 row-GEMMs (M = 423 k rows):
what a kernel with NO arithmetic and perfect coalescing reaches on this GPU for the same traffic.  torch elementwise
kernels only (measurement aid, not product code)."""
import torch

M = 423608
dev = "cuda"
x64 = torch.randn(M, 64, device=dev)
y64 = torch.randn(M, 64, device=dev)
z64 = torch.randn(M, 64, device=dev)
x128 = torch.randn(M, 128, device=dev)
o64 = torch.empty(M, 64, device=dev)
o128 = torch.empty(M, 128, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)


def timeit(fn, nbytes, name, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    t = ts[len(ts) // 2]
    print("%-44s %7.1f us  %6.0f GB/s" % (name, t, nbytes / t / 1e3))


MB64, MB128 = M * 64 * 4, M * 128 * 4
timeit(lambda: o64.copy_(x64), 2 * MB64, "copy 64 -> 64 (1r:1w, fwd 64->64)")
timeit(lambda: torch.cat([x64, x64], 1, out=o128), MB64 + MB128, "64 -> 128 (1r:2w, fwd 64->128)")
timeit(lambda: torch.add(x64, y64, out=o64), 3 * MB64, "64 + 64 -> 64 (2r:1w, dX 64->64)")
timeit(lambda: torch.add(x128[:, :64], x128[:, 64:], out=o64), MB128 + MB64, "128 -> 64 (2r:1w)")
timeit(lambda: torch.addcmul(z64, x128[:, :64], x128[:, 64:], out=o64), MB128 + 2 * MB64, "128 + 64 -> 64 (3r:1w, dX 128->64)")
timeit(lambda: x128.sum(), MB128, "read 128 (pool_fwd)")
timeit(lambda: (x128.sum(), x64.sum()), MB128 + MB64, "read 128 + 64 (dW 128x64)")
timeit(lambda: o128.zero_(), MB128, "write 128")
timeit(lambda: o64.zero_(), MB64, "write 64 (sa1_l1_fwd)")
