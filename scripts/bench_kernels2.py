"""Micro-benchmarks (GPU box): FFMA vs tcgen05 row-GEMMs at the shapes of one cfg2 update step (SA1 / SA2 / SA3 / FC).
TC levels: 0 FFMA, 1 resident-weight tcgen05 (SA1 only), 3 default (+K-chunked NT, +TN), 4 K-chunked NT everywhere."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace as NS
from gaddpg_b200 import engine
from gaddpg_b200.capi import lib
from gaddpg_b200.engine import nt, nt_problem, tn, op_plain, op_bnrelu, op_bnbwd, OP_PLAIN, OP_BNRELU, OP_BNBWD, EPI_STORE, EPI_DMASK

dev = torch.device("cuda")
ws = engine.Workspace(dev)
torch.manual_seed(0)
def bn(C): return NS(scale=torch.rand(C, device=dev) + 0.5, shift=torch.randn(C, device=dev) * 0.1, mean=torch.randn(C, device=dev) * 0.1, rstd=torch.rand(C, device=dev) + 0.5)
def bb(C): return NS(g=torch.rand(C, device=dev), m1=torch.randn(C, device=dev) * 0.01, m2=torch.randn(C, device=dev) * 0.01)
def timeit(fn, n=20):
    """GPU time per launch: n launches captured into one CUDA graph (no Python / launch gaps), replayed 3 times."""
    for _ in range(2): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        s.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n): fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * n)
levels = [int(x) for x in os.environ.get("LEVELS", "0,3,4").split(",")]
shapes = [(423608, 64, 64), (423608, 128, 64), (12974, 128, 132), (12974, 128, 128), (12974, 256, 128), (8192, 256, 260),
          (8192, 256, 256), (8192, 512, 256), (256, 1024, 512), (256, 512, 1024)]
print("%-44s" % "kernel (M, N, K)" + "".join("  L%d ms   " % l for l in levels))
for (M, N, K) in shapes:
    rw = torch.ones(M, device=dev)
    X = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.1; Y = torch.empty(M, N, device=dev)
    WT = W.t().contiguous()
    b = bn(K)
    D = torch.randn(M, N, device=dev); Yc = torch.randn(M, N, device=dev); dX = torch.empty(M, K, device=dev)
    bN_, bbN, bK_ = bn(N), bb(N), bn(K)
    dW = torch.empty(N, K, device=dev)
    rows = {
        "nt fwd bnrelu+stats": lambda: nt([nt_problem(op_bnrelu(X, b), W, K, Y, N, M, None, N, K, stats=ws.stats, srw=rw)], OP_BNRELU, EPI_STORE),
        "nt dX  bnbwd+dmask ": lambda: nt([nt_problem(op_bnbwd(D, Yc, bN_, bbN, rw=rw), WT, N, dX, K, M, None, K, N, stats=ws.stats, Yprev=X, ldyp=K, pbn=bK_)], OP_BNBWD, EPI_DMASK),
        "tn dW  bnbwd x bnrelu": lambda: tn(ws, op_bnbwd(D, Yc, bN_, bbN, rw=rw), op_bnrelu(X, bK_), OP_BNBWD, OP_BNRELU, M, None, N, K, dW, K, N, K),
    }
    for name, fn in rows.items():
        line = "%-22s (%6d,%4d,%4d)  " % (name, M, N, K)
        for l in levels:
            lib.gaddpg_set_tensor_core(l)
            line += " %8.4f " % timeit(fn)
        print(line)
lib.gaddpg_set_tensor_core(3)
