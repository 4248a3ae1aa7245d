"""Micro-benchmarks (GPU box) of the row-GEMM kernels at the SA1 shapes that dominate the step (M ~ 420k compact rows)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace as NS
from gaddpg_b200 import engine
from gaddpg_b200.engine import nt, nt_problem, tn, op_plain, op_bnrelu, op_bnbwd, OP_PLAIN, OP_BNRELU, OP_BNBWD, EPI_STORE, EPI_DMASK

dev = torch.device("cuda")
ws = engine.Workspace(dev)
M = int(os.environ.get("M", 423608))
torch.manual_seed(0)
def bn(C): return NS(scale=torch.rand(C, device=dev) + 0.5, shift=torch.randn(C, device=dev) * 0.1, mean=torch.randn(C, device=dev) * 0.1, rstd=torch.rand(C, device=dev) + 0.5)
def bb(C): return NS(g=torch.rand(C, device=dev), m1=torch.randn(C, device=dev) * 0.01, m2=torch.randn(C, device=dev) * 0.01)
rw = torch.ones(M, device=dev)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
res = []
for (N, K) in ((64, 64), (128, 64), (64, 128), (128, 128), (256, 128)):
    X = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.1; Y = torch.empty(M, N, device=dev)
    b = bn(K)
    ms = timeit(lambda: nt([nt_problem(op_bnrelu(X, b), W, K, Y, N, M, None, N, K, stats=ws.stats, srw=rw)], OP_BNRELU, EPI_STORE))
    res.append(("nt fwd bnrelu+stats N=%d K=%d" % (N, K), ms, 2.0 * M * N * K / ms / 1e9, 4.0 * M * (N + K) / ms / 1e6))
    # backward dX with mask epilogue: A = BNBWD(D,Yc) [M,K], out [M,N] masked by Yprev [M,N]
    D = torch.randn(M, K, device=dev); Yc = torch.randn(M, K, device=dev); Yp = torch.randn(M, N, device=dev)
    bK, bbK, bN = bn(K), bb(K), bn(N)
    ms = timeit(lambda: nt([nt_problem(op_bnbwd(D, Yc, bK, bbK, rw=rw), W, K, Y, N, M, None, N, K, stats=ws.stats, Yprev=Yp, ldyp=N, pbn=bN)], OP_BNBWD, EPI_DMASK))
    res.append(("nt bwd bnbwd+dmask  N=%d K=%d" % (N, K), ms, 2.0 * M * N * K / ms / 1e9, 4.0 * M * (2 * K + 2 * N) / ms / 1e6))
    dW = torch.empty(K, N, device=dev)  # TN: P [M,K'] x Q [M,N'] -> here P=D-like (N_out=K), Q = activations (K_out=N)
    ms = timeit(lambda: tn(ws, op_bnbwd(D, Yc, bK, bbK, rw=rw), op_bnrelu(Yp, bN), OP_BNBWD, OP_BNRELU, M, None, K, N, dW, N, K, N))
    res.append(("tn dW bnbwd x bnrelu  [%dx%d]" % (K, N), ms, 2.0 * M * N * K / ms / 1e9, 4.0 * M * (2 * K + N) / ms / 1e6))
for r in res:
    print("%-36s %7.3f ms  %7.2f TFLOP/s  %7.1f GB/s" % r)
