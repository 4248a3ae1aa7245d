"""Graph-timed SA1 weight-gradient launches in the form the update step uses (sparse max-pool operand, GADDPG_OP_BNBWD_POOL)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace as NS
from gaddpg_b200 import engine
from gaddpg_b200.capi import lib
from gaddpg_b200.structs import dp
from gaddpg_b200.engine import nt, nt_problem, tn, op_bnrelu, op_bnbwd, op_bnbwd_pool, OP_BNRELU, OP_BNBWD, OP_BNBWD_POOL, EPI_DMASK

dev = torch.device("cuda")
ws = engine.Workspace(dev)
torch.manual_seed(0)
S, C = 8192, 128
cnt = torch.randint(20, 64, (S,), device=dev)
seg_off = torch.zeros(S + 1, dtype=torch.int32, device=dev)
seg_off[1:] = torch.cumsum(cnt, 0).int()
M = int(seg_off[-1])
row_seg = torch.repeat_interleave(torch.arange(S, device=dev, dtype=torch.int32), cnt.long()).contiguous()
def bn(C): return NS(scale=torch.rand(C, device=dev) + 0.5, shift=torch.randn(C, device=dev) * 0.1, mean=torch.randn(C, device=dev) * 0.1, rstd=torch.rand(C, device=dev) + 0.5)
def bb(C): return NS(g=torch.rand(C, device=dev), m1=torch.randn(C, device=dev) * 0.01, m2=torch.randn(C, device=dev) * 0.01)
Y2 = torch.randn(M, 128, device=dev); Y1 = torch.randn(M, 64, device=dev); rw = torch.ones(M, device=dev)
E = torch.randn(S, 128, device=dev)
mask = torch.randint(0, 2**31 - 1, (M, 4), device=dev, dtype=torch.int32)
b2, bb2, b1 = bn(128), bb(128), bn(64)
W = torch.randn(128, 64, device=dev) * 0.1
WT = W.t().contiguous()
dW = torch.empty(128, 64, device=dev); D1 = torch.empty(M, 64, device=dev)
dy = op_bnbwd_pool(E, mask, row_seg, Y2, b2, bb2, rw=rw)
def timeit(fn, n=20):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph(); s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); s.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n): fn()
    torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * n) * 1e3
print("M = %d" % M)
print("tn dW 128x64 pool operand   %.1f us" % timeit(lambda: tn(ws, dy, op_bnrelu(Y1, b1), OP_BNBWD_POOL, OP_BNRELU, M, None, 128, 64, dW, 64, 128, 64)))
print("nt dX 128->64 pool operand  %.1f us" % timeit(lambda: nt([nt_problem(dy, WT, 128, D1, 64, M, None, 64, 128, stats=ws.stats, Yprev=Y1, ldyp=64, pbn=b1)], OP_BNBWD_POOL, EPI_DMASK)))
