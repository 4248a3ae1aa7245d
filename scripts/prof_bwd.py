"""ncu target: the SA1 backward row-GEMMs at cfg2 size (dX 128->64, dX 64->64, dW 128x64, dW 64x64), one launch each after a warm-up."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace as NS
from gaddpg_b200 import engine
from gaddpg_b200.engine import nt, nt_problem, tn, op_bnrelu, op_bnbwd, OP_BNRELU, OP_BNBWD, EPI_DMASK

dev = torch.device("cuda")
ws = engine.Workspace(dev)
torch.manual_seed(0)
M = int(os.environ.get("M", 423608))
def bn(C): return NS(scale=torch.rand(C, device=dev) + 0.5, shift=torch.randn(C, device=dev) * 0.1, mean=torch.randn(C, device=dev) * 0.1, rstd=torch.rand(C, device=dev) + 0.5)
def bb(C): return NS(g=torch.rand(C, device=dev), m1=torch.randn(C, device=dev) * 0.01, m2=torch.randn(C, device=dev) * 0.01)
rw = torch.ones(M, device=dev)
for rep in range(int(os.environ.get("REPS", 2))):
    for (N, K) in ((128, 64), (64, 64)):
        X = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.1
        WT = W.t().contiguous()
        D = torch.randn(M, N, device=dev); Yc = torch.randn(M, N, device=dev); dX = torch.empty(M, K, device=dev)
        bN_, bbN, bK_ = bn(N), bb(N), bn(K)
        dW = torch.empty(N, K, device=dev)
        nt([nt_problem(op_bnbwd(D, Yc, bN_, bbN, rw=rw), WT, N, dX, K, M, None, K, N, stats=ws.stats, Yprev=X, ldyp=K, pbn=bK_)], OP_BNBWD, EPI_DMASK)
        tn(ws, op_bnbwd(D, Yc, bN_, bbN, rw=rw), op_bnrelu(X, bK_), OP_BNBWD, OP_BNRELU, M, None, N, K, dW, K, N, K)
        torch.cuda.synchronize()
