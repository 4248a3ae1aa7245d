"""Debug: per-role cycle accounting of tc_gemm_nt_kernel (build variant -DTC_PROFILE, see csrc/tc_gemm.cu)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from types import SimpleNamespace as NS
from gaddpg_b200 import engine
from gaddpg_b200.capi import lib
from gaddpg_b200.engine import nt, nt_problem, op_bnrelu, op_bnbwd, OP_BNRELU, OP_BNBWD, EPI_STORE, EPI_DMASK
dev = torch.device("cuda"); ws = engine.Workspace(dev); M = 423608
raw = lib.load()
raw.gaddpg_debug_tc_prof.argtypes = [ctypes.c_void_p, ctypes.c_int]
def bn(C): return NS(scale=torch.rand(C, device=dev) + 0.5, shift=torch.randn(C, device=dev) * 0.1, mean=torch.randn(C, device=dev) * 0.1, rstd=torch.rand(C, device=dev) + 0.5)
def bb(C): return NS(g=torch.rand(C, device=dev), m1=torch.randn(C, device=dev) * 0.01, m2=torch.randn(C, device=dev) * 0.01)
rw = torch.ones(M, device=dev)
names = ["prod wait a_empty", "prod total", "mma wait acc_empty", "mma wait a_full", "mma total", "epi wait acc_full", "epi work", "tiles"]
for (N, K, mode) in ((64, 64, "fwd"), (128, 64, "fwd"), (64, 128, "bwd"), (64, 64, "bwd")):
    X = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * 0.1; Y = torch.empty(M, N, device=dev)
    D = torch.randn(M, K, device=dev); Yc = torch.randn(M, K, device=dev); Yp = torch.randn(M, N, device=dev)
    b, bK, bbK, bN = bn(K), bn(K), bb(K), bn(N)
    if mode == "fwd":
        fn = lambda: nt([nt_problem(op_bnrelu(X, b), W, K, Y, N, M, None, N, K, stats=ws.stats, srw=rw)], OP_BNRELU, EPI_STORE)
    else:
        fn = lambda: nt([nt_problem(op_bnbwd(D, Yc, bK, bbK, rw=rw), W, K, Y, N, M, None, N, K, stats=ws.stats, Yprev=Yp, ldyp=N, pbn=bN)], OP_BNBWD, EPI_DMASK)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    raw.gaddpg_debug_tc_prof(None, 1)
    fn(); torch.cuda.synchronize()
    out = np.zeros(148 * 16, dtype=np.uint64)
    raw.gaddpg_debug_tc_prof(out.ctypes.data, 0)
    o = out.reshape(148, 16).astype(np.float64)
    tiles = o[:, 7].mean()
    print("%s N=%d K=%d: tiles/CTA %.1f" % (mode, N, K, tiles))
    for i, n in enumerate(names[:7]):
        print("   %-20s %9.0f cycles/CTA  = %7.0f per tile" % (n, o[:, i].mean(), o[:, i].mean() / max(tiles, 1)))
    nb = max(o[:, 11].mean(), 1)
    print("   per 32x32 block (fast path, warp 9): ldtm+wait %.0f, sts+sync+lds+sync %.0f, math+stores %.0f cycles (%.0f blocks/CTA)" % (
        o[:, 8].mean() / nb, o[:, 9].mean() / nb, o[:, 10].mean() / nb, nb))
