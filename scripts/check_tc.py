"""GPU: tcgen05 3xTF32 row-GEMM vs the FP32 FFMA kernel and a float64 reference (values + BatchNorm statistics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace as NS
from gaddpg_b200 import engine
from gaddpg_b200.capi import lib
from gaddpg_b200.engine import nt, nt_problem, op_plain, op_bnrelu, op_bnbwd, OP_PLAIN, OP_BNRELU, OP_BNBWD, EPI_STORE, EPI_DMASK
from gaddpg_b200.structs import STAT_SLOTS

dev = torch.device("cuda")
ws = engine.Workspace(dev)
torch.manual_seed(0)
def bn(C): return NS(scale=torch.rand(C, device=dev) + 0.5, shift=torch.randn(C, device=dev) * 0.1, mean=torch.randn(C, device=dev) * 0.1, rstd=torch.rand(C, device=dev) + 0.5)
def bb(C): return NS(g=torch.rand(C, device=dev) + 0.2, m1=torch.randn(C, device=dev) * 0.01, m2=torch.randn(C, device=dev) * 0.01)
ok = True
for (Mmax, M, N, K) in ((20000, 17321, 64, 64), (20000, 20000, 128, 64), (16384, 9000, 64, 128), (8192, 8192, 128, 128), (70000, 65537, 256, 32)):
    Mdev = torch.tensor([M], dtype=torch.int32, device=dev)
    X = torch.randn(Mmax, K, device=dev); W = torch.randn(N, K, device=dev) * 0.2; rw = torch.rand(Mmax, device=dev) * 3
    D = torch.randn(Mmax, K, device=dev); Yc = torch.randn(Mmax, K, device=dev); Yp = torch.randn(Mmax, N, device=dev)
    b, bK, bbK, bN = bn(K), bn(K), bb(K), bn(N)
    bias = torch.randn(N, device=dev)
    cases = {
        "bnrelu+store+stats": lambda Y: nt([nt_problem(op_bnrelu(X, b), W, K, Y, N, Mmax, Mdev.data_ptr(), N, K, stats=ws.stats, srw=rw)], OP_BNRELU, EPI_STORE),
        "plain+bias+relu": lambda Y: nt([nt_problem(op_plain(X), W, K, Y, N, Mmax, Mdev.data_ptr(), N, K, bias=bias, relu=1)], OP_PLAIN, EPI_STORE),
        "bnbwd+dmask+stats": lambda Y: nt([nt_problem(op_bnbwd(D, Yc, bK, bbK, rw=rw), W, K, Y, N, Mmax, Mdev.data_ptr(), N, K, stats=ws.stats, Yprev=Yp, ldyp=N, pbn=bN)], OP_BNBWD, EPI_DMASK),
    }
    # float64 references
    A1 = torch.relu(X[:M].double() * b.scale.double() + b.shift.double())
    ref = {"bnrelu+store+stats": A1 @ W.double().t(), "plain+bias+relu": torch.relu(X[:M].double() @ W.double().t() + bias.double())}
    dY = bbK.g.double() * (D[:M].double() - rw[:M, None].double() * (bbK.m1.double() + (Yc[:M].double() - bK.mean.double()) * bK.rstd.double() * bbK.m2.double()))
    z = Yp[:M].double() * bN.scale.double() + bN.shift.double()
    ref["bnbwd+dmask+stats"] = (dY @ W.double().t()) * (z > 0)
    for name, fn in cases.items():
        out = {}
        for tc in (0, 1):
            lib.gaddpg_set_tensor_core(tc)
            Y = torch.full((Mmax, N), 7.0, device=dev)
            ws.stats.fill_(123.0)
            fn(Y)
            torch.cuda.synchronize()
            st = ws.stats[: STAT_SLOTS * 2 * N].view(STAT_SLOTS, 2, N).double().sum(0).clone()
            out[tc] = (Y, st)
        r = ref[name]
        e0 = float((out[0][0][:M].double() - r).abs().max() / r.abs().max())
        e1 = float((out[1][0][:M].double() - r).abs().max() / r.abs().max())
        untouched = bool((out[1][0][M:] == 7.0).all())
        line = "M=%d N=%d K=%d %-20s  ffma err %.2e  tcgen05 err %.2e  rows>=M untouched %s" % (M, N, K, name, e0, e1, untouched)
        good = e1 < 5e-6 and untouched
        if "stats" in name:
            if name.startswith("bnrelu"):
                rs = torch.stack([(rw[:M, None].double() * r).sum(0), (rw[:M, None].double() * r * r).sum(0)])
            else:
                xh = (Yp[:M].double() - bN.mean.double()) * bN.rstd.double()
                rs = torch.stack([r.sum(0), (r * xh).sum(0)])
            se0 = float((out[0][1] - rs).abs().max() / rs.abs().max()); se1 = float((out[1][1] - rs).abs().max() / rs.abs().max())
            line += "  stats err ffma %.2e tc %.2e" % (se0, se1)
            good = good and se1 < 2e-5
        print(line, "OK" if good else "FAIL")
        ok = ok and good
print("ALL OK" if ok else "SOME FAILED")
