"""GPU: the K-chunked tcgen05 NT kernel and the tcgen05 TN (weight-gradient) kernel against a float64 reference and the
FP32 FFMA kernels: all prologue / epilogue combinations, ragged N / K, device-side row counts, bias sums, rotation."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from types import SimpleNamespace as NS
from gaddpg_b200 import engine
from gaddpg_b200.capi import lib
from gaddpg_b200.engine import nt, nt_problem, tn, op_plain, op_bnrelu, op_bnbwd, OP_PLAIN, OP_BNRELU, OP_BNBWD, EPI_STORE, EPI_DMASK
from gaddpg_b200.structs import STAT_SLOTS

dev = torch.device("cuda")
ws = engine.Workspace(dev)
torch.manual_seed(0)
def bn(C): return NS(scale=torch.rand(C, device=dev) + 0.5, shift=torch.randn(C, device=dev) * 0.1, mean=torch.randn(C, device=dev) * 0.1, rstd=torch.rand(C, device=dev) + 0.5)
def bb(C): return NS(g=torch.rand(C, device=dev) + 0.2, m1=torch.randn(C, device=dev) * 0.01, m2=torch.randn(C, device=dev) * 0.01)
ok = True
which = os.environ.get("WHICH", "nt,tn")
TCL = int(os.environ.get("TCL", "4"))

if "nt" in which:
    for (Mmax, M, N, K) in ((13000, 12973, 128, 132), (8192, 8192, 256, 260), (8192, 8192, 512, 256), (2048, 2048, 1024, 512),
                            (1024, 1024, 512, 1024), (8192, 8000, 260, 256), (1100, 1100, 256, 516), (20000, 17321, 64, 64),
                            (40000, 39000, 128, 64), (8192, 8192, 132, 128), (1300, 1290, 32, 36),
                            # few-hundred-row problems: skinny_nt_kernel (cp.async ring + mma.sync 3xTF32)
                            (256, 256, 512, 1024), (256, 200, 1024, 516), (256, 256, 19, 256), (512, 500, 768, 516),
                            (96, 70, 64, 36), (256, 256, 256, 20)):
        Mdev = torch.tensor([M], dtype=torch.int32, device=dev)
        X = torch.randn(Mmax, K, device=dev); W = torch.randn(N, K, device=dev) * 0.2; rw = torch.rand(Mmax, device=dev) * 3
        D = torch.randn(Mmax, K, device=dev); Yc = torch.randn(Mmax, K, device=dev); Yp = torch.randn(Mmax, N, device=dev)
        b, bK, bbK, bN = bn(K), bn(K), bb(K), bn(N)
        bias = torch.randn(N, device=dev)
        cases = {
            "bnrelu+store+stats": lambda Y: nt([nt_problem(op_bnrelu(X, b), W, K, Y, N, Mmax, Mdev.data_ptr(), N, K, stats=ws.stats, srw=rw)], OP_BNRELU, EPI_STORE),
            "plain+bias+relu": lambda Y: nt([nt_problem(op_plain(X), W, K, Y, N, Mmax, Mdev.data_ptr(), N, K, bias=bias, relu=1)], OP_PLAIN, EPI_STORE),
            "bnbwd+dmask+stats": lambda Y: nt([nt_problem(op_bnbwd(D, Yc, bK, bbK, rw=rw), W, K, Y, N, Mmax, Mdev.data_ptr(), N, K, stats=ws.stats, Yprev=Yp, ldyp=N, pbn=bN)], OP_BNBWD, EPI_DMASK),
            "bnbwd+store": lambda Y: nt([nt_problem(op_bnbwd(D, Yc, bK, bbK, rw=rw), W, K, Y, N, Mmax, Mdev.data_ptr(), N, K)], OP_BNBWD, EPI_STORE),
            "plain+dmask(relu)": lambda Y: nt([nt_problem(op_plain(X), W, K, Y, N, Mmax, Mdev.data_ptr(), N, K, Yprev=Yp, ldyp=N)], OP_PLAIN, EPI_DMASK),
        }
        A1 = torch.relu(X[:M].double() * b.scale.double() + b.shift.double())
        ref = {"bnrelu+store+stats": A1 @ W.double().t(), "plain+bias+relu": torch.relu(X[:M].double() @ W.double().t() + bias.double())}
        dY = bbK.g.double() * (D[:M].double() - rw[:M, None].double() * (bbK.m1.double() + (Yc[:M].double() - bK.mean.double()) * bK.rstd.double() * bbK.m2.double()))
        z = Yp[:M].double() * bN.scale.double() + bN.shift.double()
        ref["bnbwd+dmask+stats"] = (dY @ W.double().t()) * (z > 0)
        ref["bnbwd+store"] = dY @ W.double().t()
        ref["plain+dmask(relu)"] = (X[:M].double() @ W.double().t()) * (Yp[:M].double() > 0)
        for name, fn in cases.items():
            out = {}
            for tc in (0, TCL):
                lib.gaddpg_set_tensor_core(tc)
                Y = torch.full((Mmax, N), 7.0, device=dev)
                ws.stats.fill_(123.0)
                fn(Y)
                torch.cuda.synchronize()
                st = ws.stats[: STAT_SLOTS * 2 * N].view(STAT_SLOTS, 2, N).double().sum(0).clone()
                out[tc] = (Y, st)
            r = ref[name]
            e0 = float((out[0][0][:M].double() - r).abs().max() / r.abs().max())
            e1 = float((out[TCL][0][:M].double() - r).abs().max() / r.abs().max())
            untouched = bool((out[TCL][0][M:] == 7.0).all())
            line = "NT M=%d N=%d K=%d %-20s ffma %.2e tcgen05 %.2e untouched %s" % (M, N, K, name, e0, e1, untouched)
            good = e1 < (5e-6 if K <= 256 else 2e-5) and untouched
            if "stats" in name:
                if name.startswith("bnrelu"):
                    rs = torch.stack([(rw[:M, None].double() * r).sum(0), (rw[:M, None].double() * r * r).sum(0)])
                else:
                    xh = (Yp[:M].double() - bN.mean.double()) * bN.rstd.double()
                    rs = torch.stack([r.sum(0), (r * xh).sum(0)])
                se0 = float((out[0][1] - rs).abs().max() / rs.abs().max()); se1 = float((out[TCL][1] - rs).abs().max() / rs.abs().max())
                line += " stats ffma %.2e tc %.2e" % (se0, se1)
                good = good and se1 < (2e-5 if K <= 256 else 4e-5)
            print(line, "OK" if good else "FAIL")
            ok = ok and good

if "tn" in which:
    # dW[N,K] = sum_r pro1(P)[r,:N]^T pro2(Q)[r,:K]
    for (Mmax, M, N, K, Ktrue, rot, wb) in ((20000, 17321, 128, 64, 64, 0, False), (20000, 20000, 64, 64, 64, 0, False),
                                            (13000, 12973, 128, 132, 131, 3, False), (8192, 8192, 512, 256, 256, 0, False),
                                            (8192, 8000, 256, 260, 259, 3, False), (256, 256, 1024, 512, 512, 0, True),
                                            (256, 256, 512, 1024, 1024, 0, True), (256, 256, 256, 516, 513, 0, True),
                                            (100, 70, 32, 32, 32, 0, True), (50000, 50000, 256, 128, 128, 0, False)):
        Mdev = torch.tensor([M], dtype=torch.int32, device=dev)
        D = torch.randn(Mmax, N, device=dev); Yc = torch.randn(Mmax, N, device=dev); rw = torch.rand(Mmax, device=dev) * 3
        Xq = torch.randn(Mmax, K, device=dev)
        bNn, bbN, bKk = bn(N), bb(N), bn(K)
        dYd = bbN.g.double() * (D[:M].double() - rw[:M, None].double() * (bbN.m1.double() + (Yc[:M].double() - bNn.mean.double()) * bNn.rstd.double() * bbN.m2.double()))
        Qrelu = torch.relu(Xq[:M].double() * bKk.scale.double() + bKk.shift.double())
        modes = {
            "bnbwd x bnrelu": (op_bnbwd(D, Yc, bNn, bbN, rw=rw), op_bnrelu(Xq, bKk), OP_BNBWD, OP_BNRELU, dYd, Qrelu),
            "bnbwd x plain": (op_bnbwd(D, Yc, bNn, bbN, rw=rw), op_plain(Xq), OP_BNBWD, OP_PLAIN, dYd, Xq[:M].double()),
            "plain x plain": (op_plain(D), op_plain(Xq), OP_PLAIN, OP_PLAIN, D[:M].double(), Xq[:M].double()),
            "plain x bnrelu": (op_plain(D), op_bnrelu(Xq, bKk), OP_PLAIN, OP_BNRELU, D[:M].double(), Qrelu),
        }
        for name, (P, Q, pm, qm, Pd, Qd) in modes.items():
            full = Pd.t() @ Qd                      # [N, K]
            refW = torch.roll(full[:, :Ktrue], rot, dims=1) if rot else full[:, :Ktrue]
            refb = Pd.sum(0)
            res = {}
            for tc in (0, TCL):
                lib.gaddpg_set_tensor_core(tc)
                dW = torch.full((N, Ktrue), 5.0, device=dev)
                db = torch.full((N,), 5.0, device=dev)
                tn(ws, P, Q, pm, qm, Mmax, Mdev.data_ptr(), N, K, dW, Ktrue, N, Ktrue, rot=rot, dbias=db if wb else None)
                torch.cuda.synchronize()
                res[tc] = (dW.clone(), db.clone())
            e0 = float((res[0][0].double() - refW).abs().max() / refW.abs().max())
            e1 = float((res[TCL][0].double() - refW).abs().max() / refW.abs().max())
            line = "TN M=%d N=%d K=%d(%d,rot%d) %-15s ffma %.2e tcgen05 %.2e" % (M, N, K, Ktrue, rot, name, e0, e1)
            good = e1 < 1e-5
            if wb:
                b1 = float((res[TCL][1].double() - refb).abs().max() / refb.abs().max())
                line += " bias %.2e" % b1
                good = good and b1 < 5e-6
            print(line, "OK" if good else "FAIL")
            ok = ok and good
if "pool" in which or which == "nt,tn":
    # sparse max-pool backward: GADDPG_OP_BNBWD_POOL operands (NT K=128 -> N=64 with mask epilogue, TN 128x64) and
    # gaddpg_pool_bwd_sparse against the dense pool_bwd path — same arithmetic, so the GEMM outputs must be bit-identical
    from gaddpg_b200.structs import op_bnbwd_pool, OP_BNBWD_POOL, dp
    lib.gaddpg_set_tensor_core(3)
    C, Kout = 128, 64
    lens = torch.randint(1, 65, (700,))
    Mtrue = int(lens.sum()); Mmax = Mtrue + 300; S = lens.numel()
    seg_off = torch.zeros(S + 1, dtype=torch.int32); seg_off[1:] = torch.cumsum(lens, 0)
    row_seg = torch.repeat_interleave(torch.arange(S, dtype=torch.int32), lens)
    row_seg = torch.cat([row_seg, torch.zeros(Mmax - Mtrue, dtype=torch.int32)]).to(dev)
    seg_off = seg_off.to(dev)
    Mdev = torch.tensor([Mtrue], dtype=torch.int32, device=dev)
    Y2 = torch.randn(Mmax, C, device=dev); Y1 = torch.randn(Mmax, Kout, device=dev); rw = (torch.rand(Mmax, device=dev) * 3).round() + 1
    b2, bb2, b1 = bn(C), bb(C), bn(Kout)
    outp = torch.zeros(S, C, device=dev); arg = torch.zeros(S, C, dtype=torch.int32, device=dev)
    lib.gaddpg_pool_fwd(dp(Y2), C, dp(b2.scale), dp(b2.shift), dp(seg_off), 0, S, dp(outp), dp(arg), 0)
    dOut = torch.randn(S, C + 4, device=dev)
    Dd = torch.zeros(Mmax, C, device=dev); E = torch.zeros(S, C, device=dev); mask = torch.full((Mmax, C // 32), -1, dtype=torch.int32, device=dev)
    lib.gaddpg_pool_bwd(dp(dOut), C + 4, dp(outp), dp(arg), dp(Y2), C, dp(row_seg), 0, Mmax, Mdev.data_ptr(), dp(b2.mean), dp(b2.rstd), dp(Dd), dp(ws.stats), None, 0)
    torch.cuda.synchronize()
    st_dense = ws.stats[: STAT_SLOTS * 2 * C].view(STAT_SLOTS, 2, C).double().sum(0).clone()
    lib.gaddpg_pool_bwd_sparse(dp(dOut), C + 4, dp(outp), dp(arg), dp(Y2), C, S, dp(b2.mean), dp(b2.rstd), dp(E), dp(mask), Mmax, dp(ws.stats), None, 0)
    torch.cuda.synchronize()
    st_sparse = ws.stats[: STAT_SLOTS * 2 * C].view(STAT_SLOTS, 2, C).double().sum(0).clone()
    Dre = torch.zeros(Mmax, C, device=dev)
    Dre.scatter_(0, arg.long(), E)   # D[arg[s][c]][c] = E[s][c]
    good = bool(torch.equal(Dre[:Mtrue], Dd[:Mtrue])) and float((st_dense - st_sparse).abs().max() / st_dense.abs().max()) < 1e-5
    print("pool_bwd_sparse: D rebuilt == dense %s, stats rel diff %.2e" % (torch.equal(Dre[:Mtrue], Dd[:Mtrue]), float((st_dense - st_sparse).abs().max() / st_dense.abs().max())), "OK" if good else "FAIL")
    ok = ok and good
    W = torch.randn(Kout, C, device=dev) * 0.2
    res = {}
    for name, A, mode in (("dense", op_bnbwd(Dd, Y2, b2, bb2, rw=rw), OP_BNBWD), ("pool", op_bnbwd_pool(E, mask, row_seg, Y2, b2, bb2, rw=rw), OP_BNBWD_POOL)):
        Yo = torch.full((Mmax, Kout), 7.0, device=dev); ws.stats.fill_(5.0)
        nt([nt_problem(A, W, C, Yo, Kout, Mmax, Mdev.data_ptr(), Kout, C, stats=ws.stats, Yprev=Y1, ldyp=Kout, pbn=b1)], mode, EPI_DMASK)
        torch.cuda.synchronize()
        stn = ws.stats[: STAT_SLOTS * 2 * Kout].clone()
        dW = torch.full((C, Kout), 3.0, device=dev)
        tn(ws, A, op_bnrelu(Y1, b1), mode, OP_BNRELU, Mmax, Mdev.data_ptr(), C, Kout, dW, Kout, C, Kout)
        torch.cuda.synchronize()
        res[name] = (Yo, stn, dW.clone())
    for i, what in enumerate(("NT dX", "NT stats", "TN dW")):
        same = bool(torch.equal(res["dense"][i], res["pool"][i]))
        print("BNBWD_POOL vs dense BNBWD: %s bit-identical %s" % (what, same), "OK" if same else "FAIL")
        ok = ok and same
if "pool" in which or which == "nt,tn":
    # fused max-pool: per-segment max / min keys from the NT epilogue + gaddpg_pool_keys_finalize against gaddpg_pool_fwd on
    # the stored layer output (scales of both signs); the pooled values must be bit-identical, arg must point at a row that
    # attains them; no_store must leave C untouched.  Shapes: SA1 (K=64 -> N=128, resident-weight kernel) and SA2 (K=128 ->
    # N=256, K-chunked kernel).
    for (Kin, Nout) in ((64, 128), (128, 256)):
        lens = torch.randint(1, 65, (900,))
        Mtrue = int(lens.sum()); Mmax = max(Mtrue + 200, 8192); S = lens.numel()
        seg_off = torch.zeros(S + 1, dtype=torch.int32); seg_off[1:] = torch.cumsum(lens, 0)
        row_seg = torch.cat([torch.repeat_interleave(torch.arange(S, dtype=torch.int32), lens), torch.zeros(Mmax - Mtrue, dtype=torch.int32)]).to(dev)
        seg_off = seg_off.to(dev)
        Mdev = torch.tensor([Mtrue], dtype=torch.int32, device=dev)
        X = torch.randn(Mmax, Kin, device=dev); W = torch.randn(Nout, Kin, device=dev) * 0.2; rwp = torch.ones(Mmax, device=dev)
        bin_, bout = bn(Kin), bn(Nout)
        bout.scale = bout.scale * torch.where(torch.rand(Nout, device=dev) < 0.4, -1.0, 1.0)   # negative BatchNorm scales too
        keys = torch.zeros(S * Nout, dtype=torch.int64, device=dev)
        gam = torch.where(bout.scale < 0, -1.0, 1.0) * (torch.rand(Nout, device=dev) + 0.5)   # the BatchNorm weight: same sign as the scale
        Yf = torch.full((Mmax, Nout), 7.0, device=dev)
        nt([nt_problem(op_bnrelu(X, bin_), W, Kin, Yf, Nout, Mmax, Mdev.data_ptr(), Nout, Kin, stats=ws.stats, srw=rwp,
                       pool_keys=keys, pool_seg=row_seg, pool_gamma=gam)], OP_BNRELU, EPI_STORE)
        out_f = torch.zeros(S, Nout, device=dev); arg_f = torch.zeros(S, Nout, dtype=torch.int32, device=dev)
        lib.gaddpg_pool_keys_finalize(dp(keys), S, Nout, dp(gam), dp(bout.scale), dp(bout.shift), dp(out_f), dp(arg_f), 0)
        out_r = torch.zeros(S, Nout, device=dev); arg_r = torch.zeros(S, Nout, dtype=torch.int32, device=dev)
        lib.gaddpg_pool_fwd(dp(Yf), Nout, dp(bout.scale), dp(bout.shift), dp(seg_off), 0, S, dp(out_r), dp(arg_r), 0)
        Yn = torch.full((Mmax, Nout), 7.0, device=dev)
        nt([nt_problem(op_bnrelu(X, bin_), W, Kin, Yn, Nout, Mmax, Mdev.data_ptr(), Nout, Kin, stats=ws.stats, srw=rwp,
                       pool_keys=keys, pool_seg=row_seg, pool_gamma=gam, no_store=1)], OP_BNRELU, EPI_STORE)
        out_n = torch.zeros(S, Nout, device=dev)
        lib.gaddpg_pool_keys_finalize(dp(keys), S, Nout, dp(gam), dp(bout.scale), dp(bout.shift), dp(out_n), None, 0)
        torch.cuda.synchronize()
        at_arg = torch.relu(torch.gather(Yf, 0, arg_f.long()) * bout.scale + bout.shift)
        seg_ok = bool((torch.gather(row_seg[:, None].expand(-1, Nout), 0, arg_f.long()) == torch.arange(S, device=dev)[:, None]).all())
        good = (torch.equal(out_f, out_r) and torch.equal(out_n, out_r) and bool((Yn == 7.0).all()) and bool((keys == 0).all())
                and seg_ok and float((at_arg - out_f).abs().max()) <= 1e-6 * float(out_f.abs().max()))
        print("fused max-pool K=%d N=%d: out == pool_fwd %s, no_store out %s / C untouched %s, keys re-zeroed %s, arg in segment %s, bn(Y[arg]) - out %.1e"
              % (Kin, Nout, torch.equal(out_f, out_r), torch.equal(out_n, out_r), bool((Yn == 7.0).all()), bool((keys == 0).all()), seg_ok,
                 float((at_arg - out_f).abs().max())), "OK" if good else "FAIL")
        ok = ok and good
lib.gaddpg_set_tensor_core(3)
print("ALL OK" if ok else "SOME FAILED")
