"""ctypes mirrors of the descriptor structs in include/gaddpg_b200.h, plus small builders.

``check_sizes()`` compares ``ctypes.sizeof`` with the library's own ``sizeof`` so a header edit that is not
mirrored here fails at import on the GPU box instead of corrupting a launch.
"""
import ctypes

from .capi import lib

OP_PLAIN, OP_BNRELU, OP_BNBWD, OP_BNBWD_POOL = 0, 1, 2, 3
EPI_STORE, EPI_DMASK = 0, 1
STAT_SLOTS = 296
MAX_GROUP = 4

_P = ctypes.c_void_p
_I = ctypes.c_int


class Operand(ctypes.Structure):
    _fields_ = [("X", _P), ("ldx", _I), ("Y", _P), ("ldy", _I), ("rw", _P),
                ("c0", _P), ("c1", _P), ("c2", _P), ("c3", _P), ("c4", _P), ("pmask", _P), ("pseg", _P)]


class BNTail(ctypes.Structure):
    """gaddpg_bn_tail: BatchNorm finalize run by the last CTA of the kernel that produced the statistics."""
    _fields_ = [("kind", _I), ("accumulate", _I), ("count", ctypes.c_double), ("a", _P), ("b", _P),
                ("eps", ctypes.c_float), ("momentum", ctypes.c_float), ("running_mean", _P), ("running_var", _P),
                ("num_batches_tracked", _P), ("o0", _P), ("o1", _P), ("o2", _P), ("o3", _P), ("dgamma", _P), ("dbeta", _P),
                ("counter", _P)]


class NTProblem(ctypes.Structure):
    _fields_ = [("A", Operand), ("Bw", _P), ("ldb", _I), ("bias", _P), ("C", _P), ("ldc", _I),
                ("M_max", _I), ("M_dev", _P), ("N", _I), ("K", _I), ("relu", _I),
                ("stats", _P), ("srw", _P), ("Yprev", _P), ("ldyp", _I),
                ("psc", _P), ("psh", _P), ("pmean", _P), ("prstd", _P),
                ("pool_keys", _P), ("pool_seg", _P), ("pool_gamma", _P), ("no_store", _I), ("Bw_hi", _P), ("Bw_lo", _P),
                ("tail", BNTail)]


class NTGroup(ctypes.Structure):
    _fields_ = [("p", NTProblem * MAX_GROUP)]


class TNProblem(ctypes.Structure):
    _fields_ = [("P", Operand), ("Q", Operand), ("M_max", _I), ("M_dev", _P), ("N", _I), ("K", _I)]


_LL = ctypes.c_longlong
_D = ctypes.c_double


class SALayer(ctypes.Structure):
    """gaddpg_sa_layer (include/gaddpg_b200.h): one conv + BatchNorm of a set-abstraction level."""
    _fields_ = [("W", _P), ("WT", _P), ("N", _I), ("K", _I), ("Kp", _I),
                ("gamma", _P), ("beta", _P), ("running_mean", _P), ("running_var", _P), ("num_batches_tracked", _P),
                ("scale", _P), ("shift", _P), ("mean", _P), ("rstd", _P), ("Y", _P),
                ("D", _P), ("bw_g", _P), ("bw_m1", _P), ("bw_m2", _P), ("dW", _P), ("dgamma", _P), ("dbeta", _P)]


class SALevel(ctypes.Structure):
    """gaddpg_sa_level: what gaddpg_sa_forward / gaddpg_sa_backward take."""
    _fields_ = [("layer", SALayer * 3), ("B", _I), ("S", _I), ("M_max", _I), ("M_dev", _P), ("count", _D),
                ("seg_off", _P), ("row_seg", _P), ("row_src", _P), ("row_w", _P), ("fixed_len", _I),
                ("G", _P), ("ldg", _I), ("rot", _I),
                ("cloud", _P), ("cloud_stride_b", _LL), ("cloud_stride_c", _I), ("skip", _I), ("Cp", _I), ("bc", _P), ("Cb", _I),
                ("ctr", _P), ("npoint", _I), ("out", _P), ("arg", _P)]


def check_sizes():
    a, b, c = _I(), _I(), _I()
    lib.gaddpg_struct_sizes(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    got = (ctypes.sizeof(Operand), ctypes.sizeof(NTProblem), ctypes.sizeof(TNProblem))
    if got != (a.value, b.value, c.value):
        raise RuntimeError("ctypes struct mirror out of sync with include/gaddpg_b200.h: %r vs %r" % (got, (a.value, b.value, c.value)))
    lib.gaddpg_sa_struct_sizes(ctypes.byref(a), ctypes.byref(b))
    got = (ctypes.sizeof(SALayer), ctypes.sizeof(SALevel))
    if got != (a.value, b.value):
        raise RuntimeError("ctypes mirror of gaddpg_sa_layer / gaddpg_sa_level out of sync: %r vs %r" % (got, (a.value, b.value)))


def dp(t):
    """device pointer (int) of a tensor or None"""
    return None if t is None else t.data_ptr()


def op_plain(x, ld=None):
    return Operand(X=dp(x), ldx=ld if ld is not None else x.shape[-1])


def op_bnrelu(y, bn, ld=None):
    """relu(y*scale+shift); bn has .scale/.shift"""
    return Operand(X=dp(y), ldx=ld if ld is not None else y.shape[-1], c0=dp(bn.scale), c1=dp(bn.shift))


def op_bnbwd_pool(e, mask, row_seg, y, bn, bb, rw=None):
    """BN backward of (D, Y) where D is the max-pool gradient, never materialised: D[r][c] = E[seg(r)][c] if bit c of
    mask[r] is set (r is the arg-max row of (seg(r), c)) else 0 (GADDPG_OP_BNBWD_POOL)"""
    return Operand(X=dp(e), ldx=e.shape[-1], Y=dp(y), ldy=y.shape[-1], rw=dp(rw), pmask=dp(mask), pseg=dp(row_seg),
                   c0=dp(bb.g), c1=dp(bb.m1), c2=dp(bb.m2), c3=dp(bn.mean), c4=dp(bn.rstd))


def op_bnbwd(d, y, bn, bb, rw=None):
    """BN backward of (D, Y): bn has .mean/.rstd, bb has .g/.m1/.m2"""
    return Operand(X=dp(d), ldx=d.shape[-1], Y=dp(y), ldy=y.shape[-1], rw=dp(rw),
                   c0=dp(bb.g), c1=dp(bb.m1), c2=dp(bb.m2), c3=dp(bn.mean), c4=dp(bn.rstd))
