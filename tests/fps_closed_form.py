"""numpy mirror of the packed-key FPS used by the CUDA kernel (ga-ddpg_b200/csrc/index_ops.cu):
max over key = (float_bits(d2) << 32) | (0xFFFFFFFF - rank(k)),  rank = bitrev(k mod bs)*ceil(N/bs) + k div bs.
Checked against the literal launch-shape simulation in oracle/pointnet2_cpu.c (Spec S1)."""
import numpy as np


def _bitrev(x, bits):
    r = np.zeros_like(x)
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r


def fps_closed_form(xyz, m, bs):
    """xyz (N,3) float32 -> idx (m,) int32."""
    xyz = np.asarray(xyz, np.float32)
    N = xyz.shape[0]
    log2bs = int(np.log2(bs))
    per = (N + bs - 1) // bs
    k = np.arange(N, dtype=np.int64)
    rank = _bitrev(k & (bs - 1), log2bs) * per + (k >> log2bs)
    low = (0xFFFFFFFF - rank).astype(np.uint64)
    x, y, z = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    # fmaf chain in float32: emulate fma with float64 (exact product of two float32 fits in float64;
    # the sum of an exact product and a float32 rounds once to float32 up to double rounding, which the
    # tests avoid by construction — the GPU tests use the C oracle, this mirror is only for tie logic)
    def fma(a, b, c):
        return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)

    mag = fma(z, z, fma(y, y, x * x))
    valid = ~(mag.astype(np.float64) <= 1e-3)
    temp = np.full(N, 1e10, np.float32)
    idx = np.zeros(m, np.int32)
    old = 0
    for j in range(1, m):
        dx, dy, dz = x - x[old], y - y[old], z - z[old]
        d = fma(dz, dz, fma(dy, dy, dx * dx))
        t = np.minimum(d, temp)
        temp = np.where(valid, t, temp)
        key = (temp.view(np.uint32).astype(np.uint64) << np.uint64(32)) | low
        key = np.where(valid, key, np.uint64(0))
        best = key.max()
        if best == 0:
            old = 0
        else:
            r = int(0xFFFFFFFF - (int(best) & 0xFFFFFFFF))
            br, q = divmod(r, per)
            old = q * bs + int(_bitrev(np.array([br]), log2bs)[0])
        idx[j] = old
    return idx
