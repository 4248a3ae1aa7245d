"""Shared adversarial point-cloud cases for the FPS / ball-query known-answer tests
(SURVEY.md §8(c) "KATs the build must author")."""
import numpy as np


def fps_cases():
    rs = np.random.RandomState(7)
    cases = {}
    for N in (16, 31, 32, 33, 100, 500, 512, 513, 1000, 1024, 2048, 4096):
        cases["rand_N%d" % N] = (rs.uniform(-0.1, 0.4, (3, N, 3)).astype(np.float32), min(32, N))
    # exact ties: lattice clouds (many equal distances) and duplicated points
    g = np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(8), indexing="ij"), -1).reshape(-1, 3)
    lat = (g * 0.05 + 0.1).astype(np.float32)
    cases["lattice_512"] = (np.stack([lat, lat[::-1].copy(), lat[rs.permutation(512)]]), 32)
    g2 = np.stack(np.meshgrid(np.arange(16), np.arange(16), np.arange(4), indexing="ij"), -1).reshape(-1, 3)
    lat2 = (g2 * 0.03125 + 0.25).astype(np.float32)
    cases["lattice_1024"] = (np.stack([lat2, lat2[rs.permutation(1024)]]), 64)
    dup = rs.uniform(0.1, 0.3, (2, 40, 3)).astype(np.float32)
    dup = np.concatenate([dup] * 13, 1)[:, rs.permutation(520)]
    cases["duplicates_520"] = (dup, 48)
    # invalid points (|p|^2 <= 1e-3): some, all-but-one, invalid index 0, all invalid
    c = rs.uniform(0.1, 0.3, (4, 300, 3)).astype(np.float32)
    c[0, ::3] = 0.0
    c[1, :] = 0.0
    c[1, 137] = (0.2, 0.1, 0.3)
    c[2, 0] = 0.0
    c[3, :] = 0.01
    cases["invalid_300"] = (c, 16)
    # the |p|^2 == float32(1e-3) boundary (float32(1e-3) > 1e-3 as a double: such a point is VALID)
    b = rs.uniform(0.1, 0.3, (1, 64, 3)).astype(np.float32)
    s = np.float32(np.sqrt(np.float32(1e-3)))
    b[0, 5] = (s, 0, 0)
    b[0, 6] = (np.nextafter(s, np.float32(0)), 0, 0)
    b[0, 7] = (0, np.float32(0.0316), 0)
    cases["boundary_64"] = (b, 64)
    # npoint == N (SA2 case 32 -> 32) and npoint > #valid (repeats)
    sa2 = rs.uniform(0.0, 0.2, (5, 32, 3)).astype(np.float32)
    sa2[1, 3] = 0.0
    sa2[2, :20] = 0.001
    cases["sa2_32"] = (sa2, 32)
    return cases


def ball_query_cases():
    rs = np.random.RandomState(11)
    cases = {}
    xyz = rs.uniform(0.0, 0.1, (3, 700, 3)).astype(np.float32)
    new = xyz[:, rs.choice(700, 32, replace=False)].copy()
    cases["dense_r0.02_ns64"] = (xyz, new, 0.02, 64)      # >= nsample hits: truncation in k order
    cases["sparse_r0.005_ns64"] = (xyz, new, 0.005, 64)   # < nsample hits: padding with the first
    far = new.copy()
    far[:, :5] += 10.0
    cases["nohit_rows"] = (xyz, far, 0.02, 16)           # zero hits: rows stay 0
    dupc = new.copy()
    dupc[:, 1] = dupc[:, 0]
    cases["dup_centroids"] = (xyz, dupc, 0.03, 128)
    # d2 == r^2 boundary (strict <): points at exactly r on an axis
    r = np.float32(0.25)
    bx = np.zeros((1, 8, 3), np.float32)
    bx[0, 1, 0] = r
    bx[0, 2, 1] = np.nextafter(r, np.float32(0))
    bx[0, 3, 2] = np.nextafter(r, np.float32(1))
    cases["boundary"] = (bx, np.zeros((1, 2, 3), np.float32), 0.25, 4)
    s2 = rs.uniform(0.0, 0.15, (4, 32, 3)).astype(np.float32)
    cases["sa2_like"] = (s2, s2.copy(), 0.04, 128)
    return cases
