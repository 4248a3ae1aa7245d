/*
 * oracle/pointnet2_cpu.c — TEST INFRASTRUCTURE ONLY (the CPU oracle).
 *
 * Plain-C restatement of the index/gather arithmetic of the third-party
 * `pointnet2_ops` extension that GA-DDPG calls but does not vendor
 * (call sites: /root/reference/core/networks.py:10,66-81 and
 * /root/reference/core/utils.py:32,795-796; install note README.md:23, un-pinned
 * fork github.com/liruiw/Pointnet2_PyTorch of erikwijmans/Pointnet2_PyTorch 3.0.0).
 * The algorithm restated here is SURVEY.md §8 Spec S1 (furthest point sampling),
 * S2 (ball query) and S3 (grouping / gather and their gradients).
 *
 * PARITY UNPINNED at this boundary: the reference holds no test, golden vector
 * or fixture for these ops and the extension cannot be built or run offline, so
 * "bit-exact" is defined against this file.  To make that definition as literal
 * as possible the FPS below SIMULATES the upstream launch shape (bs strided
 * "threads", per-thread strict-> running best, then the bs/2, bs/4, ..., 1
 * shared-memory tree with the "tie keeps the lower slot" merge), instead of
 * using a closed-form tie-break.  tests/ check the closed form used by the CUDA
 * kernel against this simulation.
 *
 * Floating point: every distance is evaluated with explicit fmaf() in a fixed
 * order (x*x, then fma(y,y,.), then fma(z,z,.)) which is what nvcc's default
 * -fmad=true contraction produces for the upstream expressions; build with
 * -ffp-contract=off so gcc adds no contraction of its own.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * legs may load this library.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* upstream cuda_utils.h: opt_n_threads = clamp(2^floor(log2 n), 1, 512) */
ORACLE_API int oracle_opt_n_threads(int work_size) {
  int p = 1;
  if (work_size < 1) return 1;
  while ((p << 1) <= work_size) p <<= 1;
  if (p > 512) p = 512;
  if (p < 1) p = 1;
  return p;
}

static inline float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = ax - bx, dy = ay - by, dz = az - bz;
  return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

/*
 * Spec S1.  xyz (B,N,3) f32 -> idx (B,m) i32.  temp (B,N) scratch, set to 1e10 here
 * (upstream allocates it filled with 1e10 in the Python/C++ wrapper).
 * block_size <= 0 means "use opt_n_threads(N)" like upstream.
 */
ORACLE_API void oracle_fps(int B, int N, int m, const float* xyz, float* temp, int32_t* idx,
                           int block_size) {
  if (m <= 0) return;
  int bs = block_size > 0 ? block_size : oracle_opt_n_threads(N);
  float* dists = (float*)malloc(sizeof(float) * (size_t)bs);
  int* dists_i = (int*)malloc(sizeof(int) * (size_t)bs);
  for (int b = 0; b < B; ++b) {
    const float* p = xyz + (size_t)b * N * 3;
    float* t = temp + (size_t)b * N;
    int32_t* out = idx + (size_t)b * m;
    for (int k = 0; k < N; ++k) t[k] = 1e10f;
    int old = 0;
    out[0] = 0;
    for (int j = 1; j < m; ++j) {
      float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
      /* phase 1: every simulated thread scans its strided slice */
      for (int tid = 0; tid < bs; ++tid) {
        int besti = 0;
        float best = -1.0f;
        for (int k = tid; k < N; k += bs) {
          float x2 = p[k * 3 + 0], y2 = p[k * 3 + 1], z2 = p[k * 3 + 2];
          float mag = fmaf(z2, z2, fmaf(y2, y2, x2 * x2));
          if ((double)mag <= 1e-3) continue; /* double literal, as upstream */
          float d = sqdist3(x2, y2, z2, x1, y1, z1);
          float d2 = d < t[k] ? d : t[k]; /* min(d, temp[k]) */
          t[k] = d2;
          if (d2 > best) {
            besti = k;
            best = d2;
          }
        }
        dists[tid] = best;
        dists_i[tid] = besti;
      }
      /* phase 2: the shared-memory tree, strides bs/2 ... 1 */
      for (int stride = bs >> 1; stride >= 1; stride >>= 1) {
        for (int tid = 0; tid < stride; ++tid) {
          float v1 = dists[tid], v2 = dists[tid + stride];
          int i1 = dists_i[tid], i2 = dists_i[tid + stride];
          dists[tid] = v1 > v2 ? v1 : v2;
          dists_i[tid] = v2 > v1 ? i2 : i1;
        }
      }
      old = dists_i[0];
      out[j] = old;
    }
  }
  free(dists);
  free(dists_i);
}

/* Spec S2.  new_xyz (B,m,3), xyz (B,N,3) -> idx (B,m,nsample) i32 (zero-filled first).
 * cnt (B,m) optional: number of in-radius hits found before the scan stopped. */
ORACLE_API void oracle_ball_query(int B, int N, int m, float radius, int nsample,
                                  const float* new_xyz, const float* xyz, int32_t* idx,
                                  int32_t* cnt_out) {
  const float radius2 = radius * radius;
  memset(idx, 0, sizeof(int32_t) * (size_t)B * m * nsample);
  for (int b = 0; b < B; ++b) {
    const float* p = xyz + (size_t)b * N * 3;
    const float* q = new_xyz + (size_t)b * m * 3;
    for (int j = 0; j < m; ++j) {
      int32_t* row = idx + ((size_t)b * m + j) * nsample;
      float cx = q[j * 3 + 0], cy = q[j * 3 + 1], cz = q[j * 3 + 2];
      int cnt = 0;
      for (int k = 0; k < N && cnt < nsample; ++k) {
        float d2 = sqdist3(cx, cy, cz, p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
        if (d2 < radius2) {
          if (cnt == 0)
            for (int l = 0; l < nsample; ++l) row[l] = k;
          row[cnt] = k;
          ++cnt;
        }
      }
      if (cnt_out) cnt_out[(size_t)b * m + j] = cnt;
    }
  }
}

/* gather_points: pts (B,C,N), idx (B,m) -> out (B,C,m) */
ORACLE_API void oracle_gather_points(int B, int C, int N, int m, const float* pts,
                                     const int32_t* idx, float* out) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < m; ++j)
        out[((size_t)b * C + c) * m + j] = pts[((size_t)b * C + c) * N + idx[(size_t)b * m + j]];
}

/* gather_points_grad: grad_out (B,C,m), idx (B,m) -> grad_pts (B,C,N) (zero-filled first) */
ORACLE_API void oracle_gather_points_grad(int B, int C, int N, int m, const float* grad_out,
                                          const int32_t* idx, float* grad_pts) {
  memset(grad_pts, 0, sizeof(float) * (size_t)B * C * N);
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c)
      for (int j = 0; j < m; ++j)
        grad_pts[((size_t)b * C + c) * N + idx[(size_t)b * m + j]] +=
            grad_out[((size_t)b * C + c) * m + j];
}

/* group_points: pts (B,C,N), idx (B,m,s) -> out (B,C,m,s) */
ORACLE_API void oracle_group_points(int B, int C, int N, int m, int s, const float* pts,
                                    const int32_t* idx, float* out) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      const float* src = pts + ((size_t)b * C + c) * N;
      float* dst = out + ((size_t)b * C + c) * m * s;
      const int32_t* ib = idx + (size_t)b * m * s;
      for (int e = 0; e < m * s; ++e) dst[e] = src[ib[e]];
    }
}

/* group_points_grad: grad_out (B,C,m,s), idx (B,m,s) -> grad_pts (B,C,N).
 * Upstream uses float atomicAdd (order not defined); the oracle sums in (j,l) order. */
ORACLE_API void oracle_group_points_grad(int B, int C, int N, int m, int s, const float* grad_out,
                                         const int32_t* idx, float* grad_pts) {
  memset(grad_pts, 0, sizeof(float) * (size_t)B * C * N);
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      float* dst = grad_pts + ((size_t)b * C + c) * N;
      const float* src = grad_out + ((size_t)b * C + c) * m * s;
      const int32_t* ib = idx + (size_t)b * m * s;
      for (int e = 0; e < m * s; ++e) dst[ib[e]] += src[e];
    }
}
