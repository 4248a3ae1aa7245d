// gemm_rows.cu — FP32 row-GEMM kernels with fused BatchNorm prologues/epilogues (sm_100a).
//
// These carry the shared MLP of the PointNet++ set-abstraction modules (1x1 Conv2d + BatchNorm2d + ReLU,
// upstream pointnet2_modules.build_shared_mlp, reached from /root/reference/core/networks.py:65-92), the
// FC/BN1d head (networks.py:84-91) and the actor/critic Linear layers (networks.py:265-300,315-351), forward
// and backward.  Arithmetic is plain FP32 FFMA: the parity bar is 1e-4 against an FP32 CPU oracle through
// nine stacked layers plus train-mode BatchNorm, which single-pass TF32 does not hold.
//
// Persistent CTAs (grid <= GADDPG_STAT_SLOTS = 2 per SM) walk the tile list of a row count that lives in
// device memory, so the same launch is valid for any duplicate-folded row count and is graph-capturable.
// BatchNorm statistics are reduced deterministically: registers -> shared memory in a fixed order ->
// one slot per CTA -> a fixed-order FP64 sum in the finalize kernel.  No float atomics anywhere.
#include "common.cuh"
#include "gemm_rows.cuh"
#include <stdlib.h>

#include "impl.h"
#include "bn_tail.cuh"

namespace {

constexpr int BK = 16;
constexpr int LDS = BK + 4;  // row stride of the staged tiles (floats): conflict-free LDS.128 for 8 consecutive rows

__device__ __forceinline__ float4 ldg4(const float* p) { return *reinterpret_cast<const float4*>(p); }

template <int MODE>
__device__ __forceinline__ float4 load_operand4(const Operand& d, int row, int col, int M, int W) {
  float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
  if (row >= M || col >= W) return r;
  float4 x = ldg4(d.X + (long long)row * d.ldx + col);
  if (MODE == OP_PLAIN) {
    r = x;
  } else if (MODE == OP_BNRELU) {
    float4 s = ldg4(d.c0 + col), t = ldg4(d.c1 + col);
    r.x = fmaxf(fmaf(x.x, s.x, t.x), 0.f);
    r.y = fmaxf(fmaf(x.y, s.y, t.y), 0.f);
    r.z = fmaxf(fmaf(x.z, s.z, t.z), 0.f);
    r.w = fmaxf(fmaf(x.w, s.w, t.w), 0.f);
  } else {
    float4 y = ldg4(d.Y + (long long)row * d.ldy + col);
    float4 g = ldg4(d.c0 + col), m1 = ldg4(d.c1 + col), m2 = ldg4(d.c2 + col), mu = ldg4(d.c3 + col),
           rs = ldg4(d.c4 + col);
    float w = d.rw ? d.rw[row] : 1.f;
    r.x = g.x * (x.x - w * (m1.x + (y.x - mu.x) * rs.x * m2.x));
    r.y = g.y * (x.y - w * (m1.y + (y.y - mu.y) * rs.y * m2.y));
    r.z = g.z * (x.z - w * (m1.z + (y.z - mu.z) * rs.z * m2.z));
    r.w = g.w * (x.w - w * (m1.w + (y.w - mu.w) * rs.w * m2.w));
  }
  return r;
}

// ------------------------------------------------------------------------------------------------
// NT: C[M,N] = epi(pro(A)[M,K] . B[N,K]^T).  256 threads, thread (ty,tx) owns rows ty+TY*i, cols tx+TX*j.
// ------------------------------------------------------------------------------------------------
#ifndef GADDPG_NT_MINB
#define GADDPG_NT_MINB 1
#endif
template <int BM, int BN, int TM, int TN, int AMODE, int EMODE>
__global__ void __launch_bounds__(256, GADDPG_NT_MINB) gemm_nt_kernel(const NTGroup grp) {
  constexpr int TX = BN / TN, TY = BM / TM;
  static_assert(TX * TY == 256, "tile shape must map onto 256 threads");
  constexpr int TILE_FLOATS = (BM + BN) * LDS;
  constexpr int RED_FLOATS = 2 * TY * BN;
  constexpr int SM_FLOATS = TILE_FLOATS > RED_FLOATS ? TILE_FLOATS : RED_FLOATS;
  __shared__ __align__(16) float smem[SM_FLOATS];
  __shared__ float s_acc[2 * 1024];  // per-CTA running (sum, sum2) per output column (N <= 1024 when stats on)
  float* As = smem;
  float* Bs = smem + BM * LDS;

  const NTProblem& p = grp.p[blockIdx.y];
  const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
  int M = p.M_dev ? *p.M_dev : p.M_max;
  M = M < p.M_max ? M : p.M_max;
  const int N = p.N, K = p.K;
  const bool do_stats = (p.stats != nullptr);
  if (do_stats)
    for (int c = tid; c < 2 * N; c += 256) s_acc[c] = 0.f;
  __syncthreads();

  const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN, nk = (K + BK - 1) / BK;
  constexpr int A_LD = (BM * (BK / 4) + 255) / 256, B_LD = (BN * (BK / 4) + 255) / 256;

  for (int t = blockIdx.x; t < tiles_m * tiles_n; t += gridDim.x) {
    const int row0 = (t / tiles_n) * BM, col0 = (t % tiles_n) * BN;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra[A_LD], rb[B_LD];
    auto fetch = [&](int kc) {
#pragma unroll
      for (int i = 0; i < A_LD; ++i) {
        int f = tid + 256 * i;
        int r = f / (BK / 4), kq = (f % (BK / 4)) * 4;
        ra[i] = (f < BM * (BK / 4)) ? load_operand4<AMODE>(p.A, row0 + r, kc * BK + kq, M, K)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < B_LD; ++i) {
        int f = tid + 256 * i;
        int r = f / (BK / 4), kq = (f % (BK / 4)) * 4;
        int n = col0 + r, k = kc * BK + kq;
        rb[i] = (f < BN * (BK / 4) && n < N && k < K) ? ldg4(p.Bw + (long long)n * p.ldb + k)
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    fetch(0);
    for (int kc = 0; kc < nk; ++kc) {
#pragma unroll
      for (int i = 0; i < A_LD; ++i) {
        int f = tid + 256 * i;
        if (f < BM * (BK / 4)) *reinterpret_cast<float4*>(As + (f / (BK / 4)) * LDS + (f % (BK / 4)) * 4) = ra[i];
      }
#pragma unroll
      for (int i = 0; i < B_LD; ++i) {
        int f = tid + 256 * i;
        if (f < BN * (BK / 4)) *reinterpret_cast<float4*>(Bs + (f / (BK / 4)) * LDS + (f % (BK / 4)) * 4) = rb[i];
      }
      __syncthreads();
      if (kc + 1 < nk) fetch(kc + 1);
#pragma unroll
      for (int kk = 0; kk < BK; kk += 4) {
        float4 a[TM];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(As + (ty + TY * i) * LDS + kk);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          float4 b = *reinterpret_cast<const float4*>(Bs + (tx + TX * j) * LDS + kk);
#pragma unroll
          for (int i = 0; i < TM; ++i) {
            acc[i][j] = fmaf(a[i].x, b.x, acc[i][j]);
            acc[i][j] = fmaf(a[i].y, b.y, acc[i][j]);
            acc[i][j] = fmaf(a[i].z, b.z, acc[i][j]);
            acc[i][j] = fmaf(a[i].w, b.w, acc[i][j]);
          }
        }
      }
      __syncthreads();
    }

    // ---- epilogue ----
    float s0[TN], s1[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) s0[j] = s1[j] = 0.f;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int row = row0 + ty + TY * i;
      if (row < M) {
        float w = 1.f;
        if (EMODE == EPI_STORE && do_stats && p.srw) w = p.srw[row];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const int col = col0 + tx + TX * j;
          if (col < N) {
            float v = acc[i][j];
            if (EMODE == EPI_STORE) {
              if (p.bias) v += p.bias[col];
              if (p.relu) v = fmaxf(v, 0.f);
              p.C[(long long)row * p.ldc + col] = v;
              if (do_stats) {
                s0[j] = fmaf(w, v, s0[j]);
                s1[j] = fmaf(w * v, v, s1[j]);
              }
            } else {
              float yp = p.Yprev[(long long)row * p.ldyp + col];
              float z = p.psc ? fmaf(yp, p.psc[col], p.psh[col]) : yp;
              v = z > 0.f ? v : 0.f;
              p.C[(long long)row * p.ldc + col] = v;
              if (do_stats) {
                s0[j] += v;
                s1[j] = fmaf(v, (yp - p.pmean[col]) * p.prstd[col], s1[j]);
              }
            }
          }
        }
      }
    }
    if (do_stats) {
      float* red0 = smem;
      float* red1 = smem + TY * BN;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        red0[ty * BN + tx + TX * j] = s0[j];
        red1[ty * BN + tx + TX * j] = s1[j];
      }
      __syncthreads();
      if (tid < BN && col0 + tid < N) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int y = 0; y < TY; ++y) {
          a0 += red0[y * BN + tid];
          a1 += red1[y * BN + tid];
        }
        s_acc[col0 + tid] += a0;
        s_acc[N + col0 + tid] += a1;
      }
      __syncthreads();
    }
  }
  if (do_stats) {
    __syncthreads();
    for (int slot = blockIdx.x; slot < GADDPG_STAT_SLOTS; slot += gridDim.x) {
      const bool mine = (slot == (int)blockIdx.x);
      for (int c = tid; c < 2 * N; c += 256) p.stats[(long long)slot * 2 * N + c] = mine ? s_acc[c] : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// TN: partial[s][N][K] = sum over the rows of split s of pro1(P)[r,:]^T pro2(Q)[r,:]; optional column
// sums of pro1(P) (bias gradients).  Output tile 128x128 (or 64x64), 256 threads, 8x8 (4x4) per thread.
// ------------------------------------------------------------------------------------------------
template <int BTN, int BTK, int PMODE, int QMODE>
__global__ void __launch_bounds__(256) gemm_tn_kernel(const TNProblem p, float* __restrict__ partial,
                                                      float* __restrict__ partial_bias, int splits) {
  constexpr int BR = 16;                        // rows per staged chunk
  constexpr int TN_ = BTN / 16, TK_ = BTK / 16;  // outputs per thread in each dimension (two runs of H each)
  constexpr int HN = TN_ / 2, HK = TK_ / 2;
  static_assert(BTN % 32 == 0 && BTK % 32 == 0, "16x16 threads, two runs per dimension");
  __shared__ __align__(16) float Ps[BR * BTN];
  __shared__ __align__(16) float Qs[BR * BTK];
  const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
  int M = p.M_dev ? *p.M_dev : p.M_max;
  M = M < p.M_max ? M : p.M_max;
  const int N = p.N, K = p.K;
  const int tiles_k = (K + BTK - 1) / BTK;
  const int tile = blockIdx.x, split = blockIdx.y;
  const int n0 = (tile / tiles_k) * BTN, k0 = (tile % tiles_k) * BTK;
  const bool want_bias = (partial_bias != nullptr) && (tile % tiles_k == 0);

  float acc[TN_][TK_];
  float bsum[TN_];
#pragma unroll
  for (int i = 0; i < TN_; ++i) {
    bsum[i] = 0.f;
#pragma unroll
    for (int j = 0; j < TK_; ++j) acc[i][j] = 0.f;
  }
  const int nchunks = (M + BR - 1) / BR;
  constexpr int LDP = (BR * BTN / 4 + 255) / 256, LDQ = (BR * BTK / 4 + 255) / 256;
  float4 rp[LDP], rq[LDQ];
  auto fetch = [&](int ch) {
#pragma unroll
    for (int i = 0; i < LDP; ++i) {
      int f = tid + 256 * i;
      int r = f / (BTN / 4), c = (f % (BTN / 4)) * 4;
      rp[i] = (f < BR * BTN / 4) ? load_operand4<PMODE>(p.P, ch * BR + r, n0 + c, M, N) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < LDQ; ++i) {
      int f = tid + 256 * i;
      int r = f / (BTK / 4), c = (f % (BTK / 4)) * 4;
      rq[i] = (f < BR * BTK / 4) ? load_operand4<QMODE>(p.Q, ch * BR + r, k0 + c, M, K) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  int ch = split;
  if (ch < nchunks) fetch(ch);
  for (; ch < nchunks; ch += splits) {
#pragma unroll
    for (int i = 0; i < LDP; ++i) {
      int f = tid + 256 * i;
      if (f < BR * BTN / 4) *reinterpret_cast<float4*>(Ps + f * 4) = rp[i];
    }
#pragma unroll
    for (int i = 0; i < LDQ; ++i) {
      int f = tid + 256 * i;
      if (f < BR * BTK / 4) *reinterpret_cast<float4*>(Qs + f * 4) = rq[i];
    }
    __syncthreads();
    if (ch + splits < nchunks) fetch(ch + splits);
#pragma unroll
    for (int r = 0; r < BR; ++r) {
      float a[TN_], b[TK_];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int e = 0; e < HN; ++e) a[h * HN + e] = Ps[r * BTN + h * (BTN / 2) + ty * HN + e];
#pragma unroll
        for (int e = 0; e < HK; ++e) b[h * HK + e] = Qs[r * BTK + h * (BTK / 2) + tx * HK + e];
      }
#pragma unroll
      for (int i = 0; i < TN_; ++i) {
#pragma unroll
        for (int j = 0; j < TK_; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        bsum[i] += a[i];
      }
    }
    __syncthreads();
  }
  float* out = partial + (long long)split * N * K;
#pragma unroll
  for (int i = 0; i < TN_; ++i) {
    int n = n0 + (i / HN) * (BTN / 2) + ty * HN + (i % HN);
    if (n < N) {
#pragma unroll
      for (int j = 0; j < TK_; ++j) {
        int k = k0 + (j / HK) * (BTK / 2) + tx * HK + (j % HK);
        if (k < K) out[(long long)n * K + k] = acc[i][j];
      }
      if (want_bias && tx == 0) partial_bias[(long long)split * N + n] = bsum[i];
    }
  }
}

// dst[n][(k+rot) % Ktrue] (+)= sum_s partial[s][n][k]   for k < Ktrue.  Block = 32 outputs x 8 split lanes: lane g
// sums splits g, g+8, ... (coalesced over the 32 outputs), then the 8 lane sums are added in a fixed order.
__global__ void __launch_bounds__(256) tn_reduce_kernel(const float* __restrict__ partial, int splits, int N, int Ntrue, int K,
                                                        int Ktrue, int rot, float* __restrict__ dst, int ldd, int accumulate,
                                                        float scale) {
  __shared__ float red[8][33];
  const int ox = threadIdx.x & 31, g = threadIdx.x >> 5;
  const long long total = (long long)Ntrue * Ktrue;
  for (long long e0 = (long long)blockIdx.x * 32; e0 < total; e0 += (long long)gridDim.x * 32) {
    const long long e = e0 + ox;
    float s = 0.f;
    int n = 0, k = 0;
    if (e < total) {
      n = (int)(e / Ktrue);
      k = (int)(e % Ktrue);
      const float* src = partial + (long long)n * K + k;
      const long long stride = (long long)N * K;
      s = ordered_sum<4>((splits - g + 7) / 8, [&](int i) { return src[(long long)(g + 8 * i) * stride]; });
    }
    red[g][ox] = s;
    __syncthreads();
    if (g == 0 && e < total) {
      float t = ((red[0][ox] + red[1][ox]) + (red[2][ox] + red[3][ox])) + ((red[4][ox] + red[5][ox]) + (red[6][ox] + red[7][ox]));
      int kd = k + rot;
      if (kd >= Ktrue) kd -= Ktrue;
      float* d = dst + (long long)n * ldd + kd;
      *d = (accumulate ? *d : 0.f) + t * scale;
    }
    __syncthreads();
  }
}
// same result layout for few splits (the M = B layers: 4-8 row splits, up to 512 x 1024 outputs): one thread per output,
// splits summed in index order — no shared-memory round trip, fully coalesced.
__global__ void __launch_bounds__(256) tn_reduce_few_kernel(const float* __restrict__ partial, int splits, int N, int Ntrue, int K,
                                                            int Ktrue, int rot, float* __restrict__ dst, int ldd, int accumulate,
                                                            float scale) {
  const long long total = (long long)Ntrue * Ktrue, stride = (long long)N * K;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(e / Ktrue), k = (int)(e % Ktrue);
    const float* src = partial + (long long)n * K + k;
    float s = 0.f;
#pragma unroll 8
    for (int sp = 0; sp < splits; ++sp) s += src[sp * stride];
    int kd = k + rot;
    if (kd >= Ktrue) kd -= Ktrue;
    float* d = dst + (long long)n * ldd + kd;
    *d = (accumulate ? *d : 0.f) + s * scale;
  }
}
__global__ void bias_reduce_kernel(const float* __restrict__ partial, int splits, int N, int Ntrue, float* __restrict__ dst,
                                   int accumulate) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= Ntrue) return;
  const float s = ordered_sum<8>(splits, [&](int sp) { return partial[(long long)sp * N + n]; });
  dst[n] = (accumulate ? dst[n] : 0.f) + s;
}

// ---- BatchNorm finalize kernels ------------------------------------------------------------------
// forward: slots of (sum w*y, sum w*y^2) -> batch mean / biased var -> scale, shift (+ mean, rstd for
// backward) and the running-stat update PyTorch performs (momentum 0.1, UNBIASED var, counter += 1;
// torch.nn.BatchNorm2d defaults as instantiated by upstream build_shared_mlp / networks.py:86,89).
// 1024 threads = 32 consecutive channels (lane) x 32 slot groups (warp): every load instruction reads 32 consecutive
// floats of one slot (coalesced), each thread keeps FP64 partials over its <= 10 slots (independent loads, all in flight
// together), and warp 0 folds the 32 groups in a fixed order — deterministic.
constexpr int BNF_GROUPS = 32;
__device__ __forceinline__ bool slot_sums(const float* __restrict__ stats, int C, int c, double& s, double& q) {
  __shared__ double sh[BNF_GROUPS][2][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double s0 = 0.0, q0 = 0.0;
  if (c < C) {
    constexpr int NIT = (GADDPG_STAT_SLOTS + BNF_GROUPS - 1) / BNF_GROUPS;
    float a[NIT], b[NIT];
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      const int slot = warp + i * BNF_GROUPS;
      a[i] = slot < GADDPG_STAT_SLOTS ? stats[(long long)slot * 2 * C + c] : 0.f;
      b[i] = slot < GADDPG_STAT_SLOTS ? stats[(long long)slot * 2 * C + C + c] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
      s0 += (double)a[i];
      q0 += (double)b[i];
    }
  }
  sh[warp][0][lane] = s0;
  sh[warp][1][lane] = q0;
  __syncthreads();
  if (warp != 0 || c >= C) return false;
  s = 0.0;
  q = 0.0;
#pragma unroll 8
  for (int w = 0; w < BNF_GROUPS; ++w) {
    s += sh[w][0][lane];
    q += sh[w][1][lane];
  }
  return true;
}

__global__ void __launch_bounds__(1024) bn_finalize_fwd_kernel(const float* __restrict__ stats, int C, double count,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              float eps, float momentum, float* __restrict__ running_mean,
                                                              float* __restrict__ running_var,
                                                              long long* __restrict__ num_batches_tracked, int training,
                                                              float* __restrict__ scale, float* __restrict__ shift,
                                                              float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  if (blockIdx.x == 0 && threadIdx.x == 0 && training && num_batches_tracked) *num_batches_tracked += 1;
  float mean, var;
  if (training) {
    double s, q;
    if (!slot_sums(stats, C, c, s, q)) return;
    double m = s / count;
    double v = q / count - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    var = (float)v;
    if (running_mean) {
      double unbiased = count > 1.0 ? v * (count / (count - 1.0)) : v;
      // momentum == 1: staging mode (deferred update, gaddpg_bn_running_update) — store the batch values verbatim
      running_mean[c] = (momentum == 1.f) ? mean : (1.f - momentum) * running_mean[c] + momentum * mean;
      running_var[c] = (momentum == 1.f) ? (float)unbiased : (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
  } else {
    if (threadIdx.x >= 32 || c >= C) return;
    mean = running_mean[c];
    var = running_var[c];
  }
  float rstd = 1.0f / sqrtf(var + eps);
  float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  if (mean_out) mean_out[c] = mean;
  if (rstd_out) rstd_out[c] = rstd;
}

// backward: slots of (sum D, sum D*xhat) -> m1, m2, g = gamma*rstd, and dgamma / dbeta
__global__ void __launch_bounds__(1024) bn_finalize_bwd_kernel(const float* __restrict__ stats, int C, double count,
                                                              const float* __restrict__ gamma, const float* __restrict__ rstd,
                                                              float* __restrict__ g, float* __restrict__ m1,
                                                              float* __restrict__ m2, float* __restrict__ dgamma,
                                                              float* __restrict__ dbeta, int accumulate) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  double s, q;
  if (!slot_sums(stats, C, c, s, q)) return;
  m1[c] = (float)(s / count);
  m2[c] = (float)(q / count);
  g[c] = gamma[c] * rstd[c];
  if (dgamma) dgamma[c] = (accumulate ? dgamma[c] : 0.f) + (float)q;
  if (dbeta) dbeta[c] = (accumulate ? dbeta[c] : 0.f) + (float)s;
}

// deferred running-statistics update over a whole running-stat arena (see gaddpg_bn_running_update)
__global__ void __launch_bounds__(256) bn_running_update_kernel(float* __restrict__ running, const float* __restrict__ staged,
                                                                long long n, float momentum, long long* __restrict__ nbt,
                                                                int n_layers) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) running[i] = (1.f - momentum) * running[i] + momentum * staged[i];
  if (blockIdx.x == 0 && nbt && (int)threadIdx.x < n_layers) nbt[threadIdx.x] += 1;
}

template <int AMODE, int EMODE>
int launch_nt(const NTGroup& g, int nprob, cudaStream_t st) {
  int maxM = 0, maxN = 0;
  for (int i = 0; i < nprob; ++i) {
    maxM = g.p[i].M_max > maxM ? g.p[i].M_max : maxM;
    maxN = g.p[i].N > maxN ? g.p[i].N : maxN;
  }
  if (maxM == 0) return GADDPG_OK;
  if (maxM <= 1024 || maxN < 64) {  // small problems: 32x64 tiles so a few hundred rows still spread over the chip
    int tiles = ceil_div(maxM, 32) * ceil_div(maxN, 64);
    dim3 grid(tiles < GADDPG_STAT_SLOTS ? tiles : GADDPG_STAT_SLOTS, nprob);
    gemm_nt_kernel<32, 64, 2, 4, AMODE, EMODE><<<grid, 256, 0, st>>>(g);
  } else if (maxN <= 64) {
    int tiles = ceil_div(maxM, 128);
    dim3 grid(tiles < GADDPG_STAT_SLOTS ? tiles : GADDPG_STAT_SLOTS, nprob);
    gemm_nt_kernel<128, 64, 8, 4, AMODE, EMODE><<<grid, 256, 0, st>>>(g);
  } else {
    int tiles = ceil_div(maxM, 128) * ceil_div(maxN, 128);
    dim3 grid(tiles < GADDPG_STAT_SLOTS ? tiles : GADDPG_STAT_SLOTS, nprob);
    gemm_nt_kernel<128, 128, 8, 8, AMODE, EMODE><<<grid, 256, 0, st>>>(g);
  }
  GADDPG_CHECK_LAUNCH("gemm_nt_kernel");
  return GADDPG_OK;
}

int check_operand(const Operand& o, int mode, int W, const char* what) {
  GADDPG_CHECK_ARG(o.X && (o.ldx % 4) == 0 && o.ldx >= W, "%s: bad source (ld=%d, width=%d)", what, o.ldx, W);
  GADDPG_CHECK_ARG(((uintptr_t)o.X % 16) == 0, "%s: source not 16-byte aligned", what);
  if (mode == OP_BNRELU) GADDPG_CHECK_ARG(o.c0 && o.c1, "%s: BNRELU needs scale/shift", what);
  if (mode == OP_BNBWD || mode == OP_BNBWD_POOL)
    GADDPG_CHECK_ARG(o.Y && (o.ldy % 4) == 0 && o.c0 && o.c1 && o.c2 && o.c3 && o.c4, "%s: BNBWD operand incomplete", what);
  if (mode == OP_BNBWD_POOL) GADDPG_CHECK_ARG(o.pmask && o.pseg, "%s: BNBWD_POOL needs the arg-max bit mask and the row -> segment map", what);
  return GADDPG_OK;
}

}  // namespace

static int g_tc_enabled = -1;
// level: 0 = FP32 FFMA only; 1 = + resident-weight tcgen05 NT (tc_gemm.cu); 2 = + K-chunked tcgen05 NT; 3 = + tcgen05 TN
// (default); 4 = like 3 but the K-chunked kernel also takes the shapes of the resident-weight kernel (A/B testing)
void gaddpg_set_tensor_core_impl(int level) { g_tc_enabled = level < 0 ? 0 : (level > 4 ? 4 : level); }
int gaddpg_skinny_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("GADDPG_SKINNY");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on && gaddpg_get_tensor_core_impl() >= 1;
}
int gaddpg_get_tensor_core_impl() {
  if (g_tc_enabled < 0) {
    const char* e = getenv("GADDPG_TC");
    g_tc_enabled = (e && e[0] >= '0' && e[0] <= '4') ? (e[0] - '0') : 3;
  }
  return g_tc_enabled;
}

int gaddpg_gemm_nt_impl(const NTGroup* g, int nprob, int amode, int emode, void* stream) {
  GADDPG_CHECK_ARG(g && nprob >= 1 && nprob <= GADDPG_MAX_GROUP, "gemm_nt: bad group size %d", nprob);
  for (int i = 0; i < nprob; ++i) {
    const NTProblem& p = g->p[i];
    GADDPG_CHECK_ARG(p.K >= 4 && (p.K % 4) == 0 && p.N >= 1 && p.M_max >= 0, "gemm_nt[%d]: bad shape M=%d N=%d K=%d", i,
                     p.M_max, p.N, p.K);
    int rc = check_operand(p.A, amode, p.K, "gemm_nt A");
    if (rc) return rc;
    GADDPG_CHECK_ARG(p.Bw && (p.ldb % 4) == 0 && p.ldb >= p.K && ((uintptr_t)p.Bw % 16) == 0, "gemm_nt[%d]: bad B", i);
    GADDPG_CHECK_ARG(p.C && p.ldc >= p.N, "gemm_nt[%d]: bad C", i);
    GADDPG_CHECK_ARG(!p.stats || p.N <= 1024, "gemm_nt[%d]: stats need N <= 1024", i);
    if (emode == EPI_DMASK) {
      GADDPG_CHECK_ARG(p.Yprev && p.ldyp >= p.N, "gemm_nt[%d]: DMASK needs Yprev", i);
      GADDPG_CHECK_ARG(!p.stats || (p.pmean && p.prstd), "gemm_nt[%d]: DMASK stats need mean/rstd", i);
      GADDPG_CHECK_ARG(!p.psc || p.psh, "gemm_nt[%d]: psc without psh", i);
    }
  }
  // BatchNorm finalize tails (gaddpg_bn_tail): fused into the tcgen05 / mma.sync kernels of single-problem launches; every other
  // path runs the separate finalize kernel behind the product, so the call means the same everywhere
  bool any_tail = false;
  for (int i = 0; i < nprob; ++i) {
    const NTProblem& p = g->p[i];
    if (p.tail.kind == 0) continue;
    GADDPG_CHECK_ARG(p.stats && (p.tail.kind == 1 || p.tail.kind == 2) && p.tail.a && p.tail.b && p.tail.o0 && p.tail.o1 &&
                         (p.tail.kind == 1 || p.tail.o2) && p.tail.count >= 1.0,
                     "gemm_nt[%d]: incomplete BatchNorm tail", i);
    any_tail = true;
  }
  if (any_tail) {
    const bool fused_path = nprob == 1 && bnt_fusable(g->p[0].tail, g->p[0].N) &&
                            ((gaddpg_get_tensor_core_impl() >= 1 && gaddpg_get_tensor_core_impl() <= 3 && gaddpg_tc_gemm_supported(g->p[0], amode, emode)) ||
                             (gaddpg_get_tensor_core_impl() >= 2 && gaddpg_tc_nt_kc_supported(g->p[0], amode, emode)) ||
                             (gaddpg_skinny_enabled() && gaddpg_skinny_supported(*g, nprob, amode, emode) && !g->p[0].pool_keys &&
                              !g->p[0].no_store && amode != OP_BNBWD_POOL));
    if (!fused_path) {
      NTGroup plain = *g;
      for (int i = 0; i < nprob; ++i) plain.p[i].tail.kind = 0;
      int rc = gaddpg_gemm_nt_impl(&plain, nprob, amode, emode, stream);
      for (int i = 0; i < nprob && rc == GADDPG_OK; ++i)
        if (g->p[i].tail.kind) rc = gaddpg_bn_tail_separate(g->p[i].tail, g->p[i].stats, g->p[i].N, stream);
      return rc;
    }
  }
  if (nprob == 1 && gaddpg_get_tensor_core_impl() >= 1 && gaddpg_get_tensor_core_impl() <= 3 &&
      gaddpg_tc_gemm_supported(g->p[0], amode, emode))
    return gaddpg_tc_gemm_nt_impl(&g->p[0], amode, emode, stream);  // tcgen05 3xTF32 path for the wide SA layers
  if (nprob == 1 && gaddpg_get_tensor_core_impl() >= 2 && gaddpg_tc_nt_kc_supported(g->p[0], amode, emode))
    return gaddpg_tc_nt_kc_impl(&g->p[0], amode, emode, stream);   // K-chunked tcgen05 path (SA2 / SA3 / FC / heads)
  for (int i = 0; i < nprob; ++i)
    if (g->p[i].pool_keys || g->p[i].no_store) {
      gaddpg_set_error("gemm_nt: the fused max-pool epilogue (pool_keys / no_store) is only available on the tcgen05 kernels");
      return GADDPG_ERR_UNSUPPORTED;
    }
  if (amode == OP_BNBWD_POOL) {
    gaddpg_set_error("gemm_nt: GADDPG_OP_BNBWD_POOL is only taken by the tcgen05 whole-K kernel (K = 128, N <= 128, DMASK epilogue)");
    return GADDPG_ERR_UNSUPPORTED;
  }
  if (gaddpg_skinny_enabled() && gaddpg_skinny_supported(*g, nprob, amode, emode))
    return gaddpg_skinny_nt_impl(g, nprob, amode, emode, stream);  // few-hundred-row problems: cp.async ring + mma.sync 3xTF32
  cudaStream_t st = (cudaStream_t)stream;
#define NT_CASE(A, E) \
  if (amode == A && emode == E) return launch_nt<A, E>(*g, nprob, st)
  NT_CASE(OP_PLAIN, EPI_STORE);
  NT_CASE(OP_BNRELU, EPI_STORE);
  NT_CASE(OP_PLAIN, EPI_DMASK);
  NT_CASE(OP_BNBWD, EPI_DMASK);
  NT_CASE(OP_BNBWD, EPI_STORE);
#undef NT_CASE
  gaddpg_set_error("gemm_nt: unsupported mode pair (%d,%d)", amode, emode);
  return GADDPG_ERR_UNSUPPORTED;
}

size_t gaddpg_gemm_tn_workspace_bytes_impl() {
  // splits*N*K floats with splits*tiles <= 2*GADDPG_STAT_SLOTS and N*K <= tiles*128*128, plus bias partials
  return (size_t)2 * GADDPG_STAT_SLOTS * 128 * 128 * sizeof(float) + (size_t)64 * 1024 * sizeof(float);
}

int gaddpg_gemm_tn_impl(const TNProblem* p, int pmode, int qmode, float* dW, int ldd, int Ntrue, int Ktrue, int rot,
                        float* dbias, int accumulate, float* ws, size_t ws_bytes, void* stream) {
  GADDPG_CHECK_ARG(p && dW && ws, "gemm_tn: null pointer");
  GADDPG_CHECK_ARG(p->N >= 4 && (p->N % 4) == 0 && p->K >= 4 && (p->K % 4) == 0, "gemm_tn: bad shape N=%d K=%d", p->N, p->K);
  GADDPG_CHECK_ARG(Ktrue >= 1 && Ktrue <= p->K && rot >= 0 && rot < Ktrue && ldd >= Ktrue && Ntrue >= 1 && Ntrue <= p->N,
                   "gemm_tn: bad output mapping");
  int rc = check_operand(p->P, pmode, p->N, "gemm_tn P");
  if (rc) return rc;
  rc = check_operand(p->Q, qmode, p->K, "gemm_tn Q");
  if (rc) return rc;
  if (p->M_max == 0) return GADDPG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (gaddpg_get_tensor_core_impl() >= 3 && gaddpg_tc_tn_supported(*p, pmode, qmode)) {
    GADDPG_CHECK_ARG(ws_bytes >= gaddpg_gemm_tn_workspace_bytes_impl(), "gemm_tn: workspace too small (%zu bytes given)", ws_bytes);
    float* wsb = ws + (size_t)2 * GADDPG_STAT_SLOTS * 128 * 128;
    int sp = 1;
    rc = gaddpg_tc_tn_impl(p, pmode, qmode, ws, (size_t)2 * GADDPG_STAT_SLOTS * 128 * 128, dbias ? wsb : nullptr, &sp, stream);
    if (rc) return rc;
    long long tot = (long long)Ntrue * Ktrue;
    int rg = (int)((tot + 31) / 32 < 1184 ? (tot + 31) / 32 : 1184);
    if (sp <= 16) {
      int fg = (int)((tot + 255) / 256 < 1184 ? (tot + 255) / 256 : 1184);
      tn_reduce_few_kernel<<<fg, 256, 0, st>>>(ws, sp, p->N, Ntrue, p->K, Ktrue, rot, dW, ldd, accumulate, 1.0f);
    } else {
      tn_reduce_kernel<<<rg, 256, 0, st>>>(ws, sp, p->N, Ntrue, p->K, Ktrue, rot, dW, ldd, accumulate, 1.0f);
    }
    GADDPG_CHECK_LAUNCH("tn_reduce_kernel");
    if (dbias) {
      bias_reduce_kernel<<<ceil_div(Ntrue, 128), 128, 0, st>>>(wsb, sp, p->N, Ntrue, dbias, accumulate);
      GADDPG_CHECK_LAUNCH("bias_reduce_kernel");
    }
    return GADDPG_OK;
  }
  if (pmode == OP_BNBWD_POOL) {
    gaddpg_set_error("gemm_tn: GADDPG_OP_BNBWD_POOL is only taken by the tcgen05 kernel (Q mode BNRELU)");
    return GADDPG_ERR_UNSUPPORTED;
  }
  const int BTN = p->N <= 64 ? 64 : 128, BTK = p->K <= 64 ? 64 : 128;
  int tiles = ceil_div(p->N, BTN) * ceil_div(p->K, BTK);
  int chunks = ceil_div(p->M_max, 16);
  int splits = GADDPG_STAT_SLOTS / tiles;
  if (splits > chunks) splits = chunks;
  if (splits > 64 * 1024 / p->N) splits = 64 * 1024 / p->N;
  if (splits < 1) splits = 1;
  size_t need = (size_t)splits * p->N * p->K * sizeof(float);
  float* ws_bias = ws + (size_t)2 * GADDPG_STAT_SLOTS * 128 * 128;
  GADDPG_CHECK_ARG(need <= (size_t)2 * GADDPG_STAT_SLOTS * 128 * 128 * sizeof(float) &&
                       ws_bytes >= gaddpg_gemm_tn_workspace_bytes_impl(),
                   "gemm_tn: workspace too small (%zu bytes given)", ws_bytes);
  dim3 grid(tiles, splits);
#define TN_CASE(PM, QM)                                                                                         \
  if (pmode == PM && qmode == QM) {                                                                             \
    float* wb = dbias ? ws_bias : nullptr;                                                                      \
    if (BTN == 64 && BTK == 64)                                                                                 \
      gemm_tn_kernel<64, 64, PM, QM><<<grid, 256, 0, st>>>(*p, ws, wb, splits);                                 \
    else if (BTN == 64)                                                                                         \
      gemm_tn_kernel<64, 128, PM, QM><<<grid, 256, 0, st>>>(*p, ws, wb, splits);                                \
    else if (BTK == 64)                                                                                         \
      gemm_tn_kernel<128, 64, PM, QM><<<grid, 256, 0, st>>>(*p, ws, wb, splits);                                \
    else                                                                                                        \
      gemm_tn_kernel<128, 128, PM, QM><<<grid, 256, 0, st>>>(*p, ws, wb, splits);                               \
  } else
  TN_CASE(OP_PLAIN, OP_PLAIN)
  TN_CASE(OP_PLAIN, OP_BNRELU)
  TN_CASE(OP_BNBWD, OP_PLAIN)
  TN_CASE(OP_BNBWD, OP_BNRELU) {
    gaddpg_set_error("gemm_tn: unsupported mode pair (%d,%d)", pmode, qmode);
    return GADDPG_ERR_UNSUPPORTED;
  }
#undef TN_CASE
  GADDPG_CHECK_LAUNCH("gemm_tn_kernel");
  long long total = (long long)Ntrue * Ktrue;
  int rgrid = (int)((total + 31) / 32 < 1184 ? (total + 31) / 32 : 1184);
  tn_reduce_kernel<<<rgrid, 256, 0, st>>>(ws, splits, p->N, Ntrue, p->K, Ktrue, rot, dW, ldd, accumulate, 1.0f);
  GADDPG_CHECK_LAUNCH("tn_reduce_kernel");
  if (dbias) {
    bias_reduce_kernel<<<ceil_div(Ntrue, 128), 128, 0, st>>>(ws_bias, splits, p->N, Ntrue, dbias, accumulate);
    GADDPG_CHECK_LAUNCH("bias_reduce_kernel");
  }
  return GADDPG_OK;
}

int gaddpg_bn_finalize_fwd_impl(const float* stats, int C, double count, const float* gamma, const float* beta, float eps,
                                float momentum, float* running_mean, float* running_var, long long* nbt, int training,
                                float* scale, float* shift, float* mean_out, float* rstd_out, void* stream) {
  GADDPG_CHECK_ARG(C >= 1 && gamma && beta && scale && shift, "bn_finalize_fwd: null pointer");
  GADDPG_CHECK_ARG(training ? (stats != nullptr && count >= 1.0) : (running_mean && running_var),
                   "bn_finalize_fwd: missing statistics source");
  bn_finalize_fwd_kernel<<<ceil_div(C, 32), 1024, 0, (cudaStream_t)stream>>>(stats, C, count, gamma, beta, eps, momentum,
                                                                            running_mean, running_var, nbt, training,
                                                                            scale, shift, mean_out, rstd_out);
  GADDPG_CHECK_LAUNCH("bn_finalize_fwd_kernel");
  return GADDPG_OK;
}

// a BatchNorm tail as its own launch (producers without a fused tail)
int gaddpg_bn_tail_separate(const BNTail& t, const float* stats, int C, void* stream) {
  if (t.kind == 1)
    return gaddpg_bn_finalize_fwd_impl(stats, C, t.count, t.a, t.b, t.eps, t.momentum, t.running_mean, t.running_var,
                                       t.num_batches_tracked, 1, t.o0, t.o1, t.o2, t.o3, stream);
  if (t.kind == 2)
    return gaddpg_bn_finalize_bwd_impl(stats, C, t.count, t.a, t.b, t.o0, t.o1, t.o2, t.dgamma, t.dbeta, t.accumulate, stream);
  return GADDPG_OK;
}

int gaddpg_bn_running_update_impl(float* running, const float* staged, long long n, float momentum, long long* nbt, int n_layers,
                                  void* stream) {
  GADDPG_CHECK_ARG(running && staged && n >= 1 && n_layers >= 0 && n_layers <= 256, "bn_running_update: bad argument");
  bn_running_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(running, staged, n, momentum, nbt, n_layers);
  GADDPG_CHECK_LAUNCH("bn_running_update_kernel");
  return GADDPG_OK;
}

int gaddpg_bn_finalize_bwd_impl(const float* stats, int C, double count, const float* gamma, const float* rstd, float* g,
                                float* m1, float* m2, float* dgamma, float* dbeta, int accumulate, void* stream) {
  GADDPG_CHECK_ARG(C >= 1 && stats && gamma && rstd && g && m1 && m2 && count >= 1.0, "bn_finalize_bwd: bad argument");
  bn_finalize_bwd_kernel<<<ceil_div(C, 32), 1024, 0, (cudaStream_t)stream>>>(stats, C, count, gamma, rstd, g, m1, m2,
                                                                            dgamma, dbeta, accumulate);
  GADDPG_CHECK_LAUNCH("bn_finalize_bwd_kernel");
  return GADDPG_OK;
}
