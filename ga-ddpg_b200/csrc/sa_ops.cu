// sa_ops.cu — set-abstraction glue kernels around the row-GEMMs (sm_100a).
//
// Replaces, for the compact (duplicate-folded) row list, what upstream QueryAndGroup/GroupAll +
// torch.cat + F.max_pool2d do with materialised (B,C,npoint,nsample) tensors (SURVEY.md §8 Spec S3,
// rows a9/a10; called from /root/reference/core/networks.py:217-220):
//   sa1_l1_fwd/bwd   first shared-MLP layer of SA1 straight from the channel-major cloud: K = 3+C is tiny
//                    (7 or 13), so the grouped input is never built; channels that are constant per sample
//                    (the action broadcast by concat_state_action_channelwise, utils.py:291-297) enter as a
//                    per-sample bias and get their gradient from per-sample column sums.
//   gather_rows      [feats | rel. xyz | 0-pad] rows for SA2 (ball groups) and SA3 (GroupAll)
//   scatter_rows     deterministic group_points_grad for the compact rows (one CTA per sample, fixed order)
//   pool_fwd/bwd     max over the neighbourhood of relu(bn(y)) with arg-max, and its backward
//   feat_finish      relu(bn(y)) of the FC head + the time column -> (B,516) policy/critic input
#include "common.cuh"
#include "gemm_rows.cuh"
#include "impl.h"

namespace {

constexpr int SA1_CO = 64;   // SA1 first-layer width (networks.py:70: mlp=[in, 64, 64, 128])
constexpr int SA1_KMAX = 16; // 3 + per-point channels

// One CTA = 256 threads = 16 rows x 16 channel-quads per pass.  W row n holds [dxyz(3) | per-point(Cp) | bcast(Cb)].
__global__ void __launch_bounds__(256) sa1_l1_fwd_kernel(const float* __restrict__ cloud, long long cloud_sb, int cloud_sc,
                                                         int skip, int Cp, const float* __restrict__ ctr, int npoint,
                                                         const int32_t* __restrict__ row_seg,
                                                         const int32_t* __restrict__ row_src,
                                                         const float* __restrict__ row_w, int M_max,
                                                         const int* __restrict__ M_dev, const float* __restrict__ W,
                                                         int ldw, const float* __restrict__ bcbias,
                                                         float* __restrict__ Y, float* __restrict__ stats) {
  __shared__ float sW[SA1_CO * SA1_KMAX];
  __shared__ float red[2 * 16 * SA1_CO];
  const int tid = threadIdx.x, q = tid & 15, rl = tid >> 4;
  const int K1 = 3 + Cp;
  for (int e = tid; e < SA1_CO * SA1_KMAX; e += 256) {
    int n = e / SA1_KMAX, k = e % SA1_KMAX;
    sW[e] = k < K1 ? W[n * ldw + k] : 0.f;
  }
  __syncthreads();
  float w[4][SA1_KMAX];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < SA1_KMAX; ++k) w[c][k] = sW[(q * 4 + c) * SA1_KMAX + k];
  int M = M_dev ? *M_dev : M_max;
  M = M < M_max ? M : M_max;
  float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
  constexpr int UR = 4;  // rows in flight per thread: indices -> cloud gathers -> FMAs, each stage issued for all UR rows
  const int rstep = gridDim.x * 16;
  for (int r0 = blockIdx.x * 16 + rl; r0 < M; r0 += rstep * UR) {
    int seg[UR], src[UR];
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const int r = r0 + u * rstep;
      seg[u] = r < M ? row_seg[r] : 0;
      src[u] = r < M ? row_src[r] : 0;
    }
    float in[UR][SA1_KMAX];
    float cx[UR][3], rwv[UR];
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const int b = seg[u] / npoint;
      const float* pc = cloud + (long long)b * cloud_sb + skip + src[u];
#pragma unroll
      for (int k = 0; k < SA1_KMAX; ++k) in[u][k] = 0.f;
#pragma unroll
      for (int k = 0; k < SA1_KMAX - 3; ++k)
        if (k < Cp) in[u][3 + k] = pc[(long long)k * cloud_sc];
#pragma unroll
      for (int d = 0; d < 3; ++d) cx[u][d] = ctr[(long long)seg[u] * 3 + d];
      rwv[u] = (stats && r0 + u * rstep < M) ? row_w[r0 + u * rstep] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const int r = r0 + u * rstep;
      if (r >= M) continue;
      const int b = seg[u] / npoint;
      in[u][0] = in[u][3] - cx[u][0];
      in[u][1] = in[u][4] - cx[u][1];
      in[u][2] = in[u][5] - cx[u][2];
      float4 o;
      float* op = &o.x;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float a = bcbias ? bcbias[(long long)b * SA1_CO + q * 4 + c] : 0.f;
#pragma unroll
        for (int k = 0; k < SA1_KMAX; ++k) a = fmaf(w[c][k], in[u][k], a);
        op[c] = a;
      }
      *reinterpret_cast<float4*>(Y + (long long)r * SA1_CO + q * 4) = o;
      if (stats) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          s0[c] = fmaf(rwv[u], op[c], s0[c]);
          s1[c] = fmaf(rwv[u] * op[c], op[c], s1[c]);
        }
      }
    }
  }
  if (stats) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      red[rl * SA1_CO + q * 4 + c] = s0[c];
      red[16 * SA1_CO + rl * SA1_CO + q * 4 + c] = s1[c];
    }
    __syncthreads();
    float v = 0.f;
    if (tid < 2 * SA1_CO) {
      int which = tid / SA1_CO, c = tid % SA1_CO;
      for (int y = 0; y < 16; ++y) v += red[which * 16 * SA1_CO + y * SA1_CO + c];
    }
    for (int slot = blockIdx.x; slot < GADDPG_STAT_SLOTS; slot += gridDim.x)
      if (tid < 2 * SA1_CO) stats[(long long)slot * 2 * SA1_CO + tid] = (slot == (int)blockIdx.x) ? v : 0.f;
  }
}

// per-sample bias of the broadcast channels: bcbias[b][n] = sum_c W[n][3+Cp+c] * bc[b][c]
__global__ void sa1_bcbias_kernel(const float* __restrict__ bc, int Cb, int B, const float* __restrict__ W, int ldw,
                                  int koff, float* __restrict__ bcbias) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * SA1_CO) return;
  int b = e / SA1_CO, n = e % SA1_CO;
  float a = 0.f;
  for (int c = 0; c < Cb; ++c) a = fmaf(W[n * ldw + koff + c], bc[b * Cb + c], a);
  bcbias[e] = a;
}

// dW1[n][k] partials: thread = (channel n, row lane rl of 4); dY1 = BN-backward of (D1, Y1) on the fly
__global__ void __launch_bounds__(256) sa1_l1_bwd_kernel(const float* __restrict__ cloud, long long cloud_sb, int cloud_sc,
                                                         int skip, int Cp, const float* __restrict__ ctr, int npoint,
                                                         const int32_t* __restrict__ row_seg,
                                                         const int32_t* __restrict__ row_src,
                                                         const float* __restrict__ row_w, int M_max,
                                                         const int* __restrict__ M_dev, const float* __restrict__ D,
                                                         const float* __restrict__ Y, const float* __restrict__ g,
                                                         const float* __restrict__ m1, const float* __restrict__ m2,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         float* __restrict__ dY_out, float* __restrict__ partial) {
  __shared__ float red[4 * SA1_CO * SA1_KMAX];
  const int tid = threadIdx.x, n = tid & 63, rl = tid >> 6;
  const float gn = g[n], m1n = m1[n], m2n = m2[n], mun = mean[n], rsn = rstd[n];
  int M = M_dev ? *M_dev : M_max;
  M = M < M_max ? M : M_max;
  float acc[SA1_KMAX];
#pragma unroll
  for (int k = 0; k < SA1_KMAX; ++k) acc[k] = 0.f;
  constexpr int UR = 8;  // rows in flight per thread
  const int rstep = gridDim.x * 4;
  for (int r0 = blockIdx.x * 4 + rl; r0 < M; r0 += rstep * UR) {
    int seg[UR], src[UR];
    float d[UR], y[UR], rwv[UR];
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const int r = r0 + u * rstep;
      const bool ok = r < M;
      seg[u] = ok ? row_seg[r] : 0;
      src[u] = ok ? row_src[r] : 0;
      d[u] = ok ? D[(long long)r * SA1_CO + n] : 0.f;
      y[u] = ok ? Y[(long long)r * SA1_CO + n] : 0.f;
      rwv[u] = ok ? row_w[r] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < UR; ++u) {
      const int r = r0 + u * rstep;
      if (r >= M) continue;
      const int b = seg[u] / npoint;
      const float* pc = cloud + (long long)b * cloud_sb + skip + src[u];
      float in[SA1_KMAX];
#pragma unroll
      for (int k = 0; k < SA1_KMAX; ++k) in[k] = 0.f;
#pragma unroll
      for (int k = 0; k < SA1_KMAX - 3; ++k)
        if (k < Cp) in[3 + k] = pc[(long long)k * cloud_sc];
      in[0] = in[3] - ctr[(long long)seg[u] * 3 + 0];
      in[1] = in[4] - ctr[(long long)seg[u] * 3 + 1];
      in[2] = in[5] - ctr[(long long)seg[u] * 3 + 2];
      const float dy = gn * (d[u] - rwv[u] * (m1n + (y[u] - mun) * rsn * m2n));
      if (dY_out) dY_out[(long long)r * SA1_CO + n] = dy;
#pragma unroll
      for (int k = 0; k < SA1_KMAX; ++k) acc[k] = fmaf(dy, in[k], acc[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < SA1_KMAX; ++k) red[(rl * SA1_CO + n) * SA1_KMAX + k] = acc[k];
  __syncthreads();
  for (int e = tid; e < SA1_CO * SA1_KMAX; e += 256) {
    float v = red[e] + red[SA1_CO * SA1_KMAX + e] + red[2 * SA1_CO * SA1_KMAX + e] + red[3 * SA1_CO * SA1_KMAX + e];
    partial[(long long)blockIdx.x * SA1_CO * SA1_KMAX + e] = v;
  }
}
__global__ void sa1_dw_reduce_kernel(const float* __restrict__ partial, int nblk, int K1, float* __restrict__ dW, int ldw,
                                     int accumulate) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= SA1_CO * K1) return;
  int n = e / K1, k = e % K1;
  float s = 0.f;
  for (int b = 0; b < nblk; ++b) s += partial[((long long)b * SA1_CO + n) * SA1_KMAX + k];
  float* d = dW + n * ldw + k;
  *d = (accumulate ? *d : 0.f) + s;
}

// per-sample column sums of dY1 (rows of a sample are contiguous), then dbc = Wbc^T colsum, dWbc += colsum (x) bc
__global__ void __launch_bounds__(64) sa1_dbc_kernel(const float* __restrict__ dY, const int32_t* __restrict__ seg_off,
                                                     int npoint, const float* __restrict__ W, int ldw, int koff, int Cb,
                                                     float* __restrict__ colsum, float* __restrict__ dbc) {
  __shared__ float s[SA1_CO];
  const int b = blockIdx.x, n = threadIdx.x;
  const int r0 = seg_off[b * npoint], r1 = seg_off[(b + 1) * npoint];
  float a = 0.f;
  for (int r = r0; r < r1; ++r) a += dY[(long long)r * SA1_CO + n];
  s[n] = a;
  if (colsum) colsum[(long long)b * SA1_CO + n] = a;
  __syncthreads();
  if (dbc && n < Cb) {
    float v = 0.f;
    for (int m = 0; m < SA1_CO; ++m) v = fmaf(W[m * ldw + koff + n], s[m], v);
    dbc[(long long)b * Cb + n] = v;
  }
}
// dWbc[n][c] (+)= sum_b colsum[b][n] * bc[b][c]   (fixed order over b)
__global__ void sa1_dwbc_kernel(const float* __restrict__ colsum, const float* __restrict__ bc, int B, int Cb,
                                float* __restrict__ dW, int ldw, int koff, int accumulate) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= SA1_CO * Cb) return;
  int n = e / Cb, c = e % Cb;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s = fmaf(colsum[(long long)b * SA1_CO + n], bc[(long long)b * Cb + c], s);
  float* d = dW + n * ldw + koff + c;
  *d = (accumulate ? *d : 0.f) + s;
}

// G[r] = [ feats[b, src, 0:C] | xyz[b, src] - ctr[seg] | 0 ... ]   (row tables NULL: r is the point itself, no centre)
__global__ void gather_rows_kernel(const float* __restrict__ feats, int C, const float* __restrict__ xyz, int n_src,
                                   const float* __restrict__ ctr, int npoint, const int32_t* __restrict__ row_seg,
                                   const int32_t* __restrict__ row_src, int M_max, const int* __restrict__ M_dev,
                                   float* __restrict__ G, int ldg) {
  int M = M_dev ? *M_dev : M_max;
  M = M < M_max ? M : M_max;
  const int lanes = ldg / 4;  // float4 lanes per row
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)M * lanes;
       e += (long long)gridDim.x * blockDim.x) {
    int r = (int)(e / lanes), c = (int)(e % lanes) * 4;
    long long p;  // global source point
    int seg = -1;
    if (row_seg) {
      seg = row_seg[r];
      p = (long long)(seg / npoint) * n_src + row_src[r];
    } else {
      p = r;
    }
    float4 v;
    if (c + 3 < C) {
      v = *reinterpret_cast<const float4*>(feats + p * C + c);
    } else {
      float t[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int cc = c + i;
        float x = 0.f;
        if (cc < C) {
          x = feats[p * C + cc];
        } else if (cc < C + 3) {
          x = xyz[p * 3 + (cc - C)];
          if (ctr) x -= ctr[(long long)seg * 3 + (cc - C)];
        }
        t[i] = x;
      }
      v = make_float4(t[0], t[1], t[2], t[3]);
    }
    *reinterpret_cast<float4*>(G + (long long)r * ldg + c) = v;
  }
}

// dfeats[b, src, c] = sum over the sample's rows with that source, in row order (deterministic).
// One CTA per sample, thread per channel; n_src*C floats of shared memory.
__global__ void scatter_rows_kernel(const float* __restrict__ dG, int ldg, int C, int n_src, int npoint,
                                    const int32_t* __restrict__ seg_off, const int32_t* __restrict__ row_src,
                                    float* __restrict__ dfeats) {
  extern __shared__ float sacc[];
  const int b = blockIdx.x;
  for (int e = threadIdx.x; e < n_src * C; e += blockDim.x) sacc[e] = 0.f;
  __syncthreads();
  const int r0 = seg_off[b * npoint], r1 = seg_off[(b + 1) * npoint];
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    for (int r = r0; r < r1; ++r) sacc[row_src[r] * C + c] += dG[(long long)r * ldg + c];
  __syncthreads();
  for (int e = threadIdx.x; e < n_src * C; e += blockDim.x) dfeats[(long long)b * n_src * C + e] = sacc[e];
}

// out[seg][c] = max over the segment's rows of relu(y*scale+shift); arg[seg][c] = first row attaining it.
// seg_off NULL: fixed segments of `fixed_len` rows.
__global__ void pool_fwd_kernel(const float* __restrict__ Y, int C, const float* __restrict__ scale,
                                const float* __restrict__ shift, const int32_t* __restrict__ seg_off, int fixed_len,
                                int S, float* __restrict__ out, int32_t* __restrict__ arg) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)S * C;
       e += (long long)gridDim.x * blockDim.x) {
    int seg = (int)(e / C), c = (int)(e % C);
    int r0 = seg_off ? seg_off[seg] : seg * fixed_len;
    int r1 = seg_off ? seg_off[seg + 1] : r0 + fixed_len;
    const float sc = scale[c], sh = shift[c];
    float best = -1.f;
    int bi = r0;
    for (int r = r0; r < r1; ++r) {
      float v = fmaxf(fmaf(Y[(long long)r * C + c], sc, sh), 0.f);
      if (v > best) {
        best = v;
        bi = r;
      }
    }
    out[e] = best;
    if (arg) arg[e] = bi;
  }
}

// D[r][c] = dOut[seg][c] if r is the arg-max row and the pooled value is > 0, else 0; BN-backward sums.
// A thread owns 4 consecutive channels (float4) of a fixed channel group and walks rows; CL = C/4 lanes cover a row,
// 256/CL rows are processed per pass and UR passes are in flight.  Per-thread partial sums are combined in a fixed
// order at the end (one slot per CTA).
__global__ void __launch_bounds__(256) pool_bwd_kernel(const float* __restrict__ dOut, int ldo, const float* __restrict__ out,
                                                       const int32_t* __restrict__ arg, const float* __restrict__ Y,
                                                       int C, const int32_t* __restrict__ row_seg, int fixed_len,
                                                       int M_max, const int* __restrict__ M_dev,
                                                       const float* __restrict__ mean, const float* __restrict__ rstd,
                                                       float* __restrict__ D, float* __restrict__ stats) {
  extern __shared__ float sst[];  // [256 threads][8] partial sums
  const int tid = threadIdx.x;
  int M = M_dev ? *M_dev : M_max;
  M = M < M_max ? M : M_max;
  const int CL = C >> 2;            // float4 lanes per row (C % 64 == 0 -> CL in {16,32,64,128,256})
  const int cl = tid % CL, rl = tid / CL;
  const int RPP = 256 / CL;         // rows per pass per CTA (CL <= 256)
  float a0[4] = {0.f, 0.f, 0.f, 0.f}, a1[4] = {0.f, 0.f, 0.f, 0.f};
  if (rl < RPP) {
    const int c = cl << 2;
    const float4 mu = *reinterpret_cast<const float4*>(mean + c), rs = *reinterpret_cast<const float4*>(rstd + c);
    constexpr int UR = 4;
    const int rstep = gridDim.x * RPP;
    for (int r0 = blockIdx.x * RPP + rl; r0 < M; r0 += rstep * UR) {
      int seg[UR];
      float4 y[UR];
#pragma unroll
      for (int u = 0; u < UR; ++u) {
        const int r = r0 + u * rstep;
        seg[u] = r < M ? (row_seg ? row_seg[r] : r / fixed_len) : 0;
        y[u] = r < M ? *reinterpret_cast<const float4*>(Y + (long long)r * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      int4 ag[UR];
      float4 po[UR], dg[UR];
#pragma unroll
      for (int u = 0; u < UR; ++u) {
        const long long pe = (long long)seg[u] * C + c;
        ag[u] = *reinterpret_cast<const int4*>(arg + pe);
        po[u] = *reinterpret_cast<const float4*>(out + pe);
        dg[u] = *reinterpret_cast<const float4*>(dOut + (long long)seg[u] * ldo + c);
      }
#pragma unroll
      for (int u = 0; u < UR; ++u) {
        const int r = r0 + u * rstep;
        if (r >= M) continue;
        float4 d;
        d.x = (ag[u].x == r && po[u].x > 0.f) ? dg[u].x : 0.f;
        d.y = (ag[u].y == r && po[u].y > 0.f) ? dg[u].y : 0.f;
        d.z = (ag[u].z == r && po[u].z > 0.f) ? dg[u].z : 0.f;
        d.w = (ag[u].w == r && po[u].w > 0.f) ? dg[u].w : 0.f;
        *reinterpret_cast<float4*>(D + (long long)r * C + c) = d;
        a0[0] += d.x; a0[1] += d.y; a0[2] += d.z; a0[3] += d.w;
        a1[0] = fmaf(d.x, (y[u].x - mu.x) * rs.x, a1[0]);
        a1[1] = fmaf(d.y, (y[u].y - mu.y) * rs.y, a1[1]);
        a1[2] = fmaf(d.z, (y[u].z - mu.z) * rs.z, a1[2]);
        a1[3] = fmaf(d.w, (y[u].w - mu.w) * rs.w, a1[3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    sst[tid * 8 + i] = a0[i];
    sst[tid * 8 + 4 + i] = a1[i];
  }
  __syncthreads();
  // column c = 4*cl + i is held by threads (rl', cl) for rl' < RPP: sum them in rl' order
  for (int e = tid; e < 2 * C; e += 256) {
    const int which = e / C, c = e % C;
    const int l = c >> 2, i = c & 3;
    float v = 0.f;
    for (int rr = 0; rr < RPP; ++rr) v += sst[(rr * CL + l) * 8 + which * 4 + i];
    stats[(long long)blockIdx.x * 2 * C + e] = v;
  }
  for (int slot = blockIdx.x + gridDim.x; slot < GADDPG_STAT_SLOTS; slot += gridDim.x)
    for (int e = tid; e < 2 * C; e += 256) stats[(long long)slot * 2 * C + e] = 0.f;
}

// feat[b][0:C] = relu(y*scale+shift), feat[b][C] = time[b], feat[b][C+1:ld] = 0
__global__ void feat_finish_kernel(const float* __restrict__ Y, int C, const float* __restrict__ scale,
                                   const float* __restrict__ shift, const float* __restrict__ time, float time_offset,
                                   int B, float* __restrict__ feat, int ld) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)B * ld;
       e += (long long)gridDim.x * blockDim.x) {
    int b = (int)(e / ld), c = (int)(e % ld);
    float v = 0.f;
    if (c < C)
      v = fmaxf(fmaf(Y[(long long)b * C + c], scale[c], shift[c]), 0.f);
    else if (c == C && time)
      v = time[b] + time_offset;
    feat[e] = v;
  }
}

// D = dX * [Yprev*psc+psh > 0] and the BN-backward sums (sum D, sum D*xhat_prev) -- the stand-alone form of the
// EPI_DMASK epilogue, for gradients that arrive from outside (autograd plug-in mode).  One CTA per 64 columns.
__global__ void __launch_bounds__(256) dmask_stats_kernel(const float* __restrict__ dX, int ldx, const float* __restrict__ Yprev,
                                                          int C, int M, const float* __restrict__ psc,
                                                          const float* __restrict__ psh, const float* __restrict__ pmean,
                                                          const float* __restrict__ prstd, float* __restrict__ D,
                                                          float* __restrict__ stats) {
  __shared__ float red[512];
  const int tid = threadIdx.x, cl = tid & 63, rl = tid >> 6;
  const int c = blockIdx.x * 64 + cl;
  float a0 = 0.f, a1 = 0.f;
  if (c < C) {
    const float sc = psc[c], sh = psh[c], mu = pmean[c], rs = prstd[c];
    for (int r = rl; r < M; r += 4) {
      float yp = Yprev[(long long)r * C + c];
      float d = fmaf(yp, sc, sh) > 0.f ? dX[(long long)r * ldx + c] : 0.f;
      D[(long long)r * C + c] = d;
      a0 += d;
      a1 = fmaf(d, (yp - mu) * rs, a1);
    }
  }
  red[rl * 64 + cl] = a0;
  red[256 + rl * 64 + cl] = a1;
  __syncthreads();
  if (tid < 64 && c < C) {
    stats[c] = red[cl] + red[64 + cl] + red[128 + cl] + red[192 + cl];
    stats[C + c] = red[256 + cl] + red[320 + cl] + red[384 + cl] + red[448 + cl];
  }
  // remaining slots are zero
  for (long long e = (long long)blockIdx.x * 256 + tid; e < (long long)(GADDPG_STAT_SLOTS - 1) * 2 * C; e += (long long)gridDim.x * 256)
    stats[2 * C + e] = 0.f;
}

}  // namespace

int gaddpg_dmask_stats_impl(const float* dX, int ldx, const float* Yprev, int C, int M, const float* psc, const float* psh,
                            const float* pmean, const float* prstd, float* D, float* stats, void* stream) {
  GADDPG_CHECK_ARG(dX && Yprev && psc && psh && pmean && prstd && D && stats && C >= 1 && ldx >= C, "dmask_stats: bad argument");
  dmask_stats_kernel<<<ceil_div(C, 64), 256, 0, (cudaStream_t)stream>>>(dX, ldx, Yprev, C, M, psc, psh, pmean, prstd, D, stats);
  GADDPG_CHECK_LAUNCH("dmask_stats_kernel");
  return GADDPG_OK;
}

int gaddpg_sa1_l1_fwd_impl(const float* cloud, long long cloud_sb, int cloud_sc, int skip, int Cp, const float* bc, int Cb,
                           int B, const float* ctr, int npoint, const int32_t* row_seg, const int32_t* row_src,
                           const float* row_w, int M_max, const int* M_dev, const float* W, int ldw, float* bcbias_ws,
                           float* Y, float* stats, void* stream) {
  GADDPG_CHECK_ARG(cloud && ctr && row_seg && row_src && row_w && W && Y, "sa1_l1_fwd: null pointer");
  GADDPG_CHECK_ARG(Cp >= 3 && 3 + Cp <= SA1_KMAX && Cb >= 0 && ldw >= 3 + Cp + Cb, "sa1_l1_fwd: bad channels Cp=%d Cb=%d", Cp, Cb);
  GADDPG_CHECK_ARG(Cb == 0 || (bc && bcbias_ws), "sa1_l1_fwd: broadcast channels need bc and workspace");
  cudaStream_t st = (cudaStream_t)stream;
  if (Cb > 0) {
    sa1_bcbias_kernel<<<ceil_div(B * SA1_CO, 256), 256, 0, st>>>(bc, Cb, B, W, ldw, 3 + Cp, bcbias_ws);
    GADDPG_CHECK_LAUNCH("sa1_bcbias_kernel");
  }
  if (M_max == 0) return GADDPG_OK;
  int grid = ceil_div(M_max, 16);
  grid = grid < GADDPG_STAT_SLOTS ? grid : GADDPG_STAT_SLOTS;
  sa1_l1_fwd_kernel<<<grid, 256, 0, st>>>(cloud, cloud_sb, cloud_sc, skip, Cp, ctr, npoint, row_seg, row_src, row_w, M_max,
                                          M_dev, W, ldw, Cb > 0 ? bcbias_ws : nullptr, Y, stats);
  GADDPG_CHECK_LAUNCH("sa1_l1_fwd_kernel");
  return GADDPG_OK;
}

int gaddpg_sa1_l1_bwd_impl(const float* cloud, long long cloud_sb, int cloud_sc, int skip, int Cp, const float* bc, int Cb,
                           int B, const float* ctr, int npoint, const int32_t* seg_off, const int32_t* row_seg,
                           const int32_t* row_src, const float* row_w, int M_max, const int* M_dev, const float* D,
                           const float* Y, const float* g, const float* m1, const float* m2, const float* mean,
                           const float* rstd, const float* W, int ldw, float* dW, int accumulate, float* dbc, float* dY_ws,
                           float* ws, size_t ws_bytes, void* stream) {
  GADDPG_CHECK_ARG(cloud && ctr && seg_off && row_seg && row_src && row_w && D && Y && g && m1 && m2 && mean && rstd && W && ws,
                   "sa1_l1_bwd: null pointer");
  GADDPG_CHECK_ARG(Cp >= 3 && 3 + Cp <= SA1_KMAX && Cb >= 0, "sa1_l1_bwd: bad channels");
  GADDPG_CHECK_ARG(Cb == 0 || dY_ws, "sa1_l1_bwd: broadcast channels need the dY workspace");
  const int nblk = GADDPG_STAT_SLOTS;
  size_t need = ((size_t)nblk * SA1_CO * SA1_KMAX + (size_t)B * SA1_CO) * sizeof(float);
  GADDPG_CHECK_ARG(ws_bytes >= need, "sa1_l1_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (M_max == 0) return GADDPG_OK;
  float* colsum = ws + (size_t)nblk * SA1_CO * SA1_KMAX;
  sa1_l1_bwd_kernel<<<nblk, 256, 0, st>>>(cloud, cloud_sb, cloud_sc, skip, Cp, ctr, npoint, row_seg, row_src, row_w, M_max,
                                          M_dev, D, Y, g, m1, m2, mean, rstd, Cb > 0 ? dY_ws : nullptr, ws);
  GADDPG_CHECK_LAUNCH("sa1_l1_bwd_kernel");
  if (dW) {
    sa1_dw_reduce_kernel<<<ceil_div(SA1_CO * (3 + Cp), 128), 128, 0, st>>>(ws, nblk, 3 + Cp, dW, ldw, accumulate);
    GADDPG_CHECK_LAUNCH("sa1_dw_reduce_kernel");
  }
  if (Cb > 0) {
    sa1_dbc_kernel<<<B, 64, 0, st>>>(dY_ws, seg_off, npoint, W, ldw, 3 + Cp, Cb, colsum, dbc);
    GADDPG_CHECK_LAUNCH("sa1_dbc_kernel");
    if (dW) {
      sa1_dwbc_kernel<<<ceil_div(SA1_CO * Cb, 128), 128, 0, st>>>(colsum, bc, B, Cb, dW, ldw, 3 + Cp, accumulate);
      GADDPG_CHECK_LAUNCH("sa1_dwbc_kernel");
    }
  }
  return GADDPG_OK;
}

int gaddpg_gather_rows_impl(const float* feats, int C, const float* xyz, int n_src, const float* ctr, int npoint,
                            const int32_t* row_seg, const int32_t* row_src, int M_max, const int* M_dev, float* G, int ldg,
                            void* stream) {
  GADDPG_CHECK_ARG(feats && xyz && G && (C % 4) == 0 && ldg >= C + 3 && (ldg % 4) == 0, "gather_rows: bad argument");
  GADDPG_CHECK_ARG((row_seg == nullptr) == (row_src == nullptr) && (!ctr || row_seg), "gather_rows: inconsistent row tables");
  if (M_max == 0) return GADDPG_OK;
  long long work = (long long)M_max * (ldg / 4);
  int grid = (int)((work + 255) / 256 < 148 * 8 ? (work + 255) / 256 : 148 * 8);
  gather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(feats, C, xyz, n_src, ctr, npoint, row_seg, row_src, M_max, M_dev,
                                                            G, ldg);
  GADDPG_CHECK_LAUNCH("gather_rows_kernel");
  return GADDPG_OK;
}

int gaddpg_scatter_rows_impl(const float* dG, int ldg, int C, int B, int n_src, int npoint, const int32_t* seg_off,
                             const int32_t* row_src, float* dfeats, void* stream) {
  GADDPG_CHECK_ARG(dG && seg_off && row_src && dfeats && C >= 1 && ldg >= C, "scatter_rows: bad argument");
  size_t smem = (size_t)n_src * C * sizeof(float);
  GADDPG_CHECK_ARG(smem <= 200 * 1024, "scatter_rows: n_src*C too large for shared memory");
  if (B == 0) return GADDPG_OK;
  if (smem > 48 * 1024)
    GADDPG_CUDA(cudaFuncSetAttribute(scatter_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  scatter_rows_kernel<<<B, 128, smem, (cudaStream_t)stream>>>(dG, ldg, C, n_src, npoint, seg_off, row_src, dfeats);
  GADDPG_CHECK_LAUNCH("scatter_rows_kernel");
  return GADDPG_OK;
}

int gaddpg_pool_fwd_impl(const float* Y, int C, const float* scale, const float* shift, const int32_t* seg_off, int fixed_len,
                         int S, float* out, int32_t* arg, void* stream) {
  GADDPG_CHECK_ARG(Y && scale && shift && out && C >= 1 && S >= 0 && (seg_off || fixed_len >= 1), "pool_fwd: bad argument");
  if (S == 0) return GADDPG_OK;
  long long work = (long long)S * C;
  int grid = (int)((work + 255) / 256 < 148 * 8 ? (work + 255) / 256 : 148 * 8);
  pool_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Y, C, scale, shift, seg_off, fixed_len, S, out, arg);
  GADDPG_CHECK_LAUNCH("pool_fwd_kernel");
  return GADDPG_OK;
}

int gaddpg_pool_bwd_impl(const float* dOut, int ldo, const float* out, const int32_t* arg, const float* Y, int C,
                         const int32_t* row_seg, int fixed_len, int M_max, const int* M_dev, const float* mean,
                         const float* rstd, float* D, float* stats, void* stream) {
  GADDPG_CHECK_ARG(dOut && out && arg && Y && mean && rstd && D && stats, "pool_bwd: null pointer");
  GADDPG_CHECK_ARG((C % 64) == 0 && C <= 1024 && ldo >= C && (row_seg || fixed_len >= 1), "pool_bwd: bad shape C=%d", C);
  if (M_max == 0) return GADDPG_OK;
  GADDPG_CHECK_ARG(ldo % 4 == 0 && ((uintptr_t)dOut % 16) == 0, "pool_bwd: dOut must be 16-byte aligned with ld %% 4 == 0");
  int rpp = 256 / (C / 4);
  int tiles = ceil_div(M_max, rpp * 4);
  int grid = tiles < GADDPG_STAT_SLOTS ? tiles : GADDPG_STAT_SLOTS;
  size_t smem = 256 * 8 * sizeof(float);
  pool_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(dOut, ldo, out, arg, Y, C, row_seg, fixed_len, M_max, M_dev, mean, rstd,
                                                            D, stats);
  GADDPG_CHECK_LAUNCH("pool_bwd_kernel");
  return GADDPG_OK;
}

int gaddpg_feat_finish_impl(const float* Y, int C, const float* scale, const float* shift, const float* time,
                            float time_offset, int B, float* feat, int ld, void* stream) {
  GADDPG_CHECK_ARG(Y && scale && shift && feat && ld >= C + 1, "feat_finish: bad argument");
  if (B == 0) return GADDPG_OK;
  long long work = (long long)B * ld;
  feat_finish_kernel<<<(int)((work + 255) / 256), 256, 0, (cudaStream_t)stream>>>(Y, C, scale, shift, time, time_offset, B, feat,
                                                                                 ld);
  GADDPG_CHECK_LAUNCH("feat_finish_kernel");
  return GADDPG_OK;
}
