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
#include "bn_tail.cuh"

namespace {

constexpr int SA1_CO = 64;   // SA1 first-layer width (networks.py:70: mlp=[in, 64, 64, 128])
constexpr int SA1_KMAX = 16; // 3 + per-point channels

// Work decomposition shared by the forward and backward kernels of this layer: an ITEM is (sample b, block j of 256
// consecutive compact rows of that sample) -- rows of one sample are contiguous in the row table -- so everything that
// is constant per sample (the broadcast-channel bias, the per-sample gradient column sums) is CTA-uniform.  Persistent
// CTAs (grid <= GADDPG_STAT_SLOTS) stride over B*jmax items; inside an item each of the 8 warps owns 32 rows:
//   phase 1  lane = row     : coalesced row-table reads, the 3+Cp scattered cloud reads of that row (L2-resident
//                             cloud, one 4-byte gather per channel), centroid subtraction -> padded smem tile [32][20]
//   phase 2  lane = 2 chans : walks the 32 staged rows (LDS.128 broadcasts), 2x16 FMAs per row against register
//                             weights, one coalesced 256-byte store per row; statistics stay in registers.
constexpr int SA1_ITEM = 256;  // rows per item
constexpr int SA1_LDI = 20;    // staged-input row stride (floats): conflict-free float4 stores for 8 consecutive lanes

__device__ __forceinline__ void sa1_stage_row(const float* __restrict__ cloud, long long cloud_sb, int cloud_sc, int skip, int Cp,
                                              const float* __restrict__ ctr, const int32_t* __restrict__ row_seg,
                                              const int32_t* __restrict__ row_src, const float* __restrict__ row_w, int b, int r,
                                              bool valid, float* __restrict__ dst, float* __restrict__ wdst) {
  float in[SA1_KMAX];
#pragma unroll
  for (int k = 0; k < SA1_KMAX; ++k) in[k] = 0.f;
  float w = 0.f;
  if (valid) {
    const int seg = row_seg[r], src = row_src[r];
    w = row_w[r];
    const float* pc = cloud + (long long)b * cloud_sb + skip + src;
#pragma unroll
    for (int k = 0; k < SA1_KMAX - 3; ++k)
      if (k < Cp) in[3 + k] = pc[(long long)k * cloud_sc];
    in[0] = in[3] - ctr[(long long)seg * 3 + 0];
    in[1] = in[4] - ctr[(long long)seg * 3 + 1];
    in[2] = in[5] - ctr[(long long)seg * 3 + 2];
  }
#pragma unroll
  for (int k = 0; k < SA1_KMAX; k += 4) *reinterpret_cast<float4*>(dst + k) = make_float4(in[k], in[k + 1], in[k + 2], in[k + 3]);
  *wdst = w;
}

__global__ void __launch_bounds__(256, 2) sa1_l1_fwd_kernel(const float* __restrict__ cloud, long long cloud_sb, int cloud_sc,
                                                            int skip, int Cp, const float* __restrict__ ctr, int npoint,
                                                            const int32_t* __restrict__ seg_off,
                                                            const int32_t* __restrict__ row_seg,
                                                            const int32_t* __restrict__ row_src,
                                                            const float* __restrict__ row_w, int B, int jmax,
                                                            const float* __restrict__ W, int ldw,
                                                            const float* __restrict__ bcbias, float* __restrict__ Y,
                                                            float* __restrict__ stats, const BNTail tail) {
  __shared__ __align__(16) float sIn[8 * 32 * SA1_LDI];
  __shared__ float sRw[8 * 32];
  __shared__ float red[8 * 2 * SA1_CO];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K1 = 3 + Cp;
  float w0[SA1_KMAX], w1[SA1_KMAX];
#pragma unroll
  for (int k = 0; k < SA1_KMAX; ++k) {
    w0[k] = k < K1 ? W[(2 * lane) * ldw + k] : 0.f;
    w1[k] = k < K1 ? W[(2 * lane + 1) * ldw + k] : 0.f;
  }
  float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
  float* myIn = sIn + warp * 32 * SA1_LDI;
  float* myRw = sRw + warp * 32;
  for (int item = blockIdx.x; item < B * jmax; item += gridDim.x) {
    const int b = item / jmax, j = item - b * jmax;
    const int rend = seg_off[(b + 1) * npoint];
    const int base = seg_off[b * npoint] + j * SA1_ITEM + warp * 32;
    int nrows = rend - base;
    nrows = nrows > 32 ? 32 : nrows;
    if (nrows <= 0) continue;  // warp-uniform
    sa1_stage_row(cloud, cloud_sb, cloud_sc, skip, Cp, ctr, row_seg, row_src, row_w, b, base + lane, lane < nrows,
                  myIn + lane * SA1_LDI, myRw + lane);
    float2 bias = make_float2(0.f, 0.f);
    if (bcbias) bias = *reinterpret_cast<const float2*>(bcbias + (long long)b * SA1_CO + 2 * lane);
    __syncwarp();
#pragma unroll 4
    for (int rr = 0; rr < nrows; ++rr) {
      float in[SA1_KMAX];
#pragma unroll
      for (int k = 0; k < SA1_KMAX; k += 4) {
        const float4 v = *reinterpret_cast<const float4*>(myIn + rr * SA1_LDI + k);
        in[k] = v.x;
        in[k + 1] = v.y;
        in[k + 2] = v.z;
        in[k + 3] = v.w;
      }
      float a0 = bias.x, a1 = bias.y;
#pragma unroll
      for (int k = 0; k < SA1_KMAX; ++k) {
        a0 = fmaf(w0[k], in[k], a0);
        a1 = fmaf(w1[k], in[k], a1);
      }
      *reinterpret_cast<float2*>(Y + (long long)(base + rr) * SA1_CO + 2 * lane) = make_float2(a0, a1);
      if (stats) {
        const float w = myRw[rr];
        s00 = fmaf(w, a0, s00);
        s01 = fmaf(w * a0, a0, s01);
        s10 = fmaf(w, a1, s10);
        s11 = fmaf(w * a1, a1, s11);
      }
    }
    __syncwarp();
  }
  if (stats) {
    red[warp * 2 * SA1_CO + 2 * lane] = s00;
    red[warp * 2 * SA1_CO + 2 * lane + 1] = s10;
    red[warp * 2 * SA1_CO + SA1_CO + 2 * lane] = s01;
    red[warp * 2 * SA1_CO + SA1_CO + 2 * lane + 1] = s11;
    __syncthreads();
    float v = 0.f;
    if (tid < 2 * SA1_CO)
      for (int y = 0; y < 8; ++y) v += red[y * 2 * SA1_CO + tid];
    if (tail.kind != 0) {   // BatchNorm finalize by the last CTA (the staging tile is free)
      if (tid < 2 * SA1_CO) stats[(long long)blockIdx.x * 2 * SA1_CO + tid] = v;
      bnt_run<16>(tail, stats, SA1_CO, (int)gridDim.x, gridDim.x, reinterpret_cast<double*>(sIn), (int)(sizeof(sIn) / sizeof(double)));
    } else {
      for (int slot = blockIdx.x; slot < GADDPG_STAT_SLOTS; slot += gridDim.x)
        if (tid < 2 * SA1_CO) stats[(long long)slot * 2 * SA1_CO + tid] = (slot == (int)blockIdx.x) ? v : 0.f;
    }
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

// Backward of the same layer, same item decomposition: dY1 = BN-backward of (D1, Y1) on the fly (never stored),
// dW1[n][k] += dY1[r][n] * in[r][k] in 2x16 register accumulators per lane, and -- when the layer has broadcast
// channels -- the per-sample column sums of dY1 as one [64] partial per item (fixed-order sum in sa1_dbc_kernel).
#ifndef SA1_BWD_ROWS
#define SA1_BWD_ROWS 8
#endif
__global__ void __launch_bounds__(256, 2) sa1_l1_bwd_kernel(const float* __restrict__ cloud, long long cloud_sb, int cloud_sc,
                                                            int skip, int Cp, const float* __restrict__ ctr, int npoint,
                                                            const int32_t* __restrict__ seg_off,
                                                            const int32_t* __restrict__ row_seg,
                                                            const int32_t* __restrict__ row_src,
                                                            const float* __restrict__ row_w, int B, int jmax,
                                                            const float* __restrict__ D, const float* __restrict__ Y,
                                                            const float* __restrict__ g, const float* __restrict__ m1,
                                                            const float* __restrict__ m2, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, float* __restrict__ colsum_part,
                                                            float* __restrict__ partial) {
  __shared__ __align__(16) float sbuf[8 * SA1_CO * SA1_KMAX];  // staged inputs (8*32*20 floats) / final dW reduction
  __shared__ float sRw[8 * 32];
  __shared__ float red[8 * SA1_CO];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float2 gn = *reinterpret_cast<const float2*>(g + 2 * lane), m1n = *reinterpret_cast<const float2*>(m1 + 2 * lane),
               m2n = *reinterpret_cast<const float2*>(m2 + 2 * lane), mun = *reinterpret_cast<const float2*>(mean + 2 * lane),
               rsn = *reinterpret_cast<const float2*>(rstd + 2 * lane);
  float acc0[SA1_KMAX], acc1[SA1_KMAX];
#pragma unroll
  for (int k = 0; k < SA1_KMAX; ++k) acc0[k] = acc1[k] = 0.f;
  float* myIn = sbuf + warp * 32 * SA1_LDI;
  float* myRw = sRw + warp * 32;
  for (int item = blockIdx.x; item < B * jmax; item += gridDim.x) {
    const int b = item / jmax, j = item - b * jmax;
    const int rend = seg_off[(b + 1) * npoint];
    const int start = seg_off[b * npoint] + j * SA1_ITEM;
    if (start >= rend) {  // CTA-uniform: empty item
      if (colsum_part && tid < SA1_CO) colsum_part[(long long)item * SA1_CO + tid] = 0.f;
      continue;
    }
    const int base = start + warp * 32;
    int nrows = rend - base;
    nrows = nrows > 32 ? 32 : nrows;
    float cs0 = 0.f, cs1 = 0.f;
    if (nrows > 0) {
      sa1_stage_row(cloud, cloud_sb, cloud_sc, skip, Cp, ctr, row_seg, row_src, row_w, b, base + lane, lane < nrows,
                    myIn + lane * SA1_LDI, myRw + lane);
      __syncwarp();
      // SA1_BWD_ROWS rows of D and Y (2 x 8-byte loads per row and lane) are requested before the first one is consumed:
      // the kernel is bound by the latency of these streams, not by the 40 FMAs per row
      for (int r4 = 0; r4 < nrows; r4 += SA1_BWD_ROWS) {
        float2 d[SA1_BWD_ROWS], y[SA1_BWD_ROWS];
#pragma unroll
        for (int u = 0; u < SA1_BWD_ROWS; ++u) {
          const bool ok = r4 + u < nrows;
          d[u] = ok ? *reinterpret_cast<const float2*>(D + (long long)(base + r4 + u) * SA1_CO + 2 * lane) : make_float2(0.f, 0.f);
          y[u] = ok ? *reinterpret_cast<const float2*>(Y + (long long)(base + r4 + u) * SA1_CO + 2 * lane) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < SA1_BWD_ROWS; ++u) {
          if (r4 + u < nrows) {
            const float w = myRw[r4 + u];
            const float dy0 = gn.x * (d[u].x - w * (m1n.x + (y[u].x - mun.x) * rsn.x * m2n.x));
            const float dy1 = gn.y * (d[u].y - w * (m1n.y + (y[u].y - mun.y) * rsn.y * m2n.y));
            cs0 += dy0;
            cs1 += dy1;
#pragma unroll
            for (int k = 0; k < SA1_KMAX; k += 4) {
              const float4 v = *reinterpret_cast<const float4*>(myIn + (r4 + u) * SA1_LDI + k);
              acc0[k] = fmaf(dy0, v.x, acc0[k]);
              acc0[k + 1] = fmaf(dy0, v.y, acc0[k + 1]);
              acc0[k + 2] = fmaf(dy0, v.z, acc0[k + 2]);
              acc0[k + 3] = fmaf(dy0, v.w, acc0[k + 3]);
              acc1[k] = fmaf(dy1, v.x, acc1[k]);
              acc1[k + 1] = fmaf(dy1, v.y, acc1[k + 1]);
              acc1[k + 2] = fmaf(dy1, v.z, acc1[k + 2]);
              acc1[k + 3] = fmaf(dy1, v.w, acc1[k + 3]);
            }
          }
        }
      }
      __syncwarp();
    }
    if (colsum_part) {
      red[warp * SA1_CO + 2 * lane] = cs0;
      red[warp * SA1_CO + 2 * lane + 1] = cs1;
      __syncthreads();
      if (tid < SA1_CO) {
        float v = 0.f;
        for (int y = 0; y < 8; ++y) v += red[y * SA1_CO + tid];
        colsum_part[(long long)item * SA1_CO + tid] = v;
      }
      __syncthreads();
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < SA1_KMAX; k += 4) {
    *reinterpret_cast<float4*>(sbuf + (warp * SA1_CO + 2 * lane) * SA1_KMAX + k) = make_float4(acc0[k], acc0[k + 1], acc0[k + 2], acc0[k + 3]);
    *reinterpret_cast<float4*>(sbuf + (warp * SA1_CO + 2 * lane + 1) * SA1_KMAX + k) = make_float4(acc1[k], acc1[k + 1], acc1[k + 2], acc1[k + 3]);
  }
  __syncthreads();
  for (int e = tid; e < SA1_CO * SA1_KMAX; e += 256) {
    float v = 0.f;
    for (int y = 0; y < 8; ++y) v += sbuf[y * SA1_CO * SA1_KMAX + e];
    partial[(long long)blockIdx.x * SA1_CO * SA1_KMAX + e] = v;
  }
}
__global__ void sa1_dw_reduce_kernel(const float* __restrict__ partial, int nblk, int K1, float* __restrict__ dW, int ldw,
                                     int accumulate) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= SA1_CO * K1) return;
  int n = e / K1, k = e % K1;
  const float s = ordered_sum<32>(nblk, [&](int b) { return partial[((long long)b * SA1_CO + n) * SA1_KMAX + k]; });
  float* d = dW + n * ldw + k;
  *d = (accumulate ? *d : 0.f) + s;
}

// per-sample column sums of dY1 from the per-item partials (fixed order), then dbc = Wbc^T colsum
__global__ void __launch_bounds__(64) sa1_dbc_kernel(const float* __restrict__ colsum_part, int jmax, const float* __restrict__ W,
                                                     int ldw, int koff, int Cb, float* __restrict__ colsum,
                                                     float* __restrict__ dbc) {
  __shared__ float s[SA1_CO];
  const int b = blockIdx.x, n = threadIdx.x;
  const float a = ordered_sum<4>(jmax, [&](int j) { return colsum_part[((long long)b * jmax + j) * SA1_CO + n]; });
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
  int b = 0;
  for (; b + 16 <= B; b += 16) {  // 32 loads in flight, then the FMAs in sample order (same result as the plain loop)
    float x[16], y[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      x[u] = colsum[(long long)(b + u) * SA1_CO + n];
      y[u] = bc[(long long)(b + u) * Cb + c];
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) s = fmaf(x[u], y[u], s);
  }
  for (; b < B; ++b) s = fmaf(colsum[(long long)b * SA1_CO + n], bc[(long long)b * Cb + c], s);
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
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    for (int r = r0; r < r1; r += 8) {  // 8 independent loads in flight, then the adds in row order (deterministic)
      int src[8];
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        src[u] = r + u < r1 ? row_src[r + u] : -1;
        v[u] = r + u < r1 ? dG[(long long)(r + u) * ldg + c] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (src[u] >= 0) sacc[src[u] * C + c] += v[u];
    }
  }
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
    int r = r0;
    for (; r + 8 <= r1; r += 8) {  // 8 row loads in flight per thread (the walk is a latency chain otherwise); same visiting order
      float y[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) y[u] = Y[(long long)(r + u) * C + c];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float v = fmaxf(fmaf(y[u], sc, sh), 0.f);
        if (v > best) {
          best = v;
          bi = r + u;
        }
      }
    }
    for (; r < r1; ++r) {
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

// The same walk with four consecutive channels per thread (C % 4 == 0): a warp reads whole 512-byte row segments with 16-byte
// loads, eight rows in flight per thread — four times the bytes in flight of the scalar kernel for the SA1 pool, which reads
// 217 MB per pass.  Rows are visited in the same order with the same strict comparison, so out / arg are bit-identical.
__global__ void __launch_bounds__(256) pool_fwd4_kernel(const float* __restrict__ Y, int C, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, const int32_t* __restrict__ seg_off,
                                                        int fixed_len, int S, float* __restrict__ out, int32_t* __restrict__ arg) {
  const int CQ = C >> 2;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < (long long)S * CQ; e += (long long)gridDim.x * blockDim.x) {
    const int seg = (int)(e / CQ), c = (int)(e % CQ) << 2;
    const int r0 = seg_off ? seg_off[seg] : seg * fixed_len;
    const int r1 = seg_off ? seg_off[seg + 1] : r0 + fixed_len;
    const float4 sc = *reinterpret_cast<const float4*>(scale + c), sh = *reinterpret_cast<const float4*>(shift + c);
    float b0 = -1.f, b1 = -1.f, b2 = -1.f, b3 = -1.f;
    int i0 = r0, i1 = r0, i2 = r0, i3 = r0;
    auto visit = [&](const float4 y, int row) {
      const float v0 = fmaxf(fmaf(y.x, sc.x, sh.x), 0.f), v1 = fmaxf(fmaf(y.y, sc.y, sh.y), 0.f);
      const float v2 = fmaxf(fmaf(y.z, sc.z, sh.z), 0.f), v3 = fmaxf(fmaf(y.w, sc.w, sh.w), 0.f);
      if (v0 > b0) { b0 = v0; i0 = row; }
      if (v1 > b1) { b1 = v1; i1 = row; }
      if (v2 > b2) { b2 = v2; i2 = row; }
      if (v3 > b3) { b3 = v3; i3 = row; }
    };
    int r = r0;
    for (; r + 8 <= r1; r += 8) {
      float4 y[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) y[u] = *reinterpret_cast<const float4*>(Y + (long long)(r + u) * C + c);
#pragma unroll
      for (int u = 0; u < 8; ++u) visit(y[u], r + u);
    }
    for (; r < r1; ++r) visit(*reinterpret_cast<const float4*>(Y + (long long)r * C + c), r);
    *reinterpret_cast<float4*>(out + (long long)seg * C + c) = make_float4(b0, b1, b2, b3);
    if (arg) *reinterpret_cast<int4*>(arg + (long long)seg * C + c) = make_int4(i0, i1, i2, i3);
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

// Sparse max-pool backward: the dense gradient D (M x C) has one non-zero per (segment, channel) — at the arg-max row, if
// the pooled value survived the ReLU.  E[s][c] holds that value; the BN-backward sums (sum D, sum D*xhat) only need Y at
// the S*C arg-max elements.  Consumers rebuild D on the fly (GADDPG_OP_BNBWD_POOL), so neither D nor a second pass over Y
// ever touches HBM; which rows are arg-max is handed over as one bit per (row, channel) (zeroed by the launcher).
// 256 threads = 256/C segments x C channels per pass; per-thread partials, fixed-order combine, one slot
// per CTA.
__global__ void __launch_bounds__(256) pool_bwd_sparse_kernel(const float* __restrict__ dOut, int ldo, const float* __restrict__ out,
                                                              const int32_t* __restrict__ arg, const float* __restrict__ Y, int C,
                                                              int S, const float* __restrict__ mean,
                                                              const float* __restrict__ rstd, float* __restrict__ E,
                                                              uint32_t* __restrict__ mask, float* __restrict__ stats) {
  __shared__ float red[2][256];
  const int tid = threadIdx.x;
  const int c = tid % C, sl = tid / C, SPP = 256 / C;  // C in {64, 128, 256}
  float a0 = 0.f, a1 = 0.f;
  const float mu = mean[c], rs = rstd[c];
  for (int s0 = blockIdx.x * SPP + sl; s0 < S; s0 += gridDim.x * SPP * 4) {
    float e[4], y[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int s = s0 + u * gridDim.x * SPP;
      e[u] = 0.f;
      y[u] = mu;
      if (s < S) {
        const long long pe = (long long)s * C + c;
        const int r = arg[pe];
        e[u] = out[pe] > 0.f ? dOut[(long long)s * ldo + c] : 0.f;
        y[u] = Y[(long long)r * C + c];
        E[pe] = e[u];
        atomicOr(mask + (long long)r * (C >> 5) + (c >> 5), 1u << (c & 31));  // bit set: order-independent, deterministic
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      a0 += e[u];
      a1 = fmaf(e[u], (y[u] - mu) * rs, a1);
    }
  }
  red[0][tid] = a0;
  red[1][tid] = a1;
  __syncthreads();
  for (int i = tid; i < 2 * C; i += 256) {
    const int which = i / C, cc = i % C;
    float v = 0.f;
    for (int k = 0; k < SPP; ++k) v += red[which][k * C + cc];
    stats[(long long)blockIdx.x * 2 * C + i] = v;
  }
  for (int slot = blockIdx.x + gridDim.x; slot < GADDPG_STAT_SLOTS; slot += gridDim.x)
    for (int i = tid; i < 2 * C; i += 256) stats[(long long)slot * 2 * C + i] = 0.f;
}

// Fused max-pool, second half: keys[seg][c] = {max key, min key} written by the epilogue of the last shared-MLP layer
// (tc_common.cuh: epi_pool32): the per-segment maximum of the raw output for gamma >= 0, the minimum otherwise.  BatchNorm is
// monotone per channel, so relu(bn(extreme)) is bit-identical to pooling the normalised activations.  Keys are re-zeroed
// for the next pass.
__device__ __forceinline__ float key_value(unsigned int hi) {
  return __uint_as_float((hi & 0x80000000u) ? (hi & 0x7FFFFFFFu) : ~hi);
}
__global__ void __launch_bounds__(256) pool_keys_finalize_kernel(unsigned long long* __restrict__ keys, long long total, int C,
                                                                 const float* __restrict__ gamma, const float* __restrict__ scale,
                                                                 const float* __restrict__ shift, float* __restrict__ out,
                                                                 int32_t* __restrict__ arg) {
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const unsigned long long kk = keys[e];
    const unsigned int hi = (unsigned int)(kk >> 32);
    const float v = key_value(gamma[c] < 0.f ? ~hi : hi);
    out[e] = fmaxf(fmaf(v, scale[c], shift[c]), 0.f);
    if (arg) arg[e] = (int32_t)(0xFFFFFFFFu - (unsigned int)(kk & 0xFFFFFFFFull));
    keys[e] = 0ull;
  }
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
                                                          float* __restrict__ stats, const BNTail tail) {
  __shared__ __align__(16) float red[4096];   // [512] partial sums; 16 KB so that it also serves the BatchNorm tail
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
  if (tail.kind != 0) {   // one live slot: the last CTA finalizes all channels
    bnt_run(tail, stats, C, 1, gridDim.x, reinterpret_cast<double*>(red), 2048);
    return;
  }
  // remaining slots are zero
  for (long long e = (long long)blockIdx.x * 256 + tid; e < (long long)(GADDPG_STAT_SLOTS - 1) * 2 * C; e += (long long)gridDim.x * 256)
    stats[2 * C + e] = 0.f;
}

}  // namespace

int gaddpg_dmask_stats_impl(const float* dX, int ldx, const float* Yprev, int C, int M, const float* psc, const float* psh,
                            const float* pmean, const float* prstd, float* D, float* stats, const BNTail* tail, void* stream) {
  GADDPG_CHECK_ARG(dX && Yprev && psc && psh && pmean && prstd && D && stats && C >= 1 && ldx >= C, "dmask_stats: bad argument");
  BNTail t = tail_or_none(tail);
  const bool fused = bnt_fusable(t, C);
  if (!fused) t.kind = 0;
  dmask_stats_kernel<<<ceil_div(C, 64), 256, 0, (cudaStream_t)stream>>>(dX, ldx, Yprev, C, M, psc, psh, pmean, prstd, D, stats, t);
  GADDPG_CHECK_LAUNCH("dmask_stats_kernel");
  return (tail && tail->kind && !fused) ? gaddpg_bn_tail_separate(*tail, stats, C, stream) : GADDPG_OK;
}

static int sa1_jmax(int M_max, int B) { return ceil_div(ceil_div(M_max, B > 0 ? B : 1), SA1_ITEM); }

int gaddpg_sa1_l1_fwd_impl(const float* cloud, long long cloud_sb, int cloud_sc, int skip, int Cp, const float* bc, int Cb,
                           int B, const float* ctr, int npoint, const int32_t* seg_off, const int32_t* row_seg,
                           const int32_t* row_src, const float* row_w, int M_max, const int* M_dev, const float* W, int ldw,
                           float* bcbias_ws, float* Y, float* stats, const BNTail* tail, void* stream) {
  GADDPG_CHECK_ARG(cloud && ctr && seg_off && row_seg && row_src && row_w && W && Y, "sa1_l1_fwd: null pointer");
  GADDPG_CHECK_ARG(!(tail && tail->kind) || stats, "sa1_l1_fwd: a BatchNorm tail needs the statistics buffer");
  GADDPG_CHECK_ARG(Cp >= 3 && 3 + Cp <= SA1_KMAX && Cb >= 0 && ldw >= 3 + Cp + Cb, "sa1_l1_fwd: bad channels Cp=%d Cb=%d", Cp, Cb);
  GADDPG_CHECK_ARG(Cb == 0 || (bc && bcbias_ws), "sa1_l1_fwd: broadcast channels need bc and workspace");
  (void)M_dev;  // the live row count is seg_off[B*npoint]
  cudaStream_t st = (cudaStream_t)stream;
  if (Cb > 0) {
    sa1_bcbias_kernel<<<ceil_div(B * SA1_CO, 256), 256, 0, st>>>(bc, Cb, B, W, ldw, 3 + Cp, bcbias_ws);
    GADDPG_CHECK_LAUNCH("sa1_bcbias_kernel");
  }
  if (M_max == 0 || B == 0) return GADDPG_OK;
  const int jmax = sa1_jmax(M_max, B);
  int grid = B * jmax;
  grid = grid < GADDPG_STAT_SLOTS ? grid : GADDPG_STAT_SLOTS;
  BNTail t = tail_or_none(tail);
  const bool fused = bnt_fusable(t, SA1_CO);
  if (!fused) t.kind = 0;
  sa1_l1_fwd_kernel<<<grid, 256, 0, st>>>(cloud, cloud_sb, cloud_sc, skip, Cp, ctr, npoint, seg_off, row_seg, row_src, row_w, B, jmax,
                                          W, ldw, Cb > 0 ? bcbias_ws : nullptr, Y, stats, t);
  GADDPG_CHECK_LAUNCH("sa1_l1_fwd_kernel");
  return (tail && tail->kind && !fused) ? gaddpg_bn_tail_separate(*tail, stats, SA1_CO, stream) : GADDPG_OK;
}

int gaddpg_sa1_l1_bwd_impl(const float* cloud, long long cloud_sb, int cloud_sc, int skip, int Cp, const float* bc, int Cb,
                           int B, const float* ctr, int npoint, const int32_t* seg_off, const int32_t* row_seg,
                           const int32_t* row_src, const float* row_w, int M_max, const int* M_dev, const float* D,
                           const float* Y, const float* g, const float* m1, const float* m2, const float* mean,
                           const float* rstd, const float* W, int ldw, float* dW, int accumulate, float* dbc, float* ws,
                           size_t ws_bytes, void* stream) {
  GADDPG_CHECK_ARG(cloud && ctr && seg_off && row_seg && row_src && row_w && D && Y && g && m1 && m2 && mean && rstd && W && ws,
                   "sa1_l1_bwd: null pointer");
  GADDPG_CHECK_ARG(Cp >= 3 && 3 + Cp <= SA1_KMAX && Cb >= 0, "sa1_l1_bwd: bad channels");
  (void)M_dev;
  const int nblk = GADDPG_STAT_SLOTS;
  if (M_max == 0 || B == 0) return GADDPG_OK;
  const int jmax = sa1_jmax(M_max, B);
  size_t need = ((size_t)nblk * SA1_CO * SA1_KMAX + (size_t)B * SA1_CO * (1 + jmax)) * sizeof(float);
  GADDPG_CHECK_ARG(ws_bytes >= need, "sa1_l1_bwd: workspace too small (%zu < %zu)", ws_bytes, need);
  cudaStream_t st = (cudaStream_t)stream;
  float* colsum = ws + (size_t)nblk * SA1_CO * SA1_KMAX;
  float* colsum_part = colsum + (size_t)B * SA1_CO;
  int grid = B * jmax;
  grid = grid < nblk ? grid : nblk;
  sa1_l1_bwd_kernel<<<grid, 256, 0, st>>>(cloud, cloud_sb, cloud_sc, skip, Cp, ctr, npoint, seg_off, row_seg, row_src, row_w, B, jmax,
                                          D, Y, g, m1, m2, mean, rstd, Cb > 0 ? colsum_part : nullptr, ws);
  GADDPG_CHECK_LAUNCH("sa1_l1_bwd_kernel");
  if (dW) {
    sa1_dw_reduce_kernel<<<ceil_div(SA1_CO * (3 + Cp), 128), 128, 0, st>>>(ws, grid, 3 + Cp, dW, ldw, accumulate);
    GADDPG_CHECK_LAUNCH("sa1_dw_reduce_kernel");
  }
  if (Cb > 0) {
    sa1_dbc_kernel<<<B, 64, 0, st>>>(colsum_part, jmax, W, ldw, 3 + Cp, Cb, colsum, dbc);
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
  const bool vec4 = (C % 4) == 0 && (((uintptr_t)Y | (uintptr_t)scale | (uintptr_t)shift | (uintptr_t)out | (uintptr_t)arg) & 15u) == 0;
  long long work = vec4 ? (long long)S * (C / 4) : (long long)S * C;
  int grid = (int)((work + 255) / 256 < 148 * 8 ? (work + 255) / 256 : 148 * 8);
  if (vec4)
    pool_fwd4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Y, C, scale, shift, seg_off, fixed_len, S, out, arg);
  else
    pool_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(Y, C, scale, shift, seg_off, fixed_len, S, out, arg);
  GADDPG_CHECK_LAUNCH("pool_fwd_kernel");
  return GADDPG_OK;
}

int gaddpg_pool_bwd_impl(const float* dOut, int ldo, const float* out, const int32_t* arg, const float* Y, int C,
                         const int32_t* row_seg, int fixed_len, int M_max, const int* M_dev, const float* mean,
                         const float* rstd, float* D, float* stats, const BNTail* tail, void* stream) {
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
  // the tail runs as its own launch here: this kernel fills up to 296 full-width slots (C up to 1024), which one CTA cannot sum
  // faster than the C/32-CTA finalize kernel does, and the tail's registers would cost this streaming kernel its occupancy
  return (tail && tail->kind) ? gaddpg_bn_tail_separate(*tail, stats, C, stream) : GADDPG_OK;
}

int gaddpg_pool_bwd_sparse_impl(const float* dOut, int ldo, const float* out, const int32_t* arg, const float* Y, int C, int S,
                                const float* mean, const float* rstd, float* E, uint32_t* mask, int M_max, float* stats,
                                const BNTail* tail, void* stream) {
  GADDPG_CHECK_ARG(dOut && out && arg && Y && mean && rstd && E && mask && stats && M_max >= 1, "pool_bwd_sparse: bad argument");
  GADDPG_CUDA(cudaMemsetAsync(mask, 0, (size_t)M_max * (C / 32) * sizeof(uint32_t), (cudaStream_t)stream));
  GADDPG_CHECK_ARG((C == 64 || C == 128 || C == 256) && ldo >= C && S >= 1, "pool_bwd_sparse: bad shape C=%d S=%d", C, S);
  const int spp = 256 / C;
  int grid = ceil_div(S, spp * 4);
  grid = grid < GADDPG_STAT_SLOTS ? grid : GADDPG_STAT_SLOTS;
  pool_bwd_sparse_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dOut, ldo, out, arg, Y, C, S, mean, rstd, E, mask, stats);
  GADDPG_CHECK_LAUNCH("pool_bwd_sparse_kernel");
  return (tail && tail->kind) ? gaddpg_bn_tail_separate(*tail, stats, C, stream) : GADDPG_OK;   // as in pool_bwd
}

int gaddpg_pool_keys_finalize_impl(unsigned long long* keys, int S, int C, const float* gamma, const float* scale,
                                   const float* shift, float* out, int32_t* arg, void* stream) {
  GADDPG_CHECK_ARG(keys && gamma && scale && shift && out && S >= 1 && C >= 1, "pool_keys_finalize: bad argument");
  const long long total = (long long)S * C;
  const int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  pool_keys_finalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(keys, total, C, gamma, scale, shift, out, arg);
  GADDPG_CHECK_LAUNCH("pool_keys_finalize_kernel");
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
