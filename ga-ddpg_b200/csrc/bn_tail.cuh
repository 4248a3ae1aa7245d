// bn_tail.cuh — BatchNorm finalize as the tail of the kernel that produced the statistic slots (gaddpg_bn_tail).
//
// Every statistics producer (row-GEMM epilogues, the SA1 first layer, the pool / mask backward kernels) leaves per-CTA
// partial sums in `stats` ([slot][2][C] floats).  Instead of a separate bn_finalize launch per BatchNorm layer and pass
// (77 launches of ~6 us + their launch gaps per update step), the CTA that finishes LAST sums the slots and writes the
// per-channel constants.  Slots are summed in slot order in FP64 whoever is last, so the result does not depend on
// scheduling (deterministic), and the arithmetic is the one of bn_finalize_fwd_kernel / bn_finalize_bwd_kernel
// (gemm_rows.cu): train-mode torch.nn.BatchNorm2d / BatchNorm1d batch statistics, running-stat update with the unbiased
// variance, and the BatchNorm-backward sums m1 / m2 / dgamma / dbeta.
#pragma once
#include "common.cuh"
#include "gemm_rows.cuh"

typedef gaddpg_bn_tail BNTail;

constexpr int BNT_SCRATCH_DOUBLES = 2048;  // minimum shared scratch of the tail: 2C doubles for C <= 1024 (16 KB; kernels lend memory
                                           // that is free by then; more scratch = more slot groups in flight)

// Ticket: true in exactly one CTA of the launch — the one whose ticket shows every other CTA has published its slot.
// Call with all threads of the CTA, after the CTA's own slot stores.
__device__ __forceinline__ bool bnt_last_cta(unsigned int* counter, unsigned int total_ctas) {
  __shared__ unsigned int s_last;
  __syncthreads();  // the CTA's slot stores are ordered before thread 0's fence (barrier + cumulative fence, as in a grid barrier)
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(counter, 1u);
    s_last = (t + 1u == total_ctas) ? 1u : 0u;
    if (s_last) {
      *counter = 0u;  // everybody has arrived: ready for the next launch on this statistics buffer
      __threadfence();
    }
  }
  __syncthreads();
  return s_last != 0u;
}

// Sum slots [0, nslots) of stats ([slot][2][C] floats, C % 4 == 0, C <= 1024) and finalize channels [0, C).  All threads of the
// (last) CTA call it; blockDim.x >= 256.  ONE pass: a thread owns float4 quads of the slot row (four consecutive channels of the
// sums half or of the second-statistic half) and, when there are fewer quads than threads, one of G interleaved slot groups; the
// per-channel constants (gamma, beta, running statistics) are fetched before the slot loads so that their latency overlaps; one
// barrier; then a thread per channel folds the G groups in group order (fixed summation order) and writes the results.
// U = slot loads a thread issues per round (all independent, predicated): the tcgen05 kernels have the registers for one
// round over all their slots; the small streaming kernels keep U low so that the tail does not cost them occupancy.
template <int U>
__device__ __forceinline__ void bnt_finalize(const BNTail& t, const float* __restrict__ stats, int C, int nslots, double* scratch,
                                             int scratch_doubles) {
  const int T = (int)blockDim.x, tid = (int)threadIdx.x;
  const int nq = C >> 1;                              // float4 per slot row (2C floats)
  // per-channel constants of the channels this thread finalizes (c = tid + k T), issued first
  constexpr int MAXCH = 4;                            // C <= 4 * blockDim.x
  float ca[MAXCH], cb[MAXCH], crm[MAXCH], crv[MAXCH], cdg[MAXCH], cdb[MAXCH];
  const bool upd_running = t.kind == 1 && t.running_mean && t.momentum != 1.f;
  const bool acc_grads = t.kind == 2 && t.accumulate;
#pragma unroll
  for (int k = 0; k < MAXCH; ++k) {
    const int c = tid + k * T;
    ca[k] = cb[k] = crm[k] = crv[k] = cdg[k] = cdb[k] = 0.f;
    if (c < C) {
      ca[k] = t.a[c];
      cb[k] = t.b[c];
      if (upd_running) {
        crm[k] = t.running_mean[c];
        crv[k] = t.running_var[c];
      }
      if (acc_grads) {
        if (t.dgamma) cdg[k] = t.dgamma[c];
        if (t.dbeta) cdb[k] = t.dbeta[c];
      }
    }
  }
  int G = nq <= T ? T / nq : 1;
  if (G * 2 * C > scratch_doubles) G = scratch_doubles / (2 * C);
  G = G > nslots ? nslots : G;
  G = G < 1 ? 1 : G;
  const float4* base = reinterpret_cast<const float4*>(stats);
  for (int q = (nq <= T ? tid % nq : tid), grp = (nq <= T ? tid / nq : 0); q < nq && grp < G; q += T) {
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    const float4* src = base + q;
    for (int s = grp; s < nslots; s += U * G) {       // one round: U independent 16-byte loads in flight per thread
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u)
        v[u] = (s + u * G < nslots) ? __ldcg(src + (long long)(s + u * G) * nq) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        a0 += (double)v[u].x;
        a1 += (double)v[u].y;
        a2 += (double)v[u].z;
        a3 += (double)v[u].w;
      }
    }
    double* d = scratch + (long long)grp * 2 * C + 4 * q;   // column 4q.. of the 2C-wide row: [0, C) sums, [C, 2C) second statistic
    d[0] = a0;
    d[1] = a1;
    d[2] = a2;
    d[3] = a3;
    if (nq <= T) break;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < MAXCH; ++k) {
    const int c = tid + k * T;
    if (c >= C) continue;
    double s1 = 0.0, q = 0.0;
    for (int g = 0; g < G; ++g) {
      s1 += scratch[(long long)g * 2 * C + c];
      q += scratch[(long long)g * 2 * C + C + c];
    }
    if (t.kind == 1) {
      const double m = s1 / t.count;
      double v = q / t.count - m * m;
      if (v < 0.0) v = 0.0;
      const float mean = (float)m, var = (float)v;
      if (t.running_mean) {
        const double unbiased = t.count > 1.0 ? v * (t.count / (t.count - 1.0)) : v;
        // momentum == 1: staging mode (deferred update, gaddpg_bn_running_update) — store the batch values verbatim
        t.running_mean[c] = (t.momentum == 1.f) ? mean : (1.f - t.momentum) * crm[k] + t.momentum * mean;
        t.running_var[c] = (t.momentum == 1.f) ? (float)unbiased : (1.f - t.momentum) * crv[k] + t.momentum * (float)unbiased;
      }
      const float rstd = 1.0f / sqrtf(var + t.eps);
      const float sc = ca[k] * rstd;
      t.o0[c] = sc;
      t.o1[c] = cb[k] - mean * sc;
      if (t.o2) t.o2[c] = mean;
      if (t.o3) t.o3[c] = rstd;
    } else {
      t.o1[c] = (float)(s1 / t.count);
      t.o2[c] = (float)(q / t.count);
      t.o0[c] = ca[k] * cb[k];
      if (t.dgamma) t.dgamma[c] = cdg[k] + (float)q;
      if (t.dbeta) t.dbeta[c] = cdb[k] + (float)s1;
    }
  }
  if (t.kind == 1 && t.num_batches_tracked && tid == 0) *t.num_batches_tracked += 1;
}

// The whole tail: ticket, then (in the last CTA only) the finalize.  `scratch`: >= BNT_SCRATCH_DOUBLES doubles of shared memory
// nobody else uses any more (more scratch = more slot groups in flight).  Returns after a block-wide barrier in every CTA.
template <int U = 8>
__device__ __forceinline__ void bnt_run(const BNTail& t, const float* stats, int C, int nslots, unsigned int total_ctas, double* scratch,
                                        int scratch_doubles = BNT_SCRATCH_DOUBLES) {
  if (bnt_last_cta(t.counter, total_ctas)) bnt_finalize<U>(t, stats, C, nslots, scratch, scratch_doubles);
}

static inline BNTail tail_or_none(const BNTail* t) {
  BNTail z = {};
  return t ? *t : z;
}
// host side: can this tail run fused (else the caller launches the separate finalize kernel after the producer)
static inline bool bnt_fusable(const BNTail& t, int C) { return t.kind != 0 && t.counter != nullptr && (C % 4) == 0 && C >= 4 && C <= 1024; }
int gaddpg_bn_tail_separate(const BNTail& t, const float* stats, int C, void* stream);  // gemm_rows.cu
