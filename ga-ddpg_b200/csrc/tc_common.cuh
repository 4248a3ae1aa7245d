// tc_common.cuh — PTX wrappers shared by the tcgen05 kernels (tc_gemm.cu, tc_gemm_kc.cu): mbarriers, TMEM
// allocation, tcgen05.mma/commit/ld, the SWIZZLE_128B shared-memory descriptors and the 3xTF32 hi/lo split.
#pragma once
#include "common.cuh"
#include "gemm_rows.cuh"
#ifndef TCP_ADD
#define TCP_ADD(slot, v)
#define TCP_T() 0ll
#endif

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
// the same wait for a role that has nothing to do for a long time (the TN epilogue waits for the whole row loop): back off between
// polls so that the spinning warps do not take issue slots from the producers that share their schedulers
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, unsigned ns) {
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(bar);
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) break;
    __nanosleep(ns);
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, kind::tf32, one thread issues for the CTA
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
        "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]),
        "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]),
        "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// the same load without the wait (software-pipelined TMEM reads): pair with tmem_wait_ld()
__device__ __forceinline__ void tmem_ld32_async(uint32_t taddr, float* r) {
  uint32_t* u = reinterpret_cast<uint32_t*>(r);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
        "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]),
        "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]),
        "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 |
// SBO>>4 <<32 | version 1 <<46 | layout SWIZZLE_128B(2) <<61.  SBO = 1024 B (8 rows x 128 B); LBO unused (=1).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// byte offset of the 16-byte chunk holding (row r, k..k+3) inside an operand stored as K/32 blocks of [rows x 128 B]
__device__ __forceinline__ uint32_t sw128_off(int r, int k, int rows) {
  const int kb = k >> 5, chunk = (k & 31) >> 2;
  return (uint32_t)(kb * rows * 128 + (r >> 3) * 1024 + (r & 7) * 128 + ((chunk ^ (r & 7)) << 4));
}
__device__ __forceinline__ void split_store(unsigned char* hi_base, unsigned char* lo_base, uint32_t off, float4 v) {
  float4 h, l;
  h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
  h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
  h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
  h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
  l.x = v.x - h.x;
  l.y = v.y - h.y;
  l.z = v.z - h.z;
  l.w = v.w - h.w;
  *reinterpret_cast<float4*>(hi_base + off) = h;
  *reinterpret_cast<float4*>(lo_base + off) = l;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// pull a line towards L2 ahead of the register-staged load that will need it (no functional effect)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

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
    float4 g = ldg4(d.c0 + col), m1 = ldg4(d.c1 + col), m2 = ldg4(d.c2 + col), mu = ldg4(d.c3 + col), rs = ldg4(d.c4 + col);
    float w = d.rw ? d.rw[row] : 1.f;
    r.x = g.x * (x.x - w * (m1.x + (y.x - mu.x) * rs.x * m2.x));
    r.y = g.y * (x.y - w * (m1.y + (y.y - mu.y) * rs.y * m2.y));
    r.z = g.z * (x.z - w * (m1.z + (y.z - mu.z) * rs.z * m2.z));
    r.w = g.w * (x.w - w * (m1.w + (y.w - mu.w) * rs.w * m2.w));
  }
  return r;
}


// ---------------------------------------------------------------------------------------------------------
// Shared NT epilogue: one 32-row x 32-column accumulator block of one epilogue warp.
//   tcgen05.ld (lane = row, 32 columns in registers) -> 8 x STS.128 into a [32][36] staging tile -> 8 x LDS.128 in
//   the store layout (lane -> row 4i + (lane>>3), columns 4*(lane&7)..+3) -> 8 x STG.128, each covering four complete
//   128-byte row segments.  Everything the block needs from global memory (mask source, row weights) is loaded
//   before the TMEM wait; the body is branch-free so the loads, shared-memory traffic and stores of a block overlap.
// BatchNorm statistics are kept per lane for its 4 columns (s0/s1) across all tiles of the CTA and reduced once at
// the end of the kernel (epi_reduce_stats), always in the same order.
// ---------------------------------------------------------------------------------------------------------
constexpr int EPI_LD = 36;  // staging row pitch in floats: 16-byte aligned rows, conflict-free STS.128 / LDS.128

// Fused max-pool, first half (gaddpg_nt_problem.pool_keys): lane = column.  The pooled activation of a column is
// relu(bn(max y)) when the BatchNorm scale gamma*rstd is non-negative and relu(bn(min y)) otherwise, and the sign of the scale is
// the sign of gamma (rstd > 0), which is known before the layer runs: each column tracks ONE extreme.  The 32 staged rows
// are loaded up front (independent LDS), then every distinct segment of the block (rows of a segment are contiguous; a block
// usually holds one or two) is reduced branch-free and flushed with one 64-bit atomicMax per column.  Key = order-preserving
// float bits (complemented for a negative gamma) in the high word, ~row in the low word: the maximum key is the extreme value
// at the lowest row that attains it, independent of the order in which CTAs arrive.
__device__ __forceinline__ uint32_t ordered_bits(float y) {
  const uint32_t u = __float_as_uint(y);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ void epi_pool32(const NTProblem& p, const float* stage, int lane, int row_base, int col0, int M, int N) {
  const int col = col0 + lane;
  const bool cval = col < N;
  const int myrow = row_base + lane;
  const int segv = myrow < M ? p.pool_seg[myrow] : -1;
  const bool neg = cval && p.pool_gamma[col] < 0.f;
  const float bias = (p.bias && cval) ? p.bias[col] : 0.f;
  uint32_t u[32];
#pragma unroll
  for (int rr = 0; rr < 32; ++rr) {
    const uint32_t o = ordered_bits(stage[rr * EPI_LD + lane] + bias);
    u[rr] = neg ? ~o : o;
  }
  int cur = __shfl_sync(0xffffffffu, segv, 0);
  while (cur >= 0) {  // warp-uniform: one pass per distinct segment of the block (segment ids increase with the row)
    uint32_t best = 0u;
    int arow = 0, nxt = -1;
#pragma unroll
    for (int rr = 0; rr < 32; ++rr) {
      const int sg = __shfl_sync(0xffffffffu, segv, rr);
      const bool better = (sg == cur) && (u[rr] > best);
      best = better ? u[rr] : best;
      arow = better ? rr : arow;
      nxt = (nxt < 0 && sg > cur) ? sg : nxt;
    }
    if (cval)
      atomicMax(p.pool_keys + (long long)cur * N + col, ((unsigned long long)best << 32) | (0xFFFFFFFFu - (uint32_t)(row_base + arow)));
    cur = nxt;
  }
}

template <int EMODE>
__device__ __forceinline__ void epi_block32(const NTProblem& p, uint32_t taddr, float* stage, int lane, int row_base, int col0,
                                            int M, int N, const float (&wr)[8], bool do_stats, float (&s0)[4], float (&s1)[4]) {
  const int rsub = lane >> 3, c4 = (lane & 7) << 2;
  const int col = col0 + c4;
  const bool cval = col < N;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 yp[8];
  if (EMODE == EPI_DMASK) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = row_base + 4 * i + rsub;
      yp[i] = (cval && row < M) ? ldg4(p.Yprev + (long long)row * p.ldyp + col) : zero4;
    }
  }
  float4 k0 = zero4, k1 = zero4, k2 = zero4, k3 = zero4;  // STORE: bias | DMASK: scale, shift, mean, rstd of the mask source
  bool has_psc = false;
  if (cval) {
    if (EMODE == EPI_STORE) {
      if (p.bias) k0 = make_float4(p.bias[col], p.bias[col + 1], p.bias[col + 2], p.bias[col + 3]);
    } else {
      has_psc = p.psc != nullptr;
      if (has_psc) {
        k0 = ldg4(p.psc + col);
        k1 = ldg4(p.psh + col);
      }
      if (do_stats) {
        k2 = ldg4(p.pmean + col);
        k3 = ldg4(p.prstd + col);
      }
    }
  }
  float r[32];
  const long long tq0 = TCP_T();
  tmem_ld32(taddr, r);
  const long long tq1 = TCP_T();
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stage + lane * EPI_LD + 4 * j) = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
  __syncwarp();
  if (EMODE == EPI_STORE && p.pool_keys != nullptr) epi_pool32(p, stage, lane, row_base, col0, M, N);
  float4 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const float4*>(stage + (4 * i + rsub) * EPI_LD + c4);
  __syncwarp();
  const long long tq2 = TCP_T();
  if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) == 9) { TCP_ADD(8, tq1 - tq0); TCP_ADD(9, tq2 - tq1); }
  const bool relu = (EMODE == EPI_STORE) && p.relu;
  // Fast path (warp-uniform): every row and column of the block is valid and, for the store epilogue, there is no bias /
  // ReLU (the shared-MLP convs have neither) — no per-row predicates, one pointer + constant stride for the stores.  The
  // epilogue warps are the busiest role of the SA1 forward kernels (ncu: never waiting on the accumulator), so the
  // instructions dropped here are kernel time.
  const bool fast = (row_base + 32 <= M) && (col0 + 32 <= N) && (EMODE == EPI_DMASK || (!relu && p.bias == nullptr));
  if (fast) {
    float* cp = p.C + (long long)(row_base + rsub) * p.ldc + col;
    const long long step = 4ll * p.ldc;
    if (EMODE == EPI_STORE) {
      if (!p.no_store) {
#pragma unroll
        for (int i = 0; i < 8; ++i) *reinterpret_cast<float4*>(cp + i * step) = v[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 x = v[i];
        const float w = wr[i];
        s0[0] = fmaf(w, x.x, s0[0]); s1[0] = fmaf(w * x.x, x.x, s1[0]);
        s0[1] = fmaf(w, x.y, s0[1]); s1[1] = fmaf(w * x.y, x.y, s1[1]);
        s0[2] = fmaf(w, x.z, s0[2]); s1[2] = fmaf(w * x.z, x.z, s1[2]);
        s0[3] = fmaf(w, x.w, s0[3]); s1[3] = fmaf(w * x.w, x.w, s1[3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 x = v[i];
        const float4 y = yp[i];
        const float zx = has_psc ? fmaf(y.x, k0.x, k1.x) : y.x, zy = has_psc ? fmaf(y.y, k0.y, k1.y) : y.y;
        const float zz = has_psc ? fmaf(y.z, k0.z, k1.z) : y.z, zw = has_psc ? fmaf(y.w, k0.w, k1.w) : y.w;
        x.x = zx > 0.f ? x.x : 0.f;
        x.y = zy > 0.f ? x.y : 0.f;
        x.z = zz > 0.f ? x.z : 0.f;
        x.w = zw > 0.f ? x.w : 0.f;
        *reinterpret_cast<float4*>(cp + i * step) = x;
        s0[0] += x.x; s1[0] = fmaf(x.x, (y.x - k2.x) * k3.x, s1[0]);
        s0[1] += x.y; s1[1] = fmaf(x.y, (y.y - k2.y) * k3.y, s1[1]);
        s0[2] += x.z; s1[2] = fmaf(x.z, (y.z - k2.z) * k3.z, s1[2]);
        s0[3] += x.w; s1[3] = fmaf(x.w, (y.w - k2.w) * k3.w, s1[3]);
      }
    }
    if ((threadIdx.x & 31) == 0 && (threadIdx.x >> 5) == 9) { TCP_ADD(10, TCP_T() - tq2); TCP_ADD(11, 1); }
    return;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = row_base + 4 * i + rsub;
    const bool ok = cval && row < M;
    float4 x = v[i];
    if (EMODE == EPI_STORE) {
      x.x += k0.x; x.y += k0.y; x.z += k0.z; x.w += k0.w;
      if (relu) {
        x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
      }
      if (ok && !p.no_store) *reinterpret_cast<float4*>(p.C + (long long)row * p.ldc + col) = x;
      if (!ok) x = zero4;  // TMEM columns >= N / rows >= M may hold anything (0 * NaN would poison the sums)
      const float w = ok ? wr[i] : 0.f;
      s0[0] = fmaf(w, x.x, s0[0]); s1[0] = fmaf(w * x.x, x.x, s1[0]);
      s0[1] = fmaf(w, x.y, s0[1]); s1[1] = fmaf(w * x.y, x.y, s1[1]);
      s0[2] = fmaf(w, x.z, s0[2]); s1[2] = fmaf(w * x.z, x.z, s1[2]);
      s0[3] = fmaf(w, x.w, s0[3]); s1[3] = fmaf(w * x.w, x.w, s1[3]);
    } else {
      const float4 y = yp[i];
      const float zx = has_psc ? fmaf(y.x, k0.x, k1.x) : y.x, zy = has_psc ? fmaf(y.y, k0.y, k1.y) : y.y;
      const float zz = has_psc ? fmaf(y.z, k0.z, k1.z) : y.z, zw = has_psc ? fmaf(y.w, k0.w, k1.w) : y.w;
      x.x = (ok && zx > 0.f) ? x.x : 0.f;
      x.y = (ok && zy > 0.f) ? x.y : 0.f;
      x.z = (ok && zz > 0.f) ? x.z : 0.f;
      x.w = (ok && zw > 0.f) ? x.w : 0.f;
      if (ok) *reinterpret_cast<float4*>(p.C + (long long)row * p.ldc + col) = x;
      s0[0] += x.x; s1[0] = fmaf(x.x, (y.x - k2.x) * k3.x, s1[0]);
      s0[1] += x.y; s1[1] = fmaf(x.y, (y.y - k2.y) * k3.y, s1[1]);
      s0[2] += x.z; s1[2] = fmaf(x.z, (y.z - k2.z) * k3.z, s1[2]);
      s0[3] += x.w; s1[3] = fmaf(x.w, (y.w - k2.w) * k3.w, s1[3]);
    }
  }
}

// fold the four row-phase lanes (lane>>3) of each column quad; afterwards every lane holds the warp's column sums
__device__ __forceinline__ void epi_reduce_stats(float (&s)[4]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    s[k] += __shfl_xor_sync(0xffffffffu, s[k], 8);
    s[k] += __shfl_xor_sync(0xffffffffu, s[k], 16);
  }
}

// the NT epilogue needs 16-byte aligned rows of C (and of the mask source): everything else stays on gemm_rows.cu
static inline bool tc_epilogue_ok(const NTProblem& p, int emode) {
  if (p.N % 4 != 0 || p.ldc % 4 != 0 || ((uintptr_t)p.C & 15u) != 0) return false;
  if (p.pool_keys && (emode != EPI_STORE || !p.pool_seg || !p.pool_gamma || p.relu)) return false;
  if (emode == EPI_DMASK) {
    if (p.ldyp % 4 != 0 || ((uintptr_t)p.Yprev & 15u) != 0) return false;
    if (p.psc && ((((uintptr_t)p.psc) | ((uintptr_t)p.psh)) & 15u) != 0) return false;
    if (p.stats && ((((uintptr_t)p.pmean) | ((uintptr_t)p.prstd)) & 15u) != 0) return false;
  }
  return true;
}

}  // namespace
