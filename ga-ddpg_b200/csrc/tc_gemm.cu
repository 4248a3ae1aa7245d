// tc_gemm.cu — tcgen05 (5th-gen tensor core) row-GEMM for the wide set-abstraction layers (sm_100a only).
//
//   C[M,N] = epi( pro(A)[M,K] . W[N,K]^T )     same operand prologues / epilogues as gemm_rows.cu
//
// used for the SA1 shared-MLP layers (1x1 convs 64->64, 64->128 and their dX), which hold ~80 % of the step's rows
// (~420 k compact rows at B=256): reached from /root/reference/core/networks.py:66-71 through upstream
// build_shared_mlp.  FP32 parity (1e-4 through nine BN layers) is kept with a 3xTF32 split: x = hi + lo with
// hi = tf32(x), lo = x - hi (exact), and  a.b ~= lo_a.hi_b + hi_a.lo_b + hi_a.hi_b  accumulated in FP32 in TMEM
// (dropped term lo.lo ~ 2^-22 relative).
//
// Warp-specialised persistent kernel, one CTA per SM, 128-row tiles:
//   warps 0-7  producers : coalesced float4 loads of the A tile, BN(+ReLU) / BN-backward prologue in registers,
//                          hi/lo split, st.shared into the canonical K-major SWIZZLE_128B UMMA layout
//                          (the weights are staged the same way once per CTA), fence.proxy.async, mbarrier arrive
//   warp  8    MMA issuer: one thread issues 3*K/8 tcgen05.mma.kind::tf32 (M=128, N=N, K=8) per tile into a
//                          double-buffered TMEM accumulator and tcgen05.commit's the smem stage / accumulator barriers
//   warps 9-12 epilogue  : tcgen05.ld 32x32b -> registers -> smem re-layout -> 16-byte global stores (four full 128-byte
//                          row segments per instruction), mask reads and per-column BatchNorm statistics (fixed order,
//                          one slot per CTA); shared with tc_gemm_kc.cu (tc_common.cuh: epi_block32)
// TMA: the A operand passes through a per-element prologue (BN+ReLU / BN-backward) before it may reach the tensor core, so the
// producer warps are its copy engine; the WEIGHTS arrive by cp.async.bulk.tensor loads when the caller provides their
// pre-split hi / lo images (gaddpg_nt_problem.Bw_hi / Bw_lo: the SA1 forward layers), and full output blocks of the plain
// store epilogue leave by cp.async.bulk.tensor stores (epi_store_tma32).
#ifdef TC_PROFILE
__device__ unsigned long long g_tc_prof[148 * 16];
#define TCP_ADD(slot, v) atomicAdd(&g_tc_prof[blockIdx.x * 16 + (slot)], (unsigned long long)(v))
#define TCP_T() clock64()
#else
#define TCP_ADD(slot, v)
#define TCP_T() 0ll
#endif

#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"
#include "tma.cuh"
#include "impl.h"
#include "bn_tail.cuh"

namespace {

constexpr int TC_BM = 128;
// 8 producer warps + 1 MMA warp + 4 epilogue warps.  Two producer warps per scheduler: one warp's stream (prologue, hi/lo
// split, swizzled STS.128) runs at ~0.2 IPC, which bounded the tile rate before HBM did (see tc_gemm_kc.cu).
constexpr int TC_PW = 8;
constexpr int TC_PT = TC_PW * 32;
constexpr int TC_THREADS = (TC_PW + 1 + 4) * 32;
#ifndef TC_U_FWD
#define TC_U_FWD 8
#endif
#ifndef TC_U_BWD
#define TC_U_BWD 4
#endif
#ifndef TC_PFD
#define TC_PFD 1   // producer units prefetched into L2 ahead of their register-staged loads (0: off).  Measured (graph-timed,
                   // M = 423 k): dX 64->64 85.9 -> 75.3 us, dX 128->64 142.9 -> 122.7 us, fwd 64->128 78.6 -> 72.5 us (1, 2 and 3
                   // units ahead within 2 %); fwd 64->64 47.4 -> 48.2 us: store-bound, left without prefetch
#endif

constexpr int TC_KS = 64;  // K width of one shared-memory A stage (a K = 128 tile is two stages: its second half loads
                           // while the tensor core works on the first)
struct TcSmemLayout {
  uint32_t w_hi, w_lo, a_hi[2], a_lo[2], stage_buf, cc, bars, total;
};
__host__ __device__ inline TcSmemLayout tc_layout(int N, int K) {
  TcSmemLayout L;
  uint32_t off = 0;
  L.w_hi = off; off += (uint32_t)N * K * 4;
  L.w_lo = off; off += (uint32_t)N * K * 4;
  for (int s = 0; s < 2; ++s) {
    L.a_hi[s] = off; off += (uint32_t)TC_BM * TC_KS * 4;
    L.a_lo[s] = off; off += (uint32_t)TC_BM * TC_KS * 4;
  }
  L.stage_buf = off; off += 4 * 5120;              // per-epilogue-warp transpose buffers (1024-byte aligned: TMA store boxes)
  L.cc = off; off += 5 * 128 * 4;                // per-column prologue constants c0..c4 [5][K <= 128]
  L.bars = off; off += 256;                     // mbarriers + tmem address + stats scratch header
  off += 2 * 4 * 256 * 4;                       // cross-warp stats combine: [2][4 warps][N<=256]
  L.total = off;
  return L;
}

// per-column prologue constants: staged once per CTA in shared memory ([5][128] floats), fetched per pipeline unit
template <int MODE>
struct ColConsts {
  float4 c0, c1, c2, c3, c4;
  __device__ __forceinline__ void load(const float* tab, int col) {
    if (MODE == OP_BNRELU) {
      c0 = *reinterpret_cast<const float4*>(tab + col);
      c1 = *reinterpret_cast<const float4*>(tab + 128 + col);
    } else if (MODE == OP_BNBWD || MODE == OP_BNBWD_POOL) {
      c0 = *reinterpret_cast<const float4*>(tab + col);
      c1 = *reinterpret_cast<const float4*>(tab + 128 + col);
      c2 = *reinterpret_cast<const float4*>(tab + 256 + col);
      c3 = *reinterpret_cast<const float4*>(tab + 384 + col);
      c4 = *reinterpret_cast<const float4*>(tab + 512 + col);
    }
  }
  __device__ __forceinline__ float4 apply(float4 x, float4 y, float w) const {
    float4 r;
    if (MODE == OP_PLAIN) {
      r = x;
    } else if (MODE == OP_BNRELU) {
      r.x = fmaxf(fmaf(x.x, c0.x, c1.x), 0.f);
      r.y = fmaxf(fmaf(x.y, c0.y, c1.y), 0.f);
      r.z = fmaxf(fmaf(x.z, c0.z, c1.z), 0.f);
      r.w = fmaxf(fmaf(x.w, c0.w, c1.w), 0.f);
    } else {
      r.x = c0.x * (x.x - w * (c1.x + (y.x - c3.x) * c4.x * c2.x));
      r.y = c0.y * (x.y - w * (c1.y + (y.y - c3.y) * c4.y * c2.y));
      r.z = c0.z * (x.z - w * (c1.z + (y.z - c3.z) * c4.z * c2.z));
      r.w = c0.w * (x.w - w * (c1.w + (y.w - c3.w) * c4.w * c2.w));
    }
    return r;
  }
};

// One full 32-row x 32-column block of a plain store epilogue through the TMA.  Staging tile = [32 rows x 128 B] with the
// SWIZZLE_128B pattern (16-byte chunk j of row r at chunk j ^ (r & 7)): conflict-free for the row-wise STS.128 of the TMEM
// layout (lane = row) and for the LDS.128 of the statistics layout (lane -> rows 4i + (lane >> 3), columns 4 (lane & 7) ..).
__device__ __forceinline__ void epi_store_tma32(uint32_t taddr, unsigned char* stage, int lane, int row_base, int col0, const CUtensorMap* tmC,
                                                const float (&wr)[8], bool do_stats, float (&s0)[4], float (&s1)[4]) {
  float r[32];
  tmem_ld32(taddr, r);
  if (lane == 0) tma_wait_read0();   // the previous block's store has read the staging tile
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4*>(stage + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    tma_store_2d(tmC, smem_u32(stage), col0, row_base);
    tma_commit();
  }
  if (do_stats) {
    const int rsub = lane >> 3, ch = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = 4 * i + rsub;
      const float4 x = *reinterpret_cast<const float4*>(stage + row * 128 + ((ch ^ (row & 7)) << 4));
      const float w = wr[i];
      s0[0] = fmaf(w, x.x, s0[0]); s1[0] = fmaf(w * x.x, x.x, s1[0]);
      s0[1] = fmaf(w, x.y, s0[1]); s1[1] = fmaf(w * x.y, x.y, s1[1]);
      s0[2] = fmaf(w, x.z, s0[2]); s1[2] = fmaf(w * x.z, x.z, s1[2]);
      s0[3] = fmaf(w, x.w, s0[3]); s1[3] = fmaf(w * x.w, x.w, s1[3]);
    }
  }
}

// NCB = 32-column accumulator blocks the epilogue keeps statistics for: 4 (N <= 128, the SA1 layers) or 8 (N <= 256)
template <int K, int AMODE, int EMODE, int NCB>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_nt_kernel(const NTProblem p, const __grid_constant__ CUtensorMap tmC,
                                                                   const int tma_store, const __grid_constant__ CUtensorMap tmWh,
                                                                   const __grid_constant__ CUtensorMap tmWl, const int tma_w) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // SWIZZLE_128B operands need 1024-byte aligned bases (the host adds 1024 bytes of slack)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int N = p.N;
  const TcSmemLayout L = tc_layout(N, K);
  constexpr int NH = K / TC_KS;  // A stages per tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* a_full = bars;           // [2]
  uint64_t* a_empty = bars + 2;      // [2]
  uint64_t* acc_full = bars + 4;     // [2]
  uint64_t* acc_empty = bars + 6;    // [2]
  uint64_t* w_full = bars + 8;       // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  float* stat_comb = reinterpret_cast<float*>(smem + L.bars + 256);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int M = p.M_dev ? *p.M_dev : p.M_max;
  M = M < p.M_max ? M : p.M_max;
  const int ntiles = (M + TC_BM - 1) / TC_BM;
  const uint32_t tmem_cols = (2 * N <= 32) ? 32 : (2 * N <= 64) ? 64 : (2 * N <= 128) ? 128 : (2 * N <= 256) ? 256 : 512;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a_full[s], TC_PT);
      mbar_init(&a_empty[s], 1);
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);
    }
    mbar_init(w_full, tma_w ? 1 : TC_PT);   // weights: one TMA transaction barrier, or one arrival per producer thread
    fence_barrier_init();
  }
  if (warp == TC_PW) tmem_alloc(tmem_slot, tmem_cols);
  {
    float* cct = reinterpret_cast<float*>(smem + L.cc);
    constexpr int NC = (AMODE == OP_PLAIN) ? 0 : (AMODE == OP_BNRELU) ? 2 : 5;
    for (int i = tid; i < NC * K; i += TC_THREADS) {
      const int a = i / K, c = i % K;
      const float* sp = a == 0 ? p.A.c0 : a == 1 ? p.A.c1 : a == 2 ? p.A.c2 : a == 3 ? p.A.c3 : p.A.c4;
      cct[a * 128 + c] = sp[c];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < TC_PW) {
    // ===================== producers =====================
    constexpr int KQ4 = K / 4;         // float4 per row of the weight matrix
    constexpr int SQ4 = TC_KS / 4;     // float4 per row of one A stage
    constexpr int RPI = TC_PT / SQ4;   // rows covered by the producer threads per iteration (16)
    constexpr int ITERS = TC_BM / RPI; // row-iterations per stage (8)
    constexpr bool BWD = (AMODE == OP_BNBWD || AMODE == OP_BNBWD_POOL);
    constexpr int U = BWD ? TC_U_BWD : TC_U_FWD;  // row-iterations per pipeline unit (two register sets in flight)
    static_assert(K % TC_KS == 0 && ITERS % U == 0, "K must be 64 or 128");
    if (!tma_w) {   // no pre-split image of this weight matrix: split it here (once per CTA)
      for (int idx = tid; idx < N * KQ4; idx += TC_PT) {
        int n = idx / KQ4, k = (idx % KQ4) << 2;
        float4 v = ldg4(p.Bw + (long long)n * p.ldb + k);
        split_store(smem + L.w_hi, smem + L.w_lo, sw128_off(n, k, N), v);
      }
      fence_proxy_async();
      mbar_arrive(w_full);
    }
    const int kc = (tid % SQ4) << 2, rsub = tid / SQ4;  // column inside the 64-wide stage, first row served
    const float* cct = reinterpret_cast<const float*>(smem + L.cc);
    const bool pf_on = BWD || N > 64;   // (uniform) the 64 -> 64 forward layer is bound by its stores, not by its loads
    // Software pipeline over "units" of U row-iterations: the global loads of unit u+1 are issued into a second
    // register set before unit u is transformed and stored, so HBM latency overlaps the prologue math and the
    // shared-memory stores (and runs ahead across stage and tile boundaries, independent of the smem-stage barriers).
    constexpr int UPS = ITERS / U;    // units per stage
    constexpr int UPT = NH * UPS;     // units per tile
    struct Regs {
      float4 x[U], y[U];
      float w[U];
      uint32_t m[U];  // BNBWD_POOL: arg-max bits of this thread's 4 columns (after the shift in process())
    };
    const int my_tiles = (ntiles > (int)blockIdx.x) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int total_units = my_tiles * UPT;
    int segn[U];  // BNBWD_POOL: segments of the rows of the NEXT unit to be issued
    auto prefetch_seg = [&](int u) {
      const int it = u / UPT, i0 = ((u % UPT) % UPS) * U;
      const int row0 = ((int)blockIdx.x + it * (int)gridDim.x) * TC_BM;
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int row = row0 + rsub + (i0 + k) * RPI;
        segn[k] = (u < total_units && row < M) ? p.A.pseg[row] : 0;
      }
    };
    auto issue = [&](Regs& R, int u) {
      const int it = u / UPT, v = u % UPT, h = v / UPS, i0 = (v % UPS) * U;
      const int row0 = ((int)blockIdx.x + it * (int)gridDim.x) * TC_BM;
      const int kcol = h * TC_KS + kc;
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int row = row0 + rsub + (i0 + k) * RPI;
        R.x[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        R.y[k] = R.x[k];
        R.w[k] = 1.f;
        if (row < M) {
          if (AMODE != OP_BNBWD_POOL) R.x[k] = ldg4(p.A.X + (long long)row * p.A.ldx + kcol);
          if (BWD) {
            R.y[k] = ldg4(p.A.Y + (long long)row * p.A.ldy + kcol);
            if (p.A.rw) R.w[k] = p.A.rw[row];
          }
        }
      }
      if (AMODE == OP_BNBWD_POOL) {
        // D is the max-pool gradient: non-zero only where this row is the arg-max of its (segment, channel).  It is
        // rebuilt from the (S, C) masked-gradient table E (L2 resident) and a 16-byte arg-max bit mask per row instead
        // of being read from a dense (M, C) tensor in HBM.  The row -> segment lookup was issued one unit ahead (segn),
        // so nothing here waits on a dependent load; the select happens in process().
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const int row = row0 + rsub + (i0 + k) * RPI;
          R.m[k] = 0u;
          if (row < M) {
            R.x[k] = ldg4(p.A.X + (long long)segn[k] * p.A.ldx + kcol);
            R.m[k] = p.A.pmask[(long long)row * (K / 32) + (kcol >> 5)];
          }
        }
        prefetch_seg(u + 1);
      }
      if (TC_PFD > 0 && pf_on && u + TC_PFD < total_units) {
        // L2 prefetch of the unit TC_PFD ahead: the producers wait on these loads (long-scoreboard stalls dominate their issue
        // slots, ncu); with the line already in L2 the register-staged load of that unit returns in a fraction of the time.  A
        // unit is U*16 consecutive rows x 64 columns (two 128-byte lines per row and operand): one line per producer thread.
        const int up = u + TC_PFD;
        const int it2 = up / UPT, v2 = up % UPT;
        constexpr int LPO = U * 32;                               // lines per operand and unit
        const int li = tid % LPO, op = tid / LPO;                 // op 0: X, op 1: Y (BatchNorm-backward forms)
        const int row = ((int)blockIdx.x + it2 * (int)gridDim.x) * TC_BM + (v2 % UPS) * U * RPI + (li >> 1);
        const int col = (v2 / UPS) * TC_KS + (li & 1) * 32;
        if (row < M) {
          if (op == 0 && AMODE != OP_BNBWD_POOL) prefetch_l2(p.A.X + (long long)row * p.A.ldx + col);
          if (op == 1 && BWD) prefetch_l2(p.A.Y + (long long)row * p.A.ldy + col);
        }
      }
    };
    auto process = [&](const Regs& R, int u) {
      const int it = u / UPT, v = u % UPT, h = v / UPS, i0 = (v % UPS) * U;
      const int g = it * NH + h, s = g & 1;  // stage counter of this CTA, ring slot
      const int row0 = ((int)blockIdx.x + it * (int)gridDim.x) * TC_BM;
      const int kcol = h * TC_KS + kc;
      ColConsts<AMODE> cc;
      cc.load(cct, kcol);
      if (i0 == 0) {
        const long long t0 = TCP_T();
        mbar_wait(&a_empty[s], ((uint32_t)(g >> 1) & 1u) ^ 1u);
        if (tid == 0) TCP_ADD(0, TCP_T() - t0);
      }
      unsigned char* ah = smem + L.a_hi[s];
      unsigned char* al = smem + L.a_lo[s];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int r = rsub + (i0 + k) * RPI;
        float4 x = R.x[k];
        if (AMODE == OP_BNBWD_POOL) {
          const uint32_t bits = R.m[k] >> (kcol & 31);
          x.x = (bits & 1u) ? x.x : 0.f;
          x.y = (bits & 2u) ? x.y : 0.f;
          x.z = (bits & 4u) ? x.z : 0.f;
          x.w = (bits & 8u) ? x.w : 0.f;
        }
        float4 vv = (row0 + r < M) ? cc.apply(x, R.y[k], R.w[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
        split_store(ah, al, sw128_off(r, kc, TC_BM), vv);
      }
      if (i0 + U == ITERS) {
        fence_proxy_async();
        mbar_arrive(&a_full[s]);
      }
    };
    Regs RA, RB;
    const long long tp0 = TCP_T();
    if (AMODE == OP_BNBWD_POOL) prefetch_seg(0);
    if (total_units > 0) issue(RA, 0);
#pragma unroll 1
    for (int u = 0; u < total_units; u += 2) {
      if (u + 1 < total_units) issue(RB, u + 1);
      process(RA, u);
      if (u + 2 < total_units) issue(RA, u + 2);
      if (u + 1 < total_units) process(RB, u + 1);
    }
    if (tid == 0) TCP_ADD(1, TCP_T() - tp0);
  } else if (warp == TC_PW) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 (1<<4), a=b=TF32 (2<<7, 2<<10), K-major A and B,
      // N>>3 at bit 17, M>>4 at bit 24
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      const uint32_t sbase = smem_u32(smem);
      if (tma_w) {   // pre-split hi / lo weight images (refreshed once per optimiser step) arrive by TMA: one [N x 32] box per
                     // 32-column K block lands directly in the K-major SWIZZLE_128B operand layout
        mbar_expect_tx(w_full, (uint32_t)(2 * N * K * 4));
        for (int kb = 0; kb < K / 32; ++kb) {
          tma_load_2d(sbase + L.w_hi + kb * N * 128, &tmWh, kb * 32, 0, w_full);
          tma_load_2d(sbase + L.w_lo + kb * N * 128, &tmWl, kb * 32, 0, w_full);
        }
      }
      mbar_wait(w_full, 0);
      tc_fence_after();
      int it = 0, g = 0;
      const long long tm0 = TCP_T();
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int b = it & 1;
        const uint32_t bph = (uint32_t)(it >> 1) & 1u;
        {
          const long long t0 = TCP_T();
          mbar_wait(&acc_empty[b], bph ^ 1u);
          TCP_ADD(2, TCP_T() - t0);
        }
        const uint32_t d_tmem = tmem_base + (uint32_t)(b * N);
        uint32_t acc = 0;
#pragma unroll
        for (int h = 0; h < NH; ++h, ++g) {
          const int s = g & 1;
          {
            const long long t0 = TCP_T();
            mbar_wait(&a_full[s], (uint32_t)(g >> 1) & 1u);
            TCP_ADD(3, TCP_T() - t0);
          }
          tc_fence_after();
#pragma unroll
          for (int kb = 0; kb < (TC_KS >> 5); ++kb) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t aoff = (uint32_t)(kb * TC_BM * 128 + ks * 32);
              const uint32_t boff = (uint32_t)((h * (TC_KS >> 5) + kb) * N * 128 + ks * 32);
              const uint64_t a_hi = make_desc(sbase + L.a_hi[s] + aoff), a_lo = make_desc(sbase + L.a_lo[s] + aoff);
              const uint64_t b_hi = make_desc(sbase + L.w_hi + boff), b_lo = make_desc(sbase + L.w_lo + boff);
              umma_tf32(d_tmem, a_lo, b_hi, idesc, acc);
              umma_tf32(d_tmem, a_hi, b_lo, idesc, 1u);
              umma_tf32(d_tmem, a_hi, b_hi, idesc, 1u);
              acc = 1u;
            }
          }
          umma_commit(&a_empty[s]);  // this smem stage may be refilled once these MMAs have read it
        }
        umma_commit(&acc_full[b]);   // accumulator ready for the epilogue
      }
      TCP_ADD(4, TCP_T() - tm0);
    }
    __syncwarp();
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    float* stage = reinterpret_cast<float*>(smem + L.stage_buf + (warp - TC_PW - 1) * 5120);
    const bool do_stats = (p.stats != nullptr);
    const int rsub = lane >> 3;
    float s0[NCB][4], s1[NCB][4];
#pragma unroll
    for (int c = 0; c < NCB; ++c)
#pragma unroll
      for (int k = 0; k < 4; ++k) s0[c][k] = s1[c][k] = 0.f;
    int it = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int b = it & 1;
      const uint32_t bph = (uint32_t)(it >> 1) & 1u;
      const int row_base = t * TC_BM + q * 32;
      // duplicate-row multiplicities of the 8 rows this lane stores (statistics weights)
      float wr[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) wr[i] = 1.f;
      if (EMODE == EPI_STORE && do_stats && p.srw) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = row_base + 4 * i + rsub;
          wr[i] = (row < M) ? p.srw[row] : 0.f;
        }
      }
      if (EMODE == EPI_DMASK && TC_PFD > 0) {
        // mask source of this CTA's NEXT tile towards L2 while the accumulator of this one is still being produced (the
        // epilogue's Yprev loads were its longest exposed latency: profiles/r1_tc_roles.md)
        const int rown = (t + (int)gridDim.x) * TC_BM + q * 32 + lane;
        if (rown < M)
          for (int c = 0; c < N; c += 32) prefetch_l2(p.Yprev + (long long)rown * p.ldyp + c);
      }
      {
        const long long t0 = TCP_T();
        mbar_wait(&acc_full[b], bph);
        if (warp == TC_PW + 1 && lane == 0) TCP_ADD(5, TCP_T() - t0);
      }
      tc_fence_after();
      const long long te0 = TCP_T();
      // full 32 x 32 blocks of a plain store epilogue leave through the TMA: tcgen05.ld -> swizzled staging tile -> ONE
      // cp.async.bulk.tensor store per block (the 8 x STG.128 per lane of the generic path were the measured limiter of the
      // forward kernels: profiles/r1_tc_roles.md); statistics are read back from the same staging tile
      const bool tma_blk = (EMODE == EPI_STORE) && tma_store && (row_base + 32 <= M) && !p.relu && p.bias == nullptr && p.pool_keys == nullptr;
#pragma unroll
      for (int cb = 0; cb < NCB; ++cb) {
        if (cb * 32 < N) {
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * N + cb * 32);
          if (EMODE == EPI_STORE && tma_blk && cb * 32 + 32 <= N)
            epi_store_tma32(ta, reinterpret_cast<unsigned char*>(stage), lane, row_base, cb * 32, &tmC, wr, do_stats, s0[cb], s1[cb]);
          else
            epi_block32<EMODE>(p, ta, stage, lane, row_base, cb * 32, M, N, wr, do_stats, s0[cb], s1[cb]);
        }
      }
      if (warp == TC_PW + 1 && lane == 0) { TCP_ADD(6, TCP_T() - te0); TCP_ADD(7, 1); }
      tc_fence_before();
      if (lane == 0) mbar_arrive(&acc_empty[b]);
    }
    if (EMODE == EPI_STORE && tma_store && lane == 0) tma_wait_all0();
    if (do_stats) {
#pragma unroll
      for (int cb = 0; cb < NCB; ++cb) {
        if (cb * 32 < N) {
          epi_reduce_stats(s0[cb]);
          epi_reduce_stats(s1[cb]);
          if (lane < 8) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int col = cb * 32 + lane * 4 + k;
              if (col < N) {
                stat_comb[(warp - TC_PW - 1) * 256 + col] = s0[cb][k];
                stat_comb[1024 + (warp - TC_PW - 1) * 256 + col] = s1[cb][k];
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.stats) {
    for (int c = tid; c < N; c += TC_THREADS) {
      float a0 = stat_comb[c] + stat_comb[256 + c] + stat_comb[512 + c] + stat_comb[768 + c];
      float a1 = stat_comb[1024 + c] + stat_comb[1280 + c] + stat_comb[1536 + c] + stat_comb[1792 + c];
      p.stats[(long long)blockIdx.x * 2 * N + c] = a0;
      p.stats[(long long)blockIdx.x * 2 * N + N + c] = a1;
    }
    if (p.tail.kind == 0)   // (a fused tail reads the gridDim.x live slots only)
      for (int slot = blockIdx.x + gridDim.x; slot < GADDPG_STAT_SLOTS; slot += gridDim.x)
        for (int c = tid; c < 2 * N; c += TC_THREADS) p.stats[(long long)slot * 2 * N + c] = 0.f;
  }
  if (warp == TC_PW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
  // BatchNorm finalize by the last CTA (the operand stages are free: every MMA and store of this CTA has completed)
  if (p.stats && p.tail.kind != 0) bnt_run<13>(p.tail, p.stats, N, ntiles < (int)gridDim.x ? ntiles : (int)gridDim.x, gridDim.x, reinterpret_cast<double*>(smem), 8192);
}

}  // namespace

// Shapes the tensor-core kernel takes; everything else stays on the FP32 FFMA kernel of gemm_rows.cu.
bool gaddpg_tc_gemm_supported(const NTProblem& p, int amode, int emode) {
  if (p.M_max < 8192) return false;  // small problems are latency bound either way
  if (p.K != 64 && p.K != 128) return false;
  if (p.N % 16 != 0 || p.N < 16 || p.N > 256) return false;
  if (!tc_epilogue_ok(p, emode)) return false;
  if (amode == OP_BNBWD_POOL &&
      !(p.K == 128 && p.N <= 128 && emode == EPI_DMASK && p.A.pmask && p.A.pseg && p.A.ldx == 128))
    return false;
  if (p.ldb != p.K) {
    if (p.ldb % 4 != 0) return false;
  }
  if (tc_layout(p.N, p.K).total + 1024 > 227 * 1024) return false;
  (void)amode;
  (void)emode;
  return true;
}

int gaddpg_tc_gemm_nt_impl(const NTProblem* p, int amode, int emode, void* stream) {
  const TcSmemLayout L = tc_layout(p->N, p->K);
  const size_t smem = L.total + 1024;  // slack for the 1024-byte alignment of the operand regions
  int tiles = ceil_div(p->M_max, TC_BM);
  int grid = tiles < gaddpg_sm_count() ? tiles : gaddpg_sm_count();
  cudaStream_t st = (cudaStream_t)stream;
  // C tiles of the plain store epilogue leave through cp.async.bulk.tensor stores (GADDPG_TMA_STORE=0: the STG path, for A/B)
  static const bool tma_env = []() { const char* e = getenv("GADDPG_TMA_STORE"); return !(e && e[0] == '0'); }();
  CUtensorMap tmC;
  memset(&tmC, 0, sizeof(tmC));
  int tma_store = 0;
  if (tma_env && emode == EPI_STORE && !p->no_store && p->C && (p->N % 32) == 0 && (p->ldc % 4) == 0 && ((uintptr_t)p->C & 15u) == 0)
    tma_store = make_map(&tmC, p->C, p->M_max, p->N, p->ldc, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B) ? 1 : 0;
  CUtensorMap tmWh, tmWl;
  memset(&tmWh, 0, sizeof(tmWh));
  memset(&tmWl, 0, sizeof(tmWl));
  int tma_w = 0;
  if (tma_env && p->Bw_hi && p->Bw_lo && p->N <= 256)
    tma_w = (make_map(&tmWh, p->Bw_hi, p->N, p->K, p->K, p->N, 32, CU_TENSOR_MAP_SWIZZLE_128B) &&
             make_map(&tmWl, p->Bw_lo, p->N, p->K, p->K, p->N, 32, CU_TENSOR_MAP_SWIZZLE_128B)) ? 1 : 0;
#define TC_LAUNCH(KK, A, E, NCB_)                                                                              \
  {                                                                                                            \
    auto kern = tc_gemm_nt_kernel<KK, A, E, NCB_>;                                                             \
    GADDPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    kern<<<grid, TC_THREADS, smem, st>>>(*p, tmC, tma_store, tmWh, tmWl, tma_w);                       \
    GADDPG_CHECK_LAUNCH("tc_gemm_nt_kernel");                                                                  \
    return GADDPG_OK;                                                                                          \
  }
#define TC_CASE(A, E)                                                                                          \
  if (amode == A && emode == E) {                                                                              \
    if (p->K == 64 && p->N <= 128) TC_LAUNCH(64, A, E, 4)                                                      \
    if (p->K == 64) TC_LAUNCH(64, A, E, 8)                                                                     \
    if (p->K == 128 && p->N <= 128) TC_LAUNCH(128, A, E, 4)                                                    \
    if (p->K == 128) TC_LAUNCH(128, A, E, 8)                                                                   \
  }
  if (amode == OP_BNBWD_POOL && emode == EPI_DMASK && p->K == 128 && p->N <= 128) TC_LAUNCH(128, OP_BNBWD_POOL, EPI_DMASK, 4)
  TC_CASE(OP_PLAIN, EPI_STORE)
  TC_CASE(OP_BNRELU, EPI_STORE)
  TC_CASE(OP_BNBWD, EPI_DMASK)
  TC_CASE(OP_BNBWD, EPI_STORE)
  TC_CASE(OP_PLAIN, EPI_DMASK)
#undef TC_CASE
#undef TC_LAUNCH
  gaddpg_set_error("tc_gemm_nt: unsupported mode pair (%d,%d)", amode, emode);
  return GADDPG_ERR_UNSUPPORTED;
}

#ifdef TC_PROFILE
extern "C" __attribute__((visibility("default"))) int gaddpg_debug_tc_prof(unsigned long long* host_out, int reset) {
  if (host_out) cudaMemcpyFromSymbol(host_out, g_tc_prof, sizeof(unsigned long long) * 148 * 16);
  if (reset) {
    static unsigned long long z[148 * 16];
    cudaMemcpyToSymbol(g_tc_prof, z, sizeof(z));
  }
  return 0;
}
#endif
