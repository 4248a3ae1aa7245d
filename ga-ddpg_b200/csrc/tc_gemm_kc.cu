// tc_gemm_kc.cu — K-chunked tcgen05 (3xTF32) row-GEMMs for every shape tc_gemm.cu does not take (sm_100a only).
//
//   NT:  C[M,N]  = epi( pro(A)[M,K] . W[N,K]^T )            any K (multiple of 4), any N, M >= 128
//   TN:  dW[N,K] = sum_r pro1(P)[r,:]^T pro2(Q)[r,:]         weight gradients (split over rows, fixed-order reduce)
//
// These carry SA2 / SA3 / the FC-BN head / the actor-critic Linear layers (K = 132 ... 1024) and ALL weight
// gradients, i.e. what upstream build_shared_mlp / nn.Linear do through cuDNN / cuBLAS for
// /root/reference/core/networks.py:65-92,265-300,315-351.  Same operand prologues / epilogues and the same
// deterministic statistics slots as gemm_rows.cu; same 3xTF32 split as tc_gemm.cu (x = hi + lo, three MMAs).
//
// NT kernel: grid (m-tiles (persistent), n-tiles).  288 threads: warps 0-3 stream A[128 x 32] and W[BN x 32] K-chunks
// (prologue in registers, hi/lo split, st.shared into the K-major SWIZZLE_128B UMMA layout) through an mbarrier
// ring; warp 4 issues 12 tcgen05.mma per chunk into a double-buffered TMEM accumulator; warps 5-8 run the
// epilogue (tcgen05.ld -> smem transpose -> coalesced stores, mask reads, BN statistics).
//
// TN kernel: grid (output tiles [128 x BKT], row splits).  The reduction runs over ROWS, so both operands are fed
// to the tensor core MN-major: the natural [rows x channels] staging (rows 128 B apart) is the canonical MN-major
// SWIZZLE_128B_BASE32B layout (4-row K atoms, two per K = 8 instruction); no transpose anywhere.  The accumulator
// (dW tile) stays in TMEM for the whole kernel and is written once as a per-split partial.
#include "tc_common.cuh"
#include "impl.h"
#include "bn_tail.cuh"

namespace {

constexpr int KC_BM = 128;
// 8 producer warps + 1 MMA warp + 4 epilogue warps (NT and TN kernels).  A producer warp's instruction stream (address math, prologue, hi/lo split, two STS.128 per
// float4) has little ILP and runs at ~0.2 IPC, and with one producer warp per scheduler that — not HBM or the tensor
// pipe — set the time per K chunk (ncu: ~1.4-2.6 us per chunk with 4 warps); two warps per scheduler halve it.
constexpr int KC_PW = 8;
constexpr int KC_NT_THREADS = (KC_PW + 1 + 4) * 32;

// byte offset of (row r, float4 index q) inside one [rows x 128 B] SWIZZLE_128B block
__device__ __forceinline__ uint32_t blk_off(int r, int q) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((q ^ (r & 7)) << 4));
}
// MN-major operands of kind::tf32 have exactly one legal shared-memory layout: SWIZZLE_128B_BASE32B (layout type 1;
// cutlass sm100_common.inl: "for mn-major tf32 operands, SW128_32B is the only available smem layout").  Rows of the
// reduction index are 128 B (32 channels) apart, the swizzle XORs the 32-byte unit index (address bits [5,7)) with the
// row index mod 4 (bits [7,9)), the K atom is 4 rows: LBO = byte stride between 32-channel blocks, SBO = byte stride
// between 4-row groups (cute::UMMA::make_umma_desc<Major::MN>: ((8,n),(4,k)) in uint128 units).
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
         (1ull << 61);
}
// byte offset of (row r, float4 index q) inside one [rows x 128 B] SWIZZLE_128B_BASE32B block
__device__ __forceinline__ uint32_t blk_off_mn(int r, int q) {
  return (uint32_t)(r * 128 + ((((q >> 1) ^ (r & 3)) << 5) | ((q & 1) << 4)));
}

template <int MODE>
struct Consts4 {
  float4 c0, c1, c2, c3, c4;
  __device__ __forceinline__ void load(const Operand& d, int col, bool ok) {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    c0 = c1 = c2 = c3 = c4 = z;
    if (!ok) return;
    if (MODE == OP_BNRELU) {
      c0 = ldg4(d.c0 + col);
      c1 = ldg4(d.c1 + col);
    } else if (MODE == OP_BNBWD || MODE == OP_BNBWD_POOL) {
      c0 = ldg4(d.c0 + col);
      c1 = ldg4(d.c1 + col);
      c2 = ldg4(d.c2 + col);
      c3 = ldg4(d.c3 + col);
      c4 = ldg4(d.c4 + col);
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

// =====================================================================================================
// NT, K-chunked
// =====================================================================================================
template <int BN>
struct KcLayout {
  static constexpr int NSTAGE = (BN == 128) ? 3 : 4;
  static constexpr uint32_t A_BYTES = KC_BM * 128;          // one of hi / lo
  static constexpr uint32_t W_BYTES = BN * 128;
  static constexpr uint32_t STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
  static constexpr uint32_t OFF_STAGEBUF = NSTAGE * STAGE_BYTES;
  static constexpr uint32_t OFF_BARS = OFF_STAGEBUF + 4 * 32 * EPI_LD * 4;
  static constexpr uint32_t OFF_COMB = OFF_BARS + 256;
  static constexpr uint32_t TOTAL = OFF_COMB + 2 * 4 * BN * 4;
};

template <int BN, int AMODE, int EMODE>
__global__ void __launch_bounds__(KC_NT_THREADS, 1) tc_nt_kc_kernel(const NTProblem p) {
  using L = KcLayout<BN>;
  constexpr int NSTAGE = L::NSTAGE;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BARS);
  uint64_t* full = bars;                 // [NSTAGE]
  uint64_t* empty = bars + NSTAGE;       // [NSTAGE]
  uint64_t* acc_full = bars + 2 * NSTAGE;      // [2]
  uint64_t* acc_empty = bars + 2 * NSTAGE + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 4);
  float* stat_comb = reinterpret_cast<float*>(smem + L::OFF_COMB);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int M = p.M_dev ? *p.M_dev : p.M_max;
  M = M < p.M_max ? M : p.M_max;
  const int N = p.N, K = p.K;
  const int ntiles = (M + KC_BM - 1) / KC_BM;
  const int nk = (K + 31) >> 5;
  const int n0 = blockIdx.y * BN;
  constexpr uint32_t tmem_cols = 2 * BN;  // 128 or 256: power of two >= 32

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], KC_PW * 32);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&acc_full[s], 1);
      mbar_init(&acc_empty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == KC_PW) tmem_alloc(tmem_slot, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int my_tiles = (ntiles > (int)blockIdx.x) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp < KC_PW) {
    // ===================== producers =====================
    constexpr int RPI = KC_PW * 4;                  // rows the producer threads cover per iteration (8 threads per row)
    constexpr int ITERS = KC_BM / RPI;              // A row-iterations per K chunk
    constexpr int U = (AMODE == OP_BNBWD) ? ITERS / 2 : ITERS;  // A row-iterations per pipeline unit
    constexpr int HPC = ITERS / U;                  // units per K chunk
    constexpr int WU = (BN / RPI) / HPC;            // W row-iterations per unit
    static_assert(U >= 1 && WU >= 1, "producer mapping");
    const int kq = tid & 7, r0 = tid >> 3;          // float4 inside the 32-wide chunk; first row served
    struct Regs {
      float4 x[U], y[U];
      float w[U];
      float4 b[WU];
      Consts4<AMODE> cc;
    };
    struct Cursor {
      int it, kc, h, g;  // tile iteration, K chunk, half, global chunk counter
    };
    const int total_units = my_tiles * nk * HPC;
    auto advance = [&](Cursor& c) {
      if (++c.h == HPC) {
        c.h = 0;
        ++c.g;
        if (++c.kc == nk) {
          c.kc = 0;
          ++c.it;
        }
      }
    };
    auto issue = [&](Regs& R, const Cursor& c) {
      const int row0 = ((int)blockIdx.x + c.it * (int)gridDim.x) * KC_BM;
      const int col = (c.kc << 5) + (kq << 2);
      const bool colok = col < K;
      R.cc.load(p.A, col, colok);
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int row = row0 + r0 + RPI * (c.h * U + k);
        R.x[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        R.y[k] = R.x[k];
        R.w[k] = 1.f;
        if (row < M && colok) {
          R.x[k] = ldg4(p.A.X + (long long)row * p.A.ldx + col);
          if (AMODE == OP_BNBWD) {
            R.y[k] = ldg4(p.A.Y + (long long)row * p.A.ldy + col);
            if (p.A.rw) R.w[k] = p.A.rw[row];
          }
        }
      }
#pragma unroll
      for (int k = 0; k < WU; ++k) {
        const int n = n0 + r0 + RPI * (c.h * WU + k);
        R.b[k] = (n < N && colok) ? ldg4(p.Bw + (long long)n * p.ldb + col) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    auto process = [&](const Regs& R, const Cursor& c) {
      const int s = c.g % NSTAGE;
      const int row0 = ((int)blockIdx.x + c.it * (int)gridDim.x) * KC_BM;
      const bool colok = ((c.kc << 5) + (kq << 2)) < K;
      if (c.h == 0) mbar_wait(&empty[s], ((uint32_t)(c.g / NSTAGE) & 1u) ^ 1u);
      unsigned char* st = smem + (uint32_t)s * L::STAGE_BYTES;
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const int r = r0 + RPI * (c.h * U + k);
        float4 v = (row0 + r < M && colok) ? R.cc.apply(R.x[k], R.y[k], R.w[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
        split_store(st, st + L::A_BYTES, blk_off(r, kq), v);
      }
#pragma unroll
      for (int k = 0; k < WU; ++k) {
        const int r = r0 + RPI * (c.h * WU + k);
        split_store(st + 2 * L::A_BYTES, st + 2 * L::A_BYTES + L::W_BYTES, blk_off(r, kq), R.b[k]);
      }
      if (c.h == HPC - 1) {
        fence_proxy_async();
        mbar_arrive(&full[s]);
      }
    };
    Regs RA, RB;
    Cursor ci = {0, 0, 0, 0}, cp = {0, 0, 0, 0};
    if (total_units > 0) {
      issue(RA, ci);
      advance(ci);
    }
#pragma unroll 1
    for (int u = 0; u < total_units; u += 2) {
      if (u + 1 < total_units) {
        issue(RB, ci);
        advance(ci);
      }
      process(RA, cp);
      advance(cp);
      if (u + 2 < total_units) {
        issue(RA, ci);
        advance(ci);
      }
      if (u + 1 < total_units) {
        process(RB, cp);
        advance(cp);
      }
    }
  } else if (warp == KC_PW) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(KC_BM >> 4) << 24);
      const uint32_t sbase = smem_u32(smem);
      int g = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int b = it & 1;
        mbar_wait(&acc_empty[b], ((uint32_t)(it >> 1) & 1u) ^ 1u);
        const uint32_t d_tmem = tmem_base + (uint32_t)(b * BN);
        uint32_t acc = 0;
        for (int kc = 0; kc < nk; ++kc, ++g) {
          const int s = g % NSTAGE;
          mbar_wait(&full[s], (uint32_t)(g / NSTAGE) & 1u);
          tc_fence_after();
          const uint32_t st = sbase + (uint32_t)s * L::STAGE_BYTES;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t a_hi = make_desc(st + ks * 32), a_lo = make_desc(st + L::A_BYTES + ks * 32);
            const uint64_t b_hi = make_desc(st + 2 * L::A_BYTES + ks * 32), b_lo = make_desc(st + 2 * L::A_BYTES + L::W_BYTES + ks * 32);
            umma_tf32(d_tmem, a_lo, b_hi, idesc, acc);
            umma_tf32(d_tmem, a_hi, b_lo, idesc, 1u);
            umma_tf32(d_tmem, a_hi, b_hi, idesc, 1u);
            acc = 1u;
          }
          umma_commit(&empty[s]);
        }
        umma_commit(&acc_full[b]);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int ew = warp - KC_PW - 1;
    float* stage = reinterpret_cast<float*>(smem + L::OFF_STAGEBUF) + ew * 32 * EPI_LD;
    const bool do_stats = (p.stats != nullptr);
    const int rsub = lane >> 3;
    constexpr int NCB = BN / 32;
    float s0[NCB][4], s1[NCB][4];
#pragma unroll
    for (int c = 0; c < NCB; ++c)
#pragma unroll
      for (int k = 0; k < 4; ++k) s0[c][k] = s1[c][k] = 0.f;
    for (int it = 0; it < my_tiles; ++it) {
      const int t = (int)blockIdx.x + it * (int)gridDim.x;
      const int b = it & 1;
      const int row_base = t * KC_BM + q * 32;
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
      mbar_wait(&acc_full[b], (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int cb = 0; cb < NCB; ++cb) {
        if (n0 + cb * 32 < N)
          epi_block32<EMODE>(p, tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * BN + cb * 32), stage, lane, row_base,
                             n0 + cb * 32, M, N, wr, do_stats, s0[cb], s1[cb]);
      }
      tc_fence_before();
      if (lane == 0) mbar_arrive(&acc_empty[b]);
    }
    if (do_stats) {
#pragma unroll
      for (int cb = 0; cb < NCB; ++cb) {
        epi_reduce_stats(s0[cb]);
        epi_reduce_stats(s1[cb]);
        if (lane < 8) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            stat_comb[ew * BN + cb * 32 + lane * 4 + k] = s0[cb][k];
            stat_comb[4 * BN + ew * BN + cb * 32 + lane * 4 + k] = s1[cb][k];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (p.stats) {
    for (int c = tid; c < BN; c += KC_NT_THREADS) {
      if (n0 + c < N) {
        float a0 = stat_comb[c] + stat_comb[BN + c] + stat_comb[2 * BN + c] + stat_comb[3 * BN + c];
        float a1 = stat_comb[4 * BN + c] + stat_comb[5 * BN + c] + stat_comb[6 * BN + c] + stat_comb[7 * BN + c];
        p.stats[(long long)blockIdx.x * 2 * N + n0 + c] = a0;
        p.stats[(long long)blockIdx.x * 2 * N + N + n0 + c] = a1;
      }
    }
    if (p.tail.kind == 0)   // (a fused tail reads the gridDim.x live slots only)
      for (int slot = blockIdx.x + gridDim.x; slot < GADDPG_STAT_SLOTS; slot += gridDim.x)
        for (int c = tid; c < 2 * BN; c += KC_NT_THREADS) {
          const int cc = (c < BN) ? c : c - BN;
          if (n0 + cc < N) p.stats[(long long)slot * 2 * N + (c < BN ? 0 : N) + n0 + cc] = 0.f;
        }
  }
  if (warp == KC_PW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
  // BatchNorm finalize by the last CTA of the (m-tile, n-tile) grid: slot = blockIdx.x, the n-tiles own disjoint columns
  if (p.stats && p.tail.kind != 0)
    bnt_run<13>(p.tail, p.stats, N, ntiles < (int)gridDim.x ? ntiles : (int)gridDim.x, gridDim.x * gridDim.y, reinterpret_cast<double*>(smem), 8192);
}

// =====================================================================================================
// TN (weight gradients): partial[split][N][K] tile [128 x BKT] accumulated in TMEM over this split's row chunks
// =====================================================================================================
constexpr int TN_R = 32;  // rows (reduction index) per stage = 4 MMA K-steps of 8

// cp.async helpers of the TN producers (LDGSTS: global -> shared without a register stop; src-size 0 zero-fills)
__device__ __forceinline__ void tn_cp16(unsigned char* dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(n) : "memory");
}
// 16-byte copy of which only the first `bytes` (0..16) are read; the rest of the destination is zero-filled
__device__ __forceinline__ void tn_cp16n(unsigned char* dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tn_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tn_cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Shared memory: NSTAGE operand stages in the UMMA layout (hi / lo of P and Q) + a RAW ring of D units.  A unit is half a
// chunk (16 rows): the raw rows of P (X, and Y / row weights / arg-max bits for the BatchNorm-backward forms) and Q exactly as
// they lie in global memory, landed by cp.async — so the bytes in flight per SM are bounded by shared memory (D units), not by
// the producers' registers.  Every thread copies, and later transforms, only its own 16-byte pieces: no barrier between the
// copy and the transform, just cp.async.wait_group.
// 16 producer warps (the cp.async staging leaves them at <= 96 registers): four per scheduler instead of two — their streams are
// chains of LDS -> FMA -> STS with little ILP, so the transform rate scales with the warps in flight.
constexpr int TN_PW = 16;
constexpr int TN_THREADS = (TN_PW + 1 + 4) * 32;

template <int BKT, int PMODE, bool P64>
struct TnLayout {
  static_assert(!(P64 && BKT != 64), "the 64-channel P mapping is built for BKT = 64");
  static constexpr bool PBWD = (PMODE == OP_BNBWD || PMODE == OP_BNBWD_POOL);
  static constexpr bool POOL = (PMODE == OP_BNBWD_POOL);
  static constexpr int NSTAGE = 2;
  static constexpr uint32_t P_BYTES = 4 * TN_R * 128;            // 128 channels = 4 blocks of [32 rows x 128 B]
  static constexpr uint32_t Q_BYTES = (BKT / 32) * TN_R * 128;
  static constexpr uint32_t STAGE_BYTES = 2 * P_BYTES + 2 * Q_BYTES;
  static constexpr int NT = TN_PW * 32;                          // producer threads (512)
  // unit = what one producer pass covers: P64: a whole chunk (32 rows x 16 float4 of P, 32 x 16 of Q); otherwise half a chunk of P
  // (16 rows x 32 float4) plus, for BKT = 64, the whole chunk of Q (32 rows x 16 float4) in the first half only, for BKT = 128
  // the same 16 rows of Q (16 x 32 float4).  One float4 of P (and of Y) and at most one of Q per thread and unit.
  static constexpr int HPC = P64 ? 1 : 2;                        // units per chunk
  static constexpr int UROWS = TN_R / HPC;                       // P rows per unit
  static constexpr int QEVERY = (!P64 && BKT == 64) ? 2 : 1;     // Q is issued in units with h % QEVERY == 0 ...
  static constexpr int QROWS = NT / (BKT / 4);                   // ... and covers this many rows (32 or 16)
  static constexpr uint32_t RAW_PX = NT * 16, RAW_PY = PBWD ? RAW_PX : 0, RAW_Q = NT * 16;
  // row weights / arg-max bit masks of a unit's rows: copied once per WARP (4-byte cp.async serialise 32-fold in the shared
  // memory pipe: measured), 16 bytes per lane, read back by every lane after a __syncwarp
  static constexpr uint32_t RAW_W = PBWD ? TN_PW * UROWS * 4 : 0, RAW_M = POOL ? TN_PW * UROWS * 16 : 0;
  static constexpr uint32_t OFF_PY = RAW_PX, OFF_Q = OFF_PY + RAW_PY, OFF_W = OFF_Q + RAW_Q, OFF_M = OFF_W + RAW_W;
  static constexpr uint32_t UNIT_BYTES = OFF_M + RAW_M;
  static constexpr uint32_t OFF_RAW = NSTAGE * STAGE_BYTES;
  static constexpr uint32_t TAIL_BYTES = 256 + NT * 16;          // barriers + bias combine
  static constexpr int D_FIT = (int)((227u * 1024u - 1024u - OFF_RAW - TAIL_BYTES) / UNIT_BYTES);
  static constexpr int D = D_FIT > 8 ? 8 : D_FIT;                // units in flight
  static constexpr uint32_t OFF_BARS = OFF_RAW + D * UNIT_BYTES;
  static constexpr uint32_t OFF_BIAS = OFF_BARS + 256;
  static constexpr uint32_t TOTAL = OFF_BIAS + NT * 16;
  static_assert(D >= 3, "raw ring too shallow");
};

// P64: the P operand has <= 64 channels (dW of a 64-wide layer): its producers cover 32 rows x 16 float4 per pass instead of
// 16 rows x 32 float4 with half of the threads idle; P blocks 2, 3 of every stage stay zero (zeroed once).
template <int BKT, int PMODE, int QMODE, bool P64>
__global__ void __launch_bounds__(TN_THREADS, 1) tc_tn_kernel(const TNProblem p, float* __restrict__ partial,
                                                               float* __restrict__ partial_bias, int splits, int tiles_k) {
  using L = TnLayout<BKT, PMODE, P64>;
  constexpr int NSTAGE = L::NSTAGE;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BARS);
  uint64_t* full = bars;             // [NSTAGE]
  uint64_t* empty = bars + NSTAGE;   // [NSTAGE]
  uint64_t* acc_full = bars + 2 * NSTAGE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * NSTAGE + 1);
  float* bias_comb = reinterpret_cast<float*>(smem + L::OFF_BIAS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int M = p.M_dev ? *p.M_dev : p.M_max;
  M = M < p.M_max ? M : p.M_max;
  const int N = p.N, K = p.K;
  const int tile = blockIdx.x, split = blockIdx.y;
  const int n0 = (tile / tiles_k) * 128, k0 = (tile % tiles_k) * BKT;
  const bool want_bias = (partial_bias != nullptr) && (tile % tiles_k == 0);
  const int nchunks = (M + TN_R - 1) / TN_R;
  const int my_chunks = (nchunks > split) ? (nchunks - 1 - split) / splits + 1 : 0;
  constexpr uint32_t tmem_cols = BKT;  // 64 or 128

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], TN_PW * 32);
      mbar_init(&empty[s], 1);
    }
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == TN_PW) tmem_alloc(tmem_slot, tmem_cols);
  if (P64) {  // P blocks 2, 3 (channels 64..127) of every stage: zero, never written again
    for (int i = tid; i < NSTAGE * 2 * (int)(L::P_BYTES / 2) / 16; i += TN_THREADS) {
      const int s = i / (2 * (int)(L::P_BYTES / 2) / 16), r = i % (2 * (int)(L::P_BYTES / 2) / 16);
      const int hl = r / ((int)(L::P_BYTES / 2) / 16), o = r % ((int)(L::P_BYTES / 2) / 16);
      *reinterpret_cast<float4*>(smem + (uint32_t)s * L::STAGE_BYTES + hl * L::P_BYTES + L::P_BYTES / 2 + o * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < TN_PW) {
    // ===================== producers =====================
    constexpr bool PBWD = L::PBWD, POOL = L::POOL;
    constexpr int HPC = L::HPC, D = L::D, QEVERY = L::QEVERY, QROWS = L::QROWS, PU = 1;
    constexpr int PQ4 = P64 ? 16 : 32;                // float4 per P row
    constexpr int PROWS = L::NT / PQ4;                // rows the producer threads cover per P pass (32 or 16)
    static_assert(PROWS == L::UROWS && QROWS * HPC == TN_R * QEVERY, "producer mapping");
    const int pq = tid % PQ4, pr = tid / PQ4;         // P: float4 column, first row
    const int qq = tid % (BKT / 4), qr = tid / (BKT / 4);
    const int pcol = n0 + (pq << 2), qcol = k0 + (qq << 2);
    const bool pok = pcol < N, qok = qcol < K;
    Consts4<PMODE> pc;
    Consts4<QMODE> qc;
    pc.load(p.P, pcol, pok);
    qc.load(p.Q, qcol, qok);
    const uint32_t poff_blk = (uint32_t)(pq >> 3) * (TN_R * 128), qoff_blk = (uint32_t)(qq >> 3) * (TN_R * 128);
    float4 bsum = make_float4(0.f, 0.f, 0.f, 0.f);
    const int total_units = my_chunks * HPC;
    unsigned char* raw = smem + L::OFF_RAW;
    const uint32_t my16 = (uint32_t)tid * 16;
    constexpr int UROWS = L::UROWS;
    const int mwords = POOL ? (p.P.ldx >> 5) : 0;     // arg-max mask words per row (<= 4)
    int segn[PU];  // BNBWD_POOL: segment of the row of the NEXT unit to be issued (dependent load, one unit ahead)
    auto prefetch_seg = [&](int u) {
      const int ci = u / HPC, h = u % HPC;
      const int rowc = (split + ci * splits) * TN_R;
#pragma unroll
      for (int k = 0; k < PU; ++k) {
        const int row = rowc + pr + PROWS * (h * PU + k);
        segn[k] = (u < total_units && row < M && pok) ? p.P.pseg[row] : 0;
      }
    };
    auto issue = [&](int u) {
      const int ci = u / HPC, h = u % HPC;
      const int rowc = (split + ci * splits) * TN_R;
      unsigned char* slot = raw + (uint32_t)(u % D) * L::UNIT_BYTES;
#pragma unroll
      for (int k = 0; k < PU; ++k) {
        const int row = rowc + pr + PROWS * (h * PU + k);
        const bool ok = row < M && pok;
        const long long r = ok ? row : 0;
        const int c = pok ? pcol : 0;
        if (!POOL) {
          tn_cp16(slot + k * L::NT * 16 + my16, p.P.X + r * p.P.ldx + c, ok);
        } else {  // max-pool gradient rebuilt from E (S x C, L2 resident) + the arg-max bit mask (see tc_gemm.cu)
          tn_cp16(slot + k * L::NT * 16 + my16, p.P.X + (long long)segn[k] * p.P.ldx + c, ok);
        }
        if (PBWD) tn_cp16(slot + L::OFF_PY + k * L::NT * 16 + my16, p.P.Y + r * p.P.ldy + c, ok);
      }
      {  // per-warp copies of the unit's row weights (16 rows x 4 B) and arg-max masks (16 rows x mwords x 4 B)
        const int R0 = rowc + UROWS * h;
        if (PBWD && p.P.rw && lane < UROWS / 4) {
          int nb = (M - (R0 + 4 * lane)) * 4;
          nb = nb < 0 ? 0 : (nb > 16 ? 16 : nb);
          tn_cp16n(slot + L::OFF_W + warp * (UROWS * 4) + lane * 16, p.P.rw + (nb > 0 ? R0 + 4 * lane : 0), nb);
        }
        if (POOL && lane < (UROWS * mwords) / 4) {
          int nb = (M - R0) * mwords * 4 - 16 * lane;
          nb = nb < 0 ? 0 : (nb > 16 ? 16 : nb);
          tn_cp16n(slot + L::OFF_M + warp * (UROWS * 16) + lane * 16, p.P.pmask + (nb > 0 ? (long long)R0 * mwords + 4 * lane : 0), nb);
        }
      }
      if (POOL) prefetch_seg(u + 1);
      if (h % QEVERY == 0) {
        const int row = rowc + qr + QROWS * (h / QEVERY);
        const bool ok = row < M && qok;
        tn_cp16(slot + L::OFF_Q + my16, p.Q.X + (long long)(ok ? row : 0) * p.Q.ldx + (qok ? qcol : 0), ok);
      }
    };
    auto process = [&](int u) {
      const int ci = u / HPC, h = u % HPC;
      const int s = ci % NSTAGE;
      const int rowc = (split + ci * splits) * TN_R;
      const unsigned char* slot = raw + (uint32_t)(u % D) * L::UNIT_BYTES;
      if (h == 0) mbar_wait(&empty[s], ((uint32_t)(ci / NSTAGE) & 1u) ^ 1u);
      unsigned char* st = smem + (uint32_t)s * L::STAGE_BYTES;
      if (PBWD) __syncwarp();   // the warp's shared row-weight / mask copies are visible to all of its lanes
#pragma unroll
      for (int k = 0; k < PU; ++k) {
        const int r = pr + PROWS * (h * PU + k);
        float4 x = *reinterpret_cast<const float4*>(slot + k * L::NT * 16 + my16);
        float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
        float w = 1.f;
        if (PBWD) {
          y = *reinterpret_cast<const float4*>(slot + L::OFF_PY + k * L::NT * 16 + my16);
          if (p.P.rw) w = *reinterpret_cast<const float*>(slot + L::OFF_W + warp * (UROWS * 4) + (r - UROWS * h) * 4);
        }
        if (POOL) {
          const uint32_t bits =
              *reinterpret_cast<const uint32_t*>(slot + L::OFF_M + warp * (UROWS * 16) + ((r - UROWS * h) * mwords + (pcol >> 5)) * 4) >> (pcol & 31);
          x.x = (bits & 1u) ? x.x : 0.f;
          x.y = (bits & 2u) ? x.y : 0.f;
          x.z = (bits & 4u) ? x.z : 0.f;
          x.w = (bits & 8u) ? x.w : 0.f;
        }
        float4 v = (rowc + r < M && pok) ? pc.apply(x, y, w) : make_float4(0.f, 0.f, 0.f, 0.f);
        bsum.x += v.x;
        bsum.y += v.y;
        bsum.z += v.z;
        bsum.w += v.w;
        split_store(st, st + L::P_BYTES, poff_blk + blk_off_mn(r, pq & 7), v);
      }
      if (h % QEVERY == 0) {
        const int r = qr + QROWS * (h / QEVERY);
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 qv = *reinterpret_cast<const float4*>(slot + L::OFF_Q + my16);
        float4 v = (rowc + r < M && qok) ? qc.apply(qv, z, 1.f) : z;
        split_store(st + 2 * L::P_BYTES, st + 2 * L::P_BYTES + L::Q_BYTES, qoff_blk + blk_off_mn(r, qq & 7), v);
      }
      if (h == HPC - 1) {
        fence_proxy_async();
        mbar_arrive(&full[s]);
      }
    };
    if (POOL) prefetch_seg(0);
#pragma unroll
    for (int u = 0; u < D - 1; ++u) {   // fill the ring: D - 1 units in flight before the first transform
      if (u < total_units) issue(u);
      tn_cp_commit();
    }
#pragma unroll 1
    for (int u = 0; u < total_units; ++u) {
      if (u + D - 1 < total_units) issue(u + D - 1);
      tn_cp_commit();                   // (possibly empty) group: exactly one per iteration keeps the wait count static
      tn_cp_wait<D - 1>();              // this thread's copies of unit u have landed
      process(u);
    }
    if (want_bias) *reinterpret_cast<float4*>(bias_comb + pr * (PQ4 * 4) + (pq << 2)) = bsum;   // [PROWS row groups][channels]
  } else if (warp == TN_PW) {
    // ===================== MMA issuer =====================
    if (lane == 0 && my_chunks > 0) {
      // both operands MN-major (bits 15, 16): the reduction index (rows) is the MMA K dimension
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BKT >> 3) << 17) |
                             ((uint32_t)(128 >> 4) << 24);
      const uint32_t sbase = smem_u32(smem);
      uint32_t acc = 0;
      for (int ci = 0; ci < my_chunks; ++ci) {
        const int s = ci % NSTAGE;
        mbar_wait(&full[s], (uint32_t)(ci / NSTAGE) & 1u);
        tc_fence_after();
        const uint32_t st = sbase + (uint32_t)s * L::STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < TN_R / 8; ++ks) {
          const uint32_t o = ks * 1024;
          const uint64_t a_hi = make_desc_mn(st + o, TN_R * 128, 512), a_lo = make_desc_mn(st + L::P_BYTES + o, TN_R * 128, 512);
          const uint64_t b_hi = make_desc_mn(st + 2 * L::P_BYTES + o, TN_R * 128, 512);
          const uint64_t b_lo = make_desc_mn(st + 2 * L::P_BYTES + L::Q_BYTES + o, TN_R * 128, 512);
          umma_tf32(tmem_base, a_lo, b_hi, idesc, acc);
          umma_tf32(tmem_base, a_hi, b_lo, idesc, 1u);
          umma_tf32(tmem_base, a_hi, b_hi, idesc, 1u);
          acc = 1u;
        }
        umma_commit(&empty[s]);
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else {
    // ===================== epilogue: TMEM -> partial[split] =====================
    const int q = warp & 3;
    const int n = n0 + q * 32 + lane;
    float* out = partial + ((long long)split * N + n) * K + k0;
    if (my_chunks > 0) {
      mbar_wait_sleep(acc_full, 0, 500);
      tc_fence_after();
    }
#pragma unroll
    for (int cb = 0; cb < BKT / 32; ++cb) {
      float r[32];
      if (my_chunks > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cb * 32), r);
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) r[c] = 0.f;
      }
      if (n < N) {
#pragma unroll
        for (int c = 0; c < 32; c += 4)
          if (k0 + cb * 32 + c < K) *reinterpret_cast<float4*>(out + cb * 32 + c) = make_float4(r[c], r[c + 1], r[c + 2], r[c + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (want_bias && tid < (P64 ? 64 : 128) && n0 + tid < N) {   // row groups folded in group order (deterministic)
    constexpr int NG = TN_PW * 32 / (P64 ? 16 : 32), W = P64 ? 64 : 128;
    float b = 0.f;
#pragma unroll 8
    for (int g = 0; g < NG; ++g) b += bias_comb[g * W + tid];
    partial_bias[(long long)split * N + n0 + tid] = b;
  }
  if (warp == TN_PW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

}  // namespace

// ----------------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------------
bool gaddpg_tc_nt_kc_supported(const NTProblem& p, int amode, int emode) {
  // below ~1k rows a tile grid cannot fill the chip and the serial K loop loses to the FFMA kernel's 32x64 tiles (measured)
  if (amode == OP_BNBWD_POOL) return false;  // only tc_gemm.cu and the TN kernel rebuild the pool gradient on the fly
  if (p.M_max < 1024 || p.N < 32 || p.K < 32) return false;
  if (p.ldb % 4 != 0) return false;
  if (!tc_epilogue_ok(p, emode)) return false;
  (void)amode;
  return true;
}

int gaddpg_tc_nt_kc_impl(const NTProblem* p, int amode, int emode, void* stream) {
  const int mtiles = ceil_div(p->M_max, KC_BM);
  const bool wide = p->N >= 128 && mtiles * ceil_div(p->N, 128) >= 64;
  const int BN = wide ? 128 : 64;
  dim3 grid(mtiles < GADDPG_STAT_SLOTS ? mtiles : GADDPG_STAT_SLOTS, ceil_div(p->N, BN));
  cudaStream_t st = (cudaStream_t)stream;
#define KC_LAUNCH(BNN, A, E)                                                                                   \
  {                                                                                                            \
    auto kern = tc_nt_kc_kernel<BNN, A, E>;                                                                    \
    const size_t smem = KcLayout<BNN>::TOTAL + 1024;                                                           \
    GADDPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    kern<<<grid, KC_NT_THREADS, smem, st>>>(*p);                                                                \
    GADDPG_CHECK_LAUNCH("tc_nt_kc_kernel");                                                                    \
    return GADDPG_OK;                                                                                          \
  }
#define KC_CASE(A, E)                                                                                          \
  if (amode == A && emode == E) {                                                                              \
    if (BN == 128) KC_LAUNCH(128, A, E) else KC_LAUNCH(64, A, E)                                               \
  }
  KC_CASE(OP_PLAIN, EPI_STORE)
  KC_CASE(OP_BNRELU, EPI_STORE)
  KC_CASE(OP_BNBWD, EPI_DMASK)
  KC_CASE(OP_BNBWD, EPI_STORE)
  KC_CASE(OP_PLAIN, EPI_DMASK)
#undef KC_CASE
#undef KC_LAUNCH
  gaddpg_set_error("tc_nt_kc: unsupported mode pair (%d,%d)", amode, emode);
  return GADDPG_ERR_UNSUPPORTED;
}

bool gaddpg_tc_tn_supported(const TNProblem& p, int pmode, int qmode) {
  if (p.N < 32 || p.K < 32 || p.M_max < 64) return false;
  if (pmode == OP_BNRELU || qmode == OP_BNBWD || qmode == OP_BNBWD_POOL) return false;
  if (pmode == OP_BNBWD_POOL && !(qmode == OP_BNRELU && p.P.pmask && p.P.pseg && p.P.ldx % 32 == 0))
    return false;
  return true;
}

// launches the split kernel only; the caller runs the fixed-order reduction over `*splits_out` partials
int gaddpg_tc_tn_impl(const TNProblem* p, int pmode, int qmode, float* ws, size_t ws_floats, float* ws_bias, int* splits_out,
                      void* stream) {
  const int BKT = (p->K <= 64) ? 64 : 128;
  const int tiles_k = ceil_div(p->K, BKT);
  const int tiles = ceil_div(p->N, 128) * tiles_k;
  const int chunks = ceil_div(p->M_max, TN_R);
  int splits = gaddpg_sm_count() / tiles;
  if (splits > chunks) splits = chunks;
  if ((size_t)splits * p->N * p->K > ws_floats) splits = (int)(ws_floats / ((size_t)p->N * p->K));
  if (splits > 64 * 1024 / p->N) splits = 64 * 1024 / p->N;
  if (splits < 1) splits = 1;
  GADDPG_CHECK_ARG((size_t)splits * p->N * p->K <= ws_floats, "tc_tn: workspace too small");
  *splits_out = splits;
  dim3 grid(tiles, splits);
  cudaStream_t st = (cudaStream_t)stream;
#define TN_LAUNCH(BK_, PM, QM)                                                                                 \
  if (p->N <= 64 && BK_ == 64) TN_LAUNCH_(64, PM, QM, true) else TN_LAUNCH_(BK_, PM, QM, false)
#define TN_LAUNCH_(BK_, PM, QM, P64_)                                                                          \
  {                                                                                                            \
    auto kern = tc_tn_kernel<BK_, PM, QM, P64_>;                                                               \
    const size_t smem = TnLayout<BK_, PM, P64_>::TOTAL + 1024;                                                 \
    GADDPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));           \
    kern<<<grid, TN_THREADS, smem, st>>>(*p, ws, ws_bias, splits, tiles_k);                                    \
    GADDPG_CHECK_LAUNCH("tc_tn_kernel");                                                                       \
    return GADDPG_OK;                                                                                          \
  }
#define TN_CASE(PM, QM)                                                                                        \
  if (pmode == PM && qmode == QM) {                                                                            \
    if (BKT == 128) TN_LAUNCH(128, PM, QM) else TN_LAUNCH(64, PM, QM)                                          \
  }
  TN_CASE(OP_PLAIN, OP_PLAIN)
  TN_CASE(OP_PLAIN, OP_BNRELU)
  TN_CASE(OP_BNBWD, OP_PLAIN)
  TN_CASE(OP_BNBWD, OP_BNRELU)
  TN_CASE(OP_BNBWD_POOL, OP_BNRELU)
#undef TN_CASE
#undef TN_LAUNCH
#undef TN_LAUNCH_
  gaddpg_set_error("tc_tn: unsupported mode pair (%d,%d)", pmode, qmode);
  return GADDPG_ERR_UNSUPPORTED;
}
