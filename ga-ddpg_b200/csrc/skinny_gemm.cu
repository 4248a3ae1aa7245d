// skinny_gemm.cu — row-GEMM for the few-hundred-row problems of the step (sm_100a).
//
//   C[M,N] = epi( pro(A)[M,K] . W[N,K]^T ),   M = replay batch (256) or a few hundred rows, K up to 1024
//
// These are the FC/BatchNorm1d head of the encoder (Linear 512->1024->512, /root/reference/core/networks.py:84-91) and
// the actor / critic Linear layers (networks.py:265-300, 315-351), forward and dX.  With M = 256 there are only a
// handful of output tiles, so the cost is the latency of the serial K loop, not arithmetic: the FP32 FFMA kernel of
// gemm_rows.cu (one register-prefetched 16-wide chunk in flight) spends ~0.9 us per chunk, 60 us for K = 1024.
// Here every CTA (32 x 64 output tile, 8 warps) streams its A and W K-chunks through a 5-stage cp.async ring, so
// ~70 KB per CTA are in flight, and the products run on the tensor cores as warp-level mma.sync m16n8k8 TF32 with the
// same 3xTF32 split as the tcgen05 kernels (x = hi + lo; lo.hi + hi.lo + hi.hi; one 32-wide chunk accumulates in the
// tensor core, chunks are folded with IEEE FP32 adds: 7e-7 relative error at K = 1024) — a 32 x 64 tile is too
// small for a tcgen05 (M = 128) instruction, and FFMA on such a tile is shared-memory-bandwidth bound.
// Operand prologues (BN+ReLU, BN-backward) are applied in place in shared memory by the thread that issued the copy;
// epilogues and the deterministic per-CTA statistics slots are those of gemm_rows.cu.
#include "common.cuh"
#include "gemm_rows.cuh"
#include "impl.h"
#include "bn_tail.cuh"

namespace {

constexpr int SK_BM = 32, SK_BK = 32, SK_STAGES = 5;  // output tile 32 x BN (BN = 64, or 32 when that is needed to fill the chip)
constexpr int SK_LD = SK_BK + 4;  // row pitch (floats): 16-byte aligned rows, conflict-free fragment loads (bank = 4g + t)
constexpr int SK_A_FLOATS = SK_BM * SK_LD;

__device__ __forceinline__ float4 sk_ldg4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int n = valid ? 16 : 0;  // src-size 0: the 16 destination bytes are zero-filled, nothing is read
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}

// per-column prologue constants of the float4 this thread copies (reloaded per K chunk, one chunk ahead)
template <int MODE>
struct SkConsts {
  float4 c0, c1, c2, c3, c4;
  __device__ __forceinline__ void load(const Operand& d, int col, bool ok) {
    if (!ok) return;
    if (MODE == OP_BNRELU) {
      c0 = sk_ldg4(d.c0 + col);
      c1 = sk_ldg4(d.c1 + col);
    } else if (MODE == OP_BNBWD) {
      c0 = sk_ldg4(d.c0 + col);
      c1 = sk_ldg4(d.c1 + col);
      c2 = sk_ldg4(d.c2 + col);
      c3 = sk_ldg4(d.c3 + col);
      c4 = sk_ldg4(d.c4 + col);
    }
  }
  __device__ __forceinline__ float4 apply(float4 x, float4 y, float w) const {
    float4 r;
    if (MODE == OP_BNRELU) {
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

// 8 warps = NWM x NWN: warp (wm, wn) owns MFW 16-row fragments x the 8 columns [8 wn, 8 wn + 8) of the 32 x BN tile.
// BN = 64: 1 x 8 warps, two m-fragments each; BN = 32: 2 x 4 warps, one m-fragment each (half the MMAs per CTA, twice the CTAs).
template <int AMODE, int EMODE, int BN>
__global__ void __launch_bounds__(256) skinny_nt_kernel(const NTGroup grp) {
  extern __shared__ __align__(16) float sk_smem[];
  __shared__ float s_acc[2 * 1024];  // per-CTA running (sum, sum2) per output column (N <= 1024 when stats on)
  __shared__ float s_part[2][2][64]; // [row half wm][sum | sum2][tile column]: fixed-order fold when NWM = 2
  constexpr int SK_BN = BN, SK_B_FLOATS = BN * SK_LD;
  constexpr int NWN = BN / 8, NWM = 8 / NWN, MFW = 2 / NWM;
  constexpr int STAGE_FLOATS = SK_A_FLOATS * (AMODE == OP_BNBWD ? 2 : 1) + SK_B_FLOATS;
  static_assert(SK_STAGES * STAGE_FLOATS / 2 >= BNT_SCRATCH_DOUBLES, "the cp.async ring must hold the BatchNorm tail's scratch");

  const NTProblem& p = grp.p[blockIdx.y];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int wn = warp % NWN, wm = warp / NWN;
  int M = p.M_dev ? *p.M_dev : p.M_max;
  M = M < p.M_max ? M : p.M_max;
  const int N = p.N, K = p.K;
  const bool do_stats = (p.stats != nullptr);
  if (do_stats)
    for (int c = tid; c < 2 * N; c += 256) s_acc[c] = 0.f;

  const int tiles_m = (M + SK_BM - 1) / SK_BM, tiles_n = (N + SK_BN - 1) / SK_BN, nk = (K + SK_BK - 1) / SK_BK;
  // copy assignment: A (and Y) tile = 32 rows x 8 float4 -> one per thread; W tile = 64 rows x 8 float4 -> two per thread
  const int ar = tid >> 3, aq = (tid & 7) << 2;

  for (int tile = blockIdx.x; tile < tiles_m * tiles_n; tile += gridDim.x) {
    const int row0 = (tile / tiles_n) * SK_BM, col0 = (tile % tiles_n) * SK_BN;
    const int arow = row0 + ar;
    const bool arow_ok = arow < M;
    float rw = 1.f;
    if (AMODE == OP_BNBWD && p.A.rw && arow_ok) rw = p.A.rw[arow];
    __syncthreads();  // previous tile's shared-memory reads (and the s_acc initialisation) are complete

    auto issue = [&](int kc) {
      float* st = sk_smem + (kc % SK_STAGES) * STAGE_FLOATS;
      const int k = kc * SK_BK + aq;
      const bool kok = k < K;
      cp_async16(st + ar * SK_LD + aq, p.A.X + (long long)(arow_ok ? arow : 0) * p.A.ldx + (kok ? k : 0), arow_ok && kok);
      if (AMODE == OP_BNBWD)
        cp_async16(st + SK_A_FLOATS + ar * SK_LD + aq, p.A.Y + (long long)(arow_ok ? arow : 0) * p.A.ldy + (kok ? k : 0),
                   arow_ok && kok);
      float* bs = st + SK_A_FLOATS * (AMODE == OP_BNBWD ? 2 : 1);
#pragma unroll
      for (int i = 0; i < BN / 32; ++i) {
        const int br = ar + 32 * i, n = col0 + br;
        const bool ok = n < N && kok;
        cp_async16(bs + br * SK_LD + aq, p.Bw + (long long)(n < N ? n : 0) * p.ldb + (kok ? k : 0), ok);
      }
    };

    float acc[MFW][4];
#pragma unroll
    for (int i = 0; i < MFW; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

#pragma unroll
    for (int s = 0; s < SK_STAGES - 1; ++s) {
      if (s < nk) issue(s);
      cp_async_commit();
    }
    SkConsts<AMODE> cc;
    cc.load(p.A, aq, aq < K);

    for (int kc = 0; kc < nk; ++kc) {
      cp_async_wait<SK_STAGES - 2>();  // this thread's copies of chunk kc have landed
      float* st = sk_smem + (kc % SK_STAGES) * STAGE_FLOATS;
      if (AMODE != OP_PLAIN) {
        const bool ok = arow_ok && (kc * SK_BK + aq) < K;
        if (ok) {  // in-place prologue on the float4 this thread copied (invalid elements stay zero)
          float4* xp = reinterpret_cast<float4*>(st + ar * SK_LD + aq);
          float4 y = make_float4(0.f, 0.f, 0.f, 0.f);
          if (AMODE == OP_BNBWD) y = *reinterpret_cast<const float4*>(st + SK_A_FLOATS + ar * SK_LD + aq);
          *xp = cc.apply(*xp, y, rw);
        }
        const int kn = (kc + 1) * SK_BK + aq;
        cc.load(p.A, kn, kn < K);  // next chunk's constants: their latency hides behind this chunk's MMAs
      }
      __syncthreads();  // chunk kc visible to everyone; everyone is done with chunk kc-1 -> its stage may be refilled
      if (kc + SK_STAGES - 1 < nk) issue(kc + SK_STAGES - 1);
      cp_async_commit();

      const float* As = st;
      const float* Bs = st + SK_A_FLOATS * (AMODE == OP_BNBWD ? 2 : 1) + (wn * 8) * SK_LD;
      // the tensor core's FP32 accumulation is not IEEE round-to-nearest (its error grows with the number of
      // accumulated products: ~1e-5 at K = 1024); accumulate one 32-wide chunk there and fold chunks with FP32 adds
      float d[MFW][4];
#pragma unroll
      for (int i = 0; i < MFW; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) d[i][j] = 0.f;
#pragma unroll
      for (int k8 = 0; k8 < SK_BK; k8 += 8) {
        uint32_t bh[2], bl[2];
        split_tf32(Bs[g * SK_LD + k8 + t], bh[0], bl[0]);
        split_tf32(Bs[g * SK_LD + k8 + t + 4], bh[1], bl[1]);
#pragma unroll
        for (int mf = 0; mf < MFW; ++mf) {
          uint32_t ah[4], al[4];
          const float* a = As + (16 * (wm * MFW + mf) + g) * SK_LD + k8 + t;
          split_tf32(a[0], ah[0], al[0]);
          split_tf32(a[8 * SK_LD], ah[1], al[1]);
          split_tf32(a[4], ah[2], al[2]);
          split_tf32(a[8 * SK_LD + 4], ah[3], al[3]);
          mma_tf32(d[mf], al, bh);
          mma_tf32(d[mf], ah, bl);
          mma_tf32(d[mf], ah, bh);
        }
      }
#pragma unroll
      for (int i = 0; i < MFW; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += d[i][j];
    }
    cp_async_wait<0>();

    // ---- epilogue: thread holds rows row0 + 16 (wm MFW + mf) + g (+8), columns col0 + 8 wn + 2t (+1) ----
    float s0[2] = {0.f, 0.f}, s1[2] = {0.f, 0.f};
#pragma unroll
    for (int mf = 0; mf < MFW; ++mf) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int row = row0 + 16 * (wm * MFW + mf) + g + 8 * h;
        if (row < M) {
          float w = 1.f;
          if (EMODE == EPI_STORE && do_stats && p.srw) w = p.srw[row];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int col = col0 + 8 * wn + 2 * t + j;
            if (col < N) {
              float v = acc[mf][2 * h + j];
              if (EMODE == EPI_STORE) {
                if (p.bias) v += p.bias[col];
                if (p.relu) v = fmaxf(v, 0.f);
                p.C[(long long)row * p.ldc + col] = v;
                s0[j] = fmaf(w, v, s0[j]);
                s1[j] = fmaf(w * v, v, s1[j]);
              } else {
                const float yp = p.Yprev[(long long)row * p.ldyp + col];
                const float z = p.psc ? fmaf(yp, p.psc[col], p.psh[col]) : yp;
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
    }
    if (do_stats) {  // fold the 8 row groups (g) of the warp, then the NWM row halves, always in the same order
#pragma unroll
      for (int j = 0; j < 2; ++j) {
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          s0[j] += __shfl_xor_sync(0xffffffffu, s0[j], o);
          s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], o);
        }
        if (g == 0) {
          s_part[wm][0][8 * wn + 2 * t + j] = s0[j];
          s_part[wm][1][8 * wn + 2 * t + j] = s1[j];
        }
      }
      __syncthreads();
      if (tid < BN && col0 + tid < N) {
        float a0 = s_part[0][0][tid], a1 = s_part[0][1][tid];
        if (NWM == 2) {
          a0 += s_part[1][0][tid];
          a1 += s_part[1][1][tid];
        }
        s_acc[col0 + tid] += a0;
        s_acc[N + col0 + tid] += a1;
      }
    }
  }
  if (do_stats) {
    __syncthreads();
    if (p.tail.kind != 0) {
      // BatchNorm finalize by the last CTA.  One tile per CTA (the usual case: a few hundred rows): the slot is the ROW tile and
      // a CTA writes its own BN columns only, so the tail sums tiles_m slots instead of one per CTA.
      const int ntile = tiles_m * tiles_n;
      const bool compact = ntile <= (int)gridDim.x;
      if (compact) {
        if ((int)blockIdx.x < ntile) {
          const int tm = blockIdx.x / tiles_n, col0 = (blockIdx.x % tiles_n) * SK_BN;
          if (tid < 2 * SK_BN) {
            const int c = col0 + (tid % SK_BN);
            if (c < N) p.stats[(long long)tm * 2 * N + (tid < SK_BN ? 0 : N) + c] = s_acc[(tid < SK_BN ? 0 : N) + c];
          }
        }
      } else {
        for (int c = tid; c < 2 * N; c += 256) p.stats[(long long)blockIdx.x * 2 * N + c] = s_acc[c];
      }
      // scratch: the cp.async ring (every copy has landed and been consumed)
      bnt_run(p.tail, p.stats, N, compact ? tiles_m : (int)gridDim.x, gridDim.x, reinterpret_cast<double*>(sk_smem), SK_STAGES * STAGE_FLOATS / 2);
    } else {
      for (int slot = blockIdx.x; slot < GADDPG_STAT_SLOTS; slot += gridDim.x) {
        const bool mine = (slot == (int)blockIdx.x);
        for (int c = tid; c < 2 * N; c += 256) p.stats[(long long)slot * 2 * N + c] = mine ? s_acc[c] : 0.f;
      }
    }
  }
}

template <int AMODE, int EMODE, int BN>
int launch_skinny_bn(const NTGroup& g, int nprob, int maxM, int maxN, cudaStream_t st) {
  constexpr int STAGE_FLOATS = SK_A_FLOATS * (AMODE == OP_BNBWD ? 2 : 1) + BN * SK_LD;
  const size_t smem = (size_t)SK_STAGES * STAGE_FLOATS * sizeof(float);
  auto kern = skinny_nt_kernel<AMODE, EMODE, BN>;
  GADDPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = ceil_div(maxM, SK_BM) * ceil_div(maxN, BN);
  dim3 grid(tiles < GADDPG_STAT_SLOTS ? tiles : GADDPG_STAT_SLOTS, nprob);
  kern<<<grid, 256, smem, st>>>(g);
  GADDPG_CHECK_LAUNCH("skinny_nt_kernel");
  return GADDPG_OK;
}

template <int AMODE, int EMODE>
int launch_skinny(const NTGroup& g, int nprob, int maxM, int maxN, cudaStream_t st) {
  // 32 x 64 tiles unless that leaves most of the 148 SMs without a CTA
  if (ceil_div(maxM, SK_BM) * ceil_div(maxN, 64) * nprob < gaddpg_sm_count() && maxN > 32)
    return launch_skinny_bn<AMODE, EMODE, 32>(g, nprob, maxM, maxN, st);
  return launch_skinny_bn<AMODE, EMODE, 64>(g, nprob, maxM, maxN, st);
}

}  // namespace

// Problems the skinny kernel takes: every problem of the group has at most 1024 rows (the FFMA kernel's own small-tile
// rule) and 16-byte aligned operand rows.
bool gaddpg_skinny_supported(const NTGroup& g, int nprob, int amode, int emode) {
  for (int i = 0; i < nprob; ++i) {
    const NTProblem& p = g.p[i];
    if (p.M_max > 1024) return false;
    if (p.K % 4 != 0 || p.ldb % 4 != 0 || p.A.ldx % 4 != 0 || ((uintptr_t)p.A.X & 15u) || ((uintptr_t)p.Bw & 15u)) return false;
    if (amode == OP_BNBWD && (p.A.ldy % 4 != 0 || ((uintptr_t)p.A.Y & 15u))) return false;
    if (amode != OP_PLAIN && (((uintptr_t)p.A.c0 | (uintptr_t)p.A.c1) & 15u)) return false;
    if (amode == OP_BNBWD && (((uintptr_t)p.A.c2 | (uintptr_t)p.A.c3 | (uintptr_t)p.A.c4) & 15u)) return false;
  }
  (void)emode;
  return true;
}

int gaddpg_skinny_nt_impl(const NTGroup* g, int nprob, int amode, int emode, void* stream) {
  int maxM = 0, maxN = 0;
  for (int i = 0; i < nprob; ++i) {
    maxM = g->p[i].M_max > maxM ? g->p[i].M_max : maxM;
    maxN = g->p[i].N > maxN ? g->p[i].N : maxN;
  }
  if (maxM == 0) return GADDPG_OK;
  cudaStream_t st = (cudaStream_t)stream;
#define SK_CASE(A, E) \
  if (amode == A && emode == E) return launch_skinny<A, E>(*g, nprob, maxM, maxN, st)
  SK_CASE(OP_PLAIN, EPI_STORE);
  SK_CASE(OP_BNRELU, EPI_STORE);
  SK_CASE(OP_PLAIN, EPI_DMASK);
  SK_CASE(OP_BNBWD, EPI_DMASK);
  SK_CASE(OP_BNBWD, EPI_STORE);
#undef SK_CASE
  gaddpg_set_error("skinny_nt: unsupported mode pair (%d,%d)", amode, emode);
  return GADDPG_ERR_UNSUPPORTED;
}
