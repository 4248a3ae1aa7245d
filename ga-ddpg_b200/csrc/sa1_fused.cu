// sa1_fused.cu — SA1 shared MLP as ONE TMA-fed, recompute-instead-of-store tcgen05 chain (sm_100a only).
//
// Replaces, for the first set-abstraction level (~423 k compact rows at B = 256), what upstream QueryAndGroup + build_shared_mlp
// (Conv2d 1x1 + train-mode BatchNorm2d + ReLU, three times) + F.max_pool2d do on a materialised (B, C, 32, 64) tensor
// (/root/reference/core/networks.py:66-71), and what this library did before with five launches
// that round-tripped every pre-BatchNorm activation through HBM (sa1_l1_fwd -> Y0, gemm_nt -> Y1, gemm_nt -> Y2, pool_fwd).
//
// Train-mode BatchNorm makes every layer a grid-wide dependency (layer l+1 needs the batch statistics of layer l), so the
// chain runs as THREE PHASES — three launches of the same kernel template — and each phase RECOMPUTES the layers below it
// from the L2-resident cloud instead of reading them back:
//   phase 1   gather -> conv0                                    -> statistics of Y0                 [keep: Y0 stored]
//   phase 2   gather -> conv0 -> BN0+ReLU -> conv1               -> statistics of Y1                 [keep: Y1 stored]
//   phase 3   gather -> conv0 -> BN0+ReLU -> conv1 -> BN1+ReLU -> conv2 -> statistics of Y2 and the per-(ball, channel)
//             extreme of Y2 (max for gamma >= 0, min otherwise: BN+ReLU is monotone per channel)   [keep: Y2 stored]
// Forward-only passes (target chain F2/F3, select_action) therefore move NO activation through HBM: algorithmic traffic is
// the row table (12 B/row), the pooled extremes and the slots.  Passes that are differentiated ("keep") store the three
// pre-BN outputs for the backward kernels through TMA (cp.async.bulk.tensor stores from a swizzled staging tile) but never
// read them back in the forward.
//
// One CTA per SM, 128-row tiles, a contiguous tile range per CTA (segments = ball groups are contiguous in the row table, so
// a group is reduced inside one CTA's registers; only the <= 147 groups straddling a CTA boundary are merged by the finalize
// kernel — no atomics, deterministic).  Warp roles (17 warps):
//   P   warps 0-3    thread = row: row-table + scattered cloud reads (3+Cp channels, L2), centroid subtraction, broadcast
//                    (action) channels, hi/lo split, swizzled STS of the [128 x 16] layer-0 operand
//   MMA warp 4       one thread: TMA loads of the pre-split weights (cp.async.bulk.tensor, SWIZZLE_64B/128B, once per CTA),
//                    tcgen05.mma kind::tf32 x3 (3xTF32) for conv0 (K=16), conv1 (K=64), conv2 (K=64, N=128) into TMEM
//   E   warps 5-8    thread = row: tcgen05.ld of conv0's accumulator, BN0+ReLU, hi/lo split, STS into the shared operand buffer
//                    (conv1's operand); then the same for conv1's accumulator (conv2's operand).  One group serves both stages:
//                    the operand of conv1(t+1) may only be written once conv2(t) has read the buffer, i.e. after this group's
//                    work on conv1(t) anyway.
//   G0  warps 9-12   consumers of the phase's LAST accumulator, alternating tiles (even / odd = the two TMEM buffers):
//   G1  warps 13-16  phases 1/2: thread = row -> swizzled staging tile -> column sums; phase 3: conv2 is issued TRANSPOSED
//                    (D[channel][row] = W2 . act^T), so a thread owns a channel and walks the tile's 128 rows straight out of
//                    tcgen05.ld registers: statistics and the running per-group extreme, flushed when the ball group changes
// The conv1 and conv2 operands share ONE 64 KB buffer (conv2's operand is produced from conv1's finished accumulator), which
// is what lets all three weight matrices (hi + lo, 104 KB), the operands and the staging tile fit in 227 KB.
#include "tc_common.cuh"
#include "tma.cuh"
#include "impl.h"

namespace {

constexpr int F_BM = 128;   // rows per tile (UMMA M)
constexpr int F_K0 = 16;    // layer-0 K, padded (3 + per-point channels + broadcast channels <= 16)
constexpr int F_C = 64;     // widths of conv0 / conv1 outputs (networks.py:70: mlp = [in, 64, 64, 128])
constexpr int F_C3 = 128;   // width of conv2's output
constexpr int F_WARPS = 17;
constexpr int F_THREADS = F_WARPS * 32;
constexpr int W_MMA = 4, W_G = 9;   // P: warps 0-3, E: warps 5-8, G0 / G1: warps 9-12 / 13-16

struct FLayout {
  uint32_t w0h, w0l, w1h, w1l, w2h, w2l, a0h[2], a0l[2], abh, abl, stage, consts, meta, bars, red, total;
};
__host__ __device__ constexpr int f_na0(int phase) { return phase == 3 ? 1 : 2; }   // layer-0 operand stages (smem is full in phase 3)
__host__ __device__ inline FLayout f_layout(int phase) {
  FLayout L;
  uint32_t off = 0;
  L.w0h = off; off += F_C * F_K0 * 4;
  L.w0l = off; off += F_C * F_K0 * 4;
  L.w1h = off; off += (phase >= 2) ? F_C * F_C * 4 : 0;
  L.w1l = off; off += (phase >= 2) ? F_C * F_C * 4 : 0;
  L.w2h = off; off += (phase >= 3) ? F_C3 * F_C * 4 : 0;
  L.w2l = off; off += (phase >= 3) ? F_C3 * F_C * 4 : 0;
  for (int i = 0; i < 2; ++i) {
    L.a0h[i] = off; off += (i < f_na0(phase)) ? F_BM * F_K0 * 4 : 0;
    L.a0l[i] = off; off += (i < f_na0(phase)) ? F_BM * F_K0 * 4 : 0;
  }
  L.abh = off; off += (phase >= 2) ? F_BM * F_C * 4 : 0;
  L.abl = off; off += (phase >= 2) ? F_BM * F_C * 4 : 0;
  L.stage = off; off += (phase <= 2) ? 65536 : 32768;   // phases 1/2: 8 warps x [2 col blocks][32 rows x 128 B]; phase 3 (keep): 8 warps x [32 rows x 128 B]
  L.consts = off; off += 2048;          // sc0, sh0, sc1, sh1 [64 each], sign(gamma2) [128]
  L.meta = off; off += (phase >= 3) ? 2 * 1536 : 0;     // phase 3: per consumer group: segment id, weight and run destination of its tile's rows
  L.red = off; off += 2 * 8 * F_C * 4;  // cross-warp combine of the column sums: phases 1/2 [2][8 warps][64]; phase 3 [2 groups][2][128]
  L.bars = off; off += 256;
  L.total = off;
  return L;
}

struct Sa1FParams {
  const float* cloud; long long cloud_sb; int cloud_sc; int skip; int Cp;
  const float* bc; int Cb;
  const float* ctr; int npoint;
  const int32_t* seg_off; const int32_t* row_seg; const int32_t* row_src; const float* row_w;
  int M_max; const int* M_dev;
  const float* sc0; const float* sh0; const float* sc1; const float* sh1; const float* gamma2;
  float* stats;
  float* ext; int32_t* arg; float* part_ext; int32_t* part_arg; int32_t* seg_part;
  int keep;
};

template <int ID>
__device__ __forceinline__ void named_bar(int nthreads) { asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(nthreads) : "memory"); }

// K-major SWIZZLE_64B descriptor (layer-0 operands: rows of 16 floats = 64 B; 8-row atoms of 512 B)
__device__ __forceinline__ uint64_t make_desc64(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (32ull << 32) | (1ull << 46) | (4ull << 61);
}
__device__ __forceinline__ uint32_t sw64_off(int r, int chunk) {  // 16-byte chunk 0..3 of row r
  return (uint32_t)(r * 64 + ((chunk ^ ((r >> 1) & 3)) << 4));
}
__device__ __forceinline__ uint32_t idesc_tf32(int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(F_BM >> 4) << 24);
}

// relu(y * sc + sh) of 4 consecutive columns, split and stored into the K-major SWIZZLE_128B operand buffer
__device__ __forceinline__ void bnrelu_split_store(unsigned char* hi, unsigned char* lo, int row, int col, const float* r4,
                                                   const float* sc, const float* sh) {
  const float4 s = *reinterpret_cast<const float4*>(sc + col), t = *reinterpret_cast<const float4*>(sh + col);
  float4 v;
  v.x = fmaxf(fmaf(r4[0], s.x, t.x), 0.f);
  v.y = fmaxf(fmaf(r4[1], s.y, t.y), 0.f);
  v.z = fmaxf(fmaf(r4[2], s.z, t.z), 0.f);
  v.w = fmaxf(fmaf(r4[3], s.w, t.w), 0.f);
  split_store(hi, lo, sw128_off(row, col, F_BM), v);
}

// Phases 1/2, thread = row with the 64 pre-BN outputs of its row in registers: weighted column sums of the warp's 32 rows
// through a swizzled [2 col blocks][32 rows x 128 B] staging tile (the TMA store box of the keep mode), lane = column pair.
__device__ __forceinline__ void stats64_tile(const float* r, float w, unsigned char* st, int lane, bool keep, const CUtensorMap* tmY,
                                             int grow0, float (&S0)[2], float (&S1)[2]) {
  if (keep && lane == 0) tma_wait_read0();   // the previous tile's store must have read the staging tile
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 16; ++j)
    *reinterpret_cast<float4*>(st + (j >> 3) * 4096 + lane * 128 + (((j & 7) ^ (lane & 7)) << 4)) =
        make_float4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
  if (keep) fence_proxy_async();   // writers: make the generic-proxy stores visible to the TMA (async proxy) before the sync
  __syncwarp();
  if (keep && lane == 0) {
    tma_store_2d(tmY, smem_u32(st), 0, grow0);
    tma_store_2d(tmY, smem_u32(st + 4096), 32, grow0);
    tma_commit();
  }
  const int cb = lane >> 4, col = (2 * lane) & 31, chunk = col >> 2;
  float t00 = 0.f, t01 = 0.f, t10 = 0.f, t11 = 0.f;
#pragma unroll 8
  for (int rr = 0; rr < 32; ++rr) {
    const float2 y = *reinterpret_cast<const float2*>(st + cb * 4096 + rr * 128 + ((chunk ^ (rr & 7)) << 4) + (col & 3) * 4);
    const float wr = __shfl_sync(0xffffffffu, w, rr);
    t00 = fmaf(wr, y.x, t00);
    t01 = fmaf(wr * y.x, y.x, t01);
    t10 = fmaf(wr, y.y, t10);
    t11 = fmaf(wr * y.y, y.y, t11);
  }
  S0[0] += t00; S1[0] += t01; S0[1] += t10; S1[1] += t11;
}

template <int PHASE>
__global__ void __launch_bounds__(F_THREADS, 1)
sa1_fused_kernel(const Sa1FParams p, const __grid_constant__ CUtensorMap tW0h, const __grid_constant__ CUtensorMap tW0l,
                 const __grid_constant__ CUtensorMap tW1h, const __grid_constant__ CUtensorMap tW1l,
                 const __grid_constant__ CUtensorMap tW2h, const __grid_constant__ CUtensorMap tW2l,
                 const __grid_constant__ CUtensorMap tY) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const FLayout L = f_layout(PHASE);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* w_full = bars + 0;
  uint64_t* a0_full = bars + 1;      // [2]
  uint64_t* a0_empty = bars + 3;     // [2]
  uint64_t* acc0_full = bars + 5;    // [2]
  uint64_t* acc0_empty = bars + 7;   // [2]
  uint64_t* ab1_full = bars + 9;
  uint64_t* ab_free = bars + 10;
  uint64_t* acc1_full = bars + 11;   // [2]
  uint64_t* acc1_empty = bars + 13;  // [2]
  uint64_t* ab2_full = bars + 15;
  uint64_t* acc2_full = bars + 16;   // [2]
  uint64_t* acc2_empty = bars + 18;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  constexpr int NA0 = f_na0(PHASE);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int M = p.M_dev ? *p.M_dev : p.M_max;
  M = M < p.M_max ? M : p.M_max;
  const int ntiles = (M + F_BM - 1) / F_BM;
  const int tpc = (ntiles + (int)gridDim.x - 1) / (int)gridDim.x;   // contiguous tile range per CTA
  const int tile0 = (int)blockIdx.x * tpc;
  int my_tiles = ntiles - tile0;
  my_tiles = my_tiles < 0 ? 0 : (my_tiles > tpc ? tpc : my_tiles);
  constexpr uint32_t TMEM_COLS = PHASE == 1 ? 128 : PHASE == 2 ? 256 : 512;
  constexpr uint32_t ACC0 = 0, ACC1 = 128, ACC2 = 256;

  if (tid == 0) {
    mbar_init(w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&a0_full[s], 128);
      mbar_init(&a0_empty[s], 1);
      mbar_init(&acc0_full[s], 1);
      mbar_init(&acc0_empty[s], 4);
      mbar_init(&acc1_full[s], 1);
      mbar_init(&acc1_empty[s], 4);
      mbar_init(&acc2_full[s], 1);
      mbar_init(&acc2_empty[s], 4);
    }
    mbar_init(ab1_full, 128);
    mbar_init(ab_free, 1);
    mbar_init(ab2_full, 128);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, TMEM_COLS);
  {
    float* cs = reinterpret_cast<float*>(smem + L.consts);
    if (PHASE >= 2)
      for (int i = tid; i < F_C; i += F_THREADS) {
        cs[i] = p.sc0[i];
        cs[64 + i] = p.sh0[i];
      }
    if (PHASE >= 3)
      for (int i = tid; i < F_C; i += F_THREADS) {
        cs[128 + i] = p.sc1[i];
        cs[192 + i] = p.sh1[i];
      }
    if (PHASE >= 3)
      for (int i = tid; i < F_C3; i += F_THREADS) cs[256 + i] = p.gamma2[i] < 0.f ? -1.f : 1.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t sbase = smem_u32(smem);

  if (warp < W_MMA) {
    // ===================== P: gather + layer-0 operand =====================
    // Three-deep software pipeline (the row -> (segment, source point) -> cloud reads are two dependent L2 round trips per
    // tile and only 128 threads issue them): indices of tile it+3, data of tiles it+1 and it+2 are in flight while tile it
    // is split and stored.
    const int r = tid;  // tile row
    const int K1 = 3 + p.Cp + p.Cb;
    struct Idx { int seg, src; };
    auto load_idx = [&](int it) {
      Idx x;
      x.seg = -1;
      x.src = 0;
      const int row = (tile0 + it) * F_BM + r;
      if (it < my_tiles && row < M) {
        x.seg = p.row_seg[row];
        x.src = p.row_src[row];
      }
      return x;
    };
    auto load_data = [&](const Idx& x, float (&in)[F_K0]) {
#pragma unroll
      for (int k = 0; k < F_K0; ++k) in[k] = 0.f;
      if (x.seg >= 0) {
        const int b = x.seg / p.npoint;
        const float* pc = p.cloud + (long long)b * p.cloud_sb + p.skip + x.src;
#pragma unroll
        for (int k = 0; k < F_K0 - 3; ++k)
          if (k < p.Cp) in[3 + k] = pc[(long long)k * p.cloud_sc];
        in[0] = p.ctr[(long long)x.seg * 3 + 0];   // centroid: subtracted when the tile is stored
        in[1] = p.ctr[(long long)x.seg * 3 + 1];
        in[2] = p.ctr[(long long)x.seg * 3 + 2];
#pragma unroll
        for (int k = 3; k < F_K0; ++k) {
          const int j = k - 3 - p.Cp;
          if (j >= 0 && k < K1) in[k] = p.bc[(long long)b * p.Cb + j];
        }
      }
    };
    float d0[F_K0], d1[F_K0], d2[F_K0];
    Idx i1 = load_idx(1), i2 = load_idx(2);
    {
      Idx i0 = load_idx(0);
      load_data(i0, d0);
    }
    load_data(i1, d1);
    for (int it = 0; it < my_tiles; ++it) {
      Idx i3 = load_idx(it + 3);
      load_data(i2, d2);                      // tile it+2
      const int sa = it % NA0;
      mbar_wait(&a0_empty[sa], ((uint32_t)(it / NA0) & 1u) ^ 1u);
      d0[0] = d0[3] - d0[0];                  // dxyz = xyz[src] - centroid (zero rows stay zero)
      d0[1] = d0[4] - d0[1];
      d0[2] = d0[5] - d0[2];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        split_store(smem + (sa ? L.a0h[1] : L.a0h[0]), smem + (sa ? L.a0l[1] : L.a0l[0]), sw64_off(r, c),
                    make_float4(d0[4 * c], d0[4 * c + 1], d0[4 * c + 2], d0[4 * c + 3]));
      fence_proxy_async();
      mbar_arrive(&a0_full[sa]);
#pragma unroll
      for (int k = 0; k < F_K0; ++k) {
        d0[k] = d1[k];
        d1[k] = d2[k];
      }
      i2 = i3;
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer (+ TMA weight loads) =====================
    if (lane == 0) {
      constexpr uint32_t WBYTES = 2 * F_C * F_K0 * 4 + (PHASE >= 2 ? 2 * F_C * F_C * 4 : 0) + (PHASE >= 3 ? 2 * F_C3 * F_C * 4 : 0);
      mbar_expect_tx(w_full, WBYTES);
      tma_load_2d(sbase + L.w0h, &tW0h, 0, 0, w_full);
      tma_load_2d(sbase + L.w0l, &tW0l, 0, 0, w_full);
      if (PHASE >= 2) {
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_2d(sbase + L.w1h + kb * F_C * 128, &tW1h, kb * 32, 0, w_full);
          tma_load_2d(sbase + L.w1l + kb * F_C * 128, &tW1l, kb * 32, 0, w_full);
        }
      }
      if (PHASE >= 3) {
        for (int kb = 0; kb < 2; ++kb) {
          tma_load_2d(sbase + L.w2h + kb * F_C3 * 128, &tW2h, kb * 32, 0, w_full);
          tma_load_2d(sbase + L.w2l + kb * F_C3 * 128, &tW2l, kb * 32, 0, w_full);
        }
      }
      mbar_wait(w_full, 0);
      tc_fence_after();
      const uint32_t id64 = idesc_tf32(F_C), id128 = idesc_tf32(F_C3);
      auto mma0 = [&](int j) {
        const int b = j & 1, sa = j % NA0;
        mbar_wait(&a0_full[sa], (uint32_t)(j / NA0) & 1u);
        mbar_wait(&acc0_empty[b], (((uint32_t)j >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + ACC0 + (uint32_t)b * F_C;
        uint32_t acc = 0;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          const uint64_t ah = make_desc64(sbase + (sa ? L.a0h[1] : L.a0h[0]) + ks * 32), al = make_desc64(sbase + (sa ? L.a0l[1] : L.a0l[0]) + ks * 32);
          const uint64_t bh = make_desc64(sbase + L.w0h + ks * 32), bl = make_desc64(sbase + L.w0l + ks * 32);
          umma_tf32(d, al, bh, id64, acc);
          umma_tf32(d, ah, bl, id64, 1u);
          umma_tf32(d, ah, bh, id64, 1u);
          acc = 1u;
        }
        umma_commit(&a0_empty[sa]);
        umma_commit(&acc0_full[b]);
      };
      // D[rows x channels] = act . W^T, or (transposed = true) D[channels x rows] = W . act^T: the accumulator then has one
      // CHANNEL per TMEM lane and the tile's 128 rows along its columns, which is the layout conv2's consumer wants (a thread
      // walks the rows of its channel straight out of tcgen05.ld registers: statistics + per-group extremes, no transpose)
      auto mma_k64 = [&](uint32_t d, uint32_t wh, uint32_t wl, int nrows_w, uint32_t idesc, bool transposed) {
        uint32_t acc = 0;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t xoff = (uint32_t)(kb * F_BM * 128 + ks * 32), woff = (uint32_t)(kb * nrows_w * 128 + ks * 32);
            const uint64_t xh = make_desc(sbase + L.abh + xoff), xl = make_desc(sbase + L.abl + xoff);
            const uint64_t wdh = make_desc(sbase + wh + woff), wdl = make_desc(sbase + wl + woff);
            if (transposed) {
              umma_tf32(d, wdl, xh, idesc, acc);
              umma_tf32(d, wdh, xl, idesc, 1u);
              umma_tf32(d, wdh, xh, idesc, 1u);
            } else {
              umma_tf32(d, xl, wdh, idesc, acc);
              umma_tf32(d, xh, wdl, idesc, 1u);
              umma_tf32(d, xh, wdh, idesc, 1u);
            }
            acc = 1u;
          }
        }
      };
      if (my_tiles > 0) mma0(0);
      for (int it = 0; it < my_tiles; ++it) {
        if (it + 1 < my_tiles) mma0(it + 1);   // conv0 of the next tile first: E0 can run ahead of conv1/conv2 of this one
        if (PHASE >= 2) {
          const int b = it & 1;
          mbar_wait(ab1_full, (uint32_t)it & 1u);
          mbar_wait(&acc1_empty[b], (((uint32_t)it >> 1) & 1u) ^ 1u);
          tc_fence_after();
          mma_k64(tmem_base + ACC1 + (uint32_t)b * F_C, L.w1h, L.w1l, F_C, id64, false);
          umma_commit(&acc1_full[b]);
          if (PHASE == 2) umma_commit(ab_free);
        }
        if (PHASE >= 3) {
          const int b = it & 1;
          mbar_wait(ab2_full, (uint32_t)it & 1u);
          mbar_wait(&acc2_empty[b], (((uint32_t)it >> 1) & 1u) ^ 1u);
          tc_fence_after();
          mma_k64(tmem_base + ACC2 + (uint32_t)b * F_BM, L.w2h, L.w2l, F_C3, id128, true);
          umma_commit(&acc2_full[b]);
          umma_commit(ab_free);
        }
      }
    }
    __syncwarp();
  } else if (warp < W_G) {
    // ===================== E: conv0 / conv1 accumulators -> operand of the next conv =====================
    if (PHASE >= 2) {
      const int q = warp & 3, row = q * 32 + lane;
      const float* cs = reinterpret_cast<const float*>(smem + L.consts);
      auto stage0 = [&](int j) {   // conv0 accumulator of tile j -> relu(bn0) -> conv1 operand
        const int b = j & 1;
        mbar_wait(&acc0_full[b], ((uint32_t)j >> 1) & 1u);
        tc_fence_after();
        float r[F_C];
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + ACC0 + (uint32_t)b * F_C;
        tmem_ld32_async(ta, r);
        tmem_ld32_async(ta + 32, r + 32);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc0_empty[b]);
        mbar_wait(ab_free, ((uint32_t)j & 1u) ^ 1u);   // conv2(j-1) (phase 2: conv1(j-1)) has read the operand buffer
#pragma unroll
        for (int k = 0; k < 16; ++k) bnrelu_split_store(smem + L.abh, smem + L.abl, row, 4 * k, r + 4 * k, cs, cs + 64);
        fence_proxy_async();
        mbar_arrive(ab1_full);
      };
      if (my_tiles > 0) stage0(0);
      for (int it = 0; it < my_tiles; ++it) {
        if (PHASE >= 3) {   // conv1 accumulator -> relu(bn1) -> conv2 operand (conv1(it) is complete: the buffer may be overwritten)
          const int b = it & 1;
          mbar_wait(&acc1_full[b], ((uint32_t)it >> 1) & 1u);
          tc_fence_after();
          float r[F_C];
          const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + ACC1 + (uint32_t)b * F_C;
          tmem_ld32_async(ta, r);
          tmem_ld32_async(ta + 32, r + 32);
          tmem_wait_ld();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc1_empty[b]);
#pragma unroll
          for (int k = 0; k < 16; ++k) bnrelu_split_store(smem + L.abh, smem + L.abl, row, 4 * k, r + 4 * k, cs + 128, cs + 192);
          fence_proxy_async();
          mbar_arrive(ab2_full);
        }
        if (it + 1 < my_tiles) stage0(it + 1);   // its TMEM read + math overlap conv2(it); only the stores wait for it
      }
    }
  } else {
    // ===================== G0 / G1: consumers of the phase's last accumulator, alternating tiles =====================
    const int grp = (warp - W_G) >> 2;   // 0: even tiles (TMEM buffer 0), 1: odd tiles
    const int q = warp & 3;              // TMEM lane quarter
    const int w8 = warp - W_G;           // 0..7
    if (PHASE <= 2) {
      // thread = row: weighted column sums of the pre-BN output (conv0 in phase 1, conv1 in phase 2) [+ TMA store in keep mode]
      const int row = q * 32 + lane;
      unsigned char* st = smem + L.stage + w8 * 8192;
      uint64_t* full = PHASE == 1 ? acc0_full : acc1_full;
      uint64_t* empty = PHASE == 1 ? acc0_empty : acc1_empty;
      constexpr uint32_t ACC = PHASE == 1 ? ACC0 : ACC1;
      float S0[2] = {0.f, 0.f}, S1[2] = {0.f, 0.f};
      for (int it = grp; it < my_tiles; it += 2) {
        const int grow = (tile0 + it) * F_BM + row;
        const float w = grow < M ? p.row_w[grow] : 0.f;
        mbar_wait(&full[grp], ((uint32_t)it >> 1) & 1u);
        tc_fence_after();
        float r[F_C];
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + ACC + (uint32_t)grp * F_C;
        tmem_ld32_async(ta, r);
        tmem_ld32_async(ta + 32, r + 32);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[grp]);
        stats64_tile(r, w, st, lane, p.keep != 0, &tY, (tile0 + it) * F_BM + q * 32, S0, S1);
      }
      if (p.keep && lane == 0) tma_wait_all0();
      float* red = reinterpret_cast<float*>(smem + L.red);
      red[w8 * 64 + 2 * lane] = S0[0];
      red[w8 * 64 + 2 * lane + 1] = S0[1];
      red[512 + w8 * 64 + 2 * lane] = S1[0];
      red[512 + w8 * 64 + 2 * lane + 1] = S1[1];
    } else {
      // thread = channel (conv2 was issued transposed): statistics + per-group extremes of the tile's 128 rows from registers.
      // Everything that depends only on the ROW (run starts, statistic weights, where a finished run is written) is worked
      // out once per tile by the group into shared memory; the per-row code of a thread is branch-free: the flush of a
      // finished run is two predicated stores.
      const int c = q * 32 + lane;
      const int gt = (warp - W_G - grp * 4) * 32 + lane;   // thread index inside the group (row whose meta it prepares)
      const float* cs = reinterpret_cast<const float*>(smem + L.consts);
      const float sgn = cs[256 + c];      // +1: the pooled value is bn(max y); -1: bn(min y)
      int32_t* mseg = reinterpret_cast<int32_t*>(smem + L.meta + grp * 1536);        // [128] segment id (-1: row >= M)
      float* mw = reinterpret_cast<float*>(smem + L.meta + grp * 1536 + 512);         // [128] statistic weight of the row
      int32_t* rdst = reinterpret_cast<int32_t*>(smem + L.meta + grp * 1536 + 1024);  // [128] destination of run k: element
                                                                                     // offset; bit 30: per-tile side table
      unsigned char* st = smem + L.stage + w8 * 4096;   // keep: [32 rows x 128 B] of this warp's 32 channels
      float S0 = 0.f, S1 = 0.f;
      for (int it = grp; it < my_tiles; it += 2) {
        const int tile = tile0 + it, trow0 = tile * F_BM;
        if (grp == 0) named_bar<1>(128); else named_bar<2>(128);    // the group is done with the previous tile's meta
        {
          const int grow = trow0 + gt;
          mseg[gt] = grow < M ? p.row_seg[grow] : -1;
          mw[gt] = grow < M ? p.row_w[grow] : 0.f;
        }
        if (grp == 0) named_bar<1>(128); else named_bar<2>(128);
        uint32_t m0, m1, m2, m3;   // bit rr of mk: row 32k+rr starts a run (a new ball group, the tile, or the dead tail)
        m0 = __ballot_sync(0xffffffffu, lane == 0 || mseg[lane] != mseg[lane - 1]);
        m1 = __ballot_sync(0xffffffffu, mseg[32 + lane] != mseg[31 + lane]);
        m2 = __ballot_sync(0xffffffffu, mseg[64 + lane] != mseg[63 + lane]);
        m3 = __ballot_sync(0xffffffffu, mseg[96 + lane] != mseg[95 + lane]);
        {   // run k (k-th set bit) -> where its result goes
          const int sg = mseg[gt];
          const bool start = gt == 0 || sg != mseg[gt - 1];
          if (start) {
            const int wq = gt >> 5;
            const uint32_t below = (wq == 0 ? m0 : wq == 1 ? m1 : wq == 2 ? m2 : m3) & ((1u << (gt & 31)) - 1u);
            const int k = __popc(below) + (wq > 0 ? __popc(m0) : 0) + (wq > 1 ? __popc(m1) : 0) + (wq > 2 ? __popc(m2) : 0);
            int d;
            if (sg < 0) {
              d = (1 << 30) | (((p.M_max + F_BM - 1) / F_BM) * F_C3);   // dead tail of the last tile: parked in the spare side-table row
            } else if (gt == 0 && p.seg_off[sg] < trow0) {
              d = (1 << 30) | (tile * F_C3);     // the group began in the previous tile: partial, merged by the finalize kernel
              p.seg_part[sg] = tile;
            } else {
              d = sg * F_C3;
            }
            rdst[k] = d;
          }
        }
        if (grp == 0) named_bar<1>(128); else named_bar<2>(128);
        float best = -3.4e38f, t0 = 0.f, t1 = 0.f;
        int barg = 0, krun = -1;
        mbar_wait(&acc2_full[grp], ((uint32_t)it >> 1) & 1u);
        tc_fence_after();
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + ACC2 + (uint32_t)grp * F_BM;
        float va[32], vb[32];
        auto emit = [&](uint32_t pred, int k) {   // predicated: store the finished run k
          const int d = rdst[k < 0 ? 0 : k];
          float* pe = ((d >> 30) & 1) ? p.part_ext : p.ext;
          int32_t* pa = ((d >> 30) & 1) ? p.part_arg : p.arg;
          const long long off = (long long)(d & 0x3FFFFFFF) + c;
          asm volatile(
              "{\n\t.reg .pred q;\n\t"
              "setp.ne.u32 q, %0, 0;\n\t"
              "@q st.global.f32 [%1], %2;\n\t"
              "@q st.global.s32 [%3], %4;\n\t}" ::"r"(pred), "l"(pe + off), "f"(best * sgn), "l"(pa + off), "r"(barg)
              : "memory");
        };
        auto chunk = [&](const float (&v)[32], int k, uint32_t m) {   // rows 32k .. 32k+31 of this thread's channel
          const int base = k * 32;
#pragma unroll
          for (int rr = 0; rr < 32; ++rr) {
            const uint32_t start = (m >> rr) & 1u;              // warp-uniform
            emit(start & (krun >= 0 ? 1u : 0u), krun);
            krun += (int)start;
            best = start ? -3.4e38f : best;
            const float w = mw[base + rr];                      // broadcast
            t0 = fmaf(w, v[rr], t0);
            t1 = fmaf(w * v[rr], v[rr], t1);
            const float key = v[rr] * sgn;
            const bool better = key > best;
            best = better ? key : best;
            barg = better ? trow0 + base + rr : barg;
          }
          if (p.keep) {   // this warp's [32 rows x 32 channels] block of the Y2 tile, swizzled for the TMA store
            if (lane == 0) tma_wait_read0();
            __syncwarp();
#pragma unroll
            for (int rr = 0; rr < 32; ++rr)
              *reinterpret_cast<float*>(st + rr * 128 + (((lane >> 2) ^ (rr & 7)) << 4) + (lane & 3) * 4) = v[rr];
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tY, smem_u32(st), q * 32, trow0 + base);
              tma_commit();
            }
          }
        };
        tmem_ld32_async(ta, va);
        tmem_wait_ld();
        tmem_ld32_async(ta + 32, vb);
        chunk(va, 0, m0);
        tmem_wait_ld();
        tmem_ld32_async(ta + 64, va);
        chunk(vb, 1, m1);
        tmem_wait_ld();
        tmem_ld32_async(ta + 96, vb);
        chunk(va, 2, m2);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&acc2_empty[grp]);
        chunk(vb, 3, m3);
        emit(1u, krun);   // the tile's last run (a complete group, the head of one that continues in the next tile, or the dead tail)
        S0 += t0;
        S1 += t1;
      }
      if (p.keep && lane == 0) tma_wait_all0();
      float* red = reinterpret_cast<float*>(smem + L.red);
      red[grp * 256 + c] = S0;
      red[grp * 256 + 128 + c] = S1;
    }
  }
  tc_fence_before();
  __syncthreads();
  {   // one slot per CTA, folded in a fixed order
    const float* red = reinterpret_cast<const float*>(smem + L.red);
    if (PHASE <= 2) {
      if (tid < 2 * F_C) {
        const int h = tid >> 6, cc = tid & 63;
        float a = 0.f;
        for (int w = 0; w < 8; ++w) {
          const bool used = my_tiles > ((w >> 2) & 1);   // group 1 (warps 4-7) only ran if the CTA had an odd tile
          a += used ? red[h * 512 + w * 64 + cc] : 0.f;
        }
        p.stats[(long long)blockIdx.x * 2 * F_C + tid] = a;
      }
    } else {
      if (tid < 2 * F_C3) {
        const int h = tid >> 7, cc = tid & 127;
        p.stats[(long long)blockIdx.x * 2 * F_C3 + tid] = red[h * 128 + cc] + red[256 + h * 128 + cc];
      }
    }
    constexpr int CW = PHASE <= 2 ? 2 * F_C : 2 * F_C3;
    for (int slot = blockIdx.x + gridDim.x; slot < GADDPG_STAT_SLOTS; slot += gridDim.x)
      for (int i = tid; i < CW; i += F_THREADS) p.stats[(long long)slot * CW + i] = 0.f;
  }
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// out = relu(bn(extreme)) + arg-max row per (segment, channel); merges the segments that straddle two CTAs' tile ranges
__global__ void __launch_bounds__(128) sa1_pool_finalize_kernel(const float* __restrict__ ext, const int32_t* __restrict__ arg,
                                                                const float* __restrict__ part_ext,
                                                                const int32_t* __restrict__ part_arg, int32_t* __restrict__ seg_part,
                                                                const float* __restrict__ gamma, const float* __restrict__ scale,
                                                                const float* __restrict__ shift, int S, float* __restrict__ out,
                                                                int32_t* __restrict__ arg_out) {
  const int c = threadIdx.x;
  const bool neg = gamma[c] < 0.f;
  const float sc = scale[c], sh = shift[c];
  for (int s = blockIdx.x; s < S; s += gridDim.x) {
    float v = ext[(long long)s * F_C3 + c];
    int a = arg[(long long)s * F_C3 + c];
    const int sp = seg_part[s];
    if (sp >= 0) {   // the tail of this group was reduced with tile sp: higher rows, so a tie keeps the head's row
      const float pv = part_ext[(long long)sp * F_C3 + c];
      if ((neg ? -pv : pv) > (neg ? -v : v)) {
        v = pv;
        a = part_arg[(long long)sp * F_C3 + c];
      }
    }
    out[(long long)s * F_C3 + c] = fmaxf(fmaf(v, sc, sh), 0.f);
    if (arg_out) arg_out[(long long)s * F_C3 + c] = a;
    __syncthreads();
    if (c == 0 && sp >= 0) seg_part[s] = -1;
  }
}

// pre-split (hi = tf32 bits, lo = x - hi) weight images for the TMA loads, once per optimiser step
__global__ void sa1f_wprep_kernel(const float* __restrict__ W0, int ld0, int K1, const float* __restrict__ W1, const float* __restrict__ W2,
                                  float* __restrict__ out) {
  const int n0 = F_C * F_K0, n1 = F_C * F_C, n2 = F_C3 * F_C;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n0 + n1 + n2; i += gridDim.x * blockDim.x) {
    float x;
    float* dst;
    if (i < n0) {
      const int n = i / F_K0, k = i % F_K0;
      x = k < K1 ? W0[(long long)n * ld0 + k] : 0.f;
      dst = out + i;
    } else if (i < n0 + n1) {
      x = W1[i - n0];
      dst = out + 2 * n0 + (i - n0);
    } else {
      x = W2[i - n0 - n1];
      dst = out + 2 * n0 + 2 * n1 + (i - n0 - n1);
    }
    const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    const int half = i < n0 ? n0 : (i < n0 + n1 ? n1 : n2);
    dst[0] = h;
    dst[half] = x - h;
  }
}

}  // namespace

int gaddpg_sa1f_wprep_impl(const float* W0, int ld0, int K1, const float* W1, const float* W2, float* wsplit, void* stream) {
  GADDPG_CHECK_ARG(W0 && W1 && W2 && wsplit && K1 >= 4 && K1 <= F_K0 && ld0 >= K1, "sa1f_wprep: bad argument (K1=%d)", K1);
  sa1f_wprep_kernel<<<24, 256, 0, (cudaStream_t)stream>>>(W0, ld0, K1, W1, W2, wsplit);
  GADDPG_CHECK_LAUNCH("sa1f_wprep_kernel");
  return GADDPG_OK;
}

long long gaddpg_sa1f_wsplit_floats_impl() { return 2ll * (F_C * F_K0 + F_C * F_C + F_C3 * F_C); }

int gaddpg_sa1_fused_fwd_impl(int phase, const float* cloud, long long cloud_sb, int cloud_sc, int skip, int Cp, const float* bc, int Cb,
                              const float* ctr, int npoint, const int32_t* seg_off, const int32_t* row_seg, const int32_t* row_src,
                              const float* row_w, int M_max, const int* M_dev, const float* wsplit, const float* sc0, const float* sh0,
                              const float* sc1, const float* sh1, const float* gamma2, float* stats, float* Ykeep, float* ext,
                              int32_t* arg, float* part_ext, int32_t* part_arg, int32_t* seg_part, void* stream) {
  GADDPG_CHECK_ARG(phase >= 1 && phase <= 3, "sa1_fused_fwd: phase must be 1, 2 or 3");
  GADDPG_CHECK_ARG(cloud && ctr && seg_off && row_seg && row_src && row_w && wsplit && stats && M_max >= 0 && npoint >= 1,
                   "sa1_fused_fwd: null pointer");
  GADDPG_CHECK_ARG(Cp >= 3 && Cb >= 0 && 3 + Cp + Cb <= F_K0 && (Cb == 0 || bc), "sa1_fused_fwd: 3 + %d + %d input channels exceed %d", Cp, Cb, F_K0);
  GADDPG_CHECK_ARG(phase < 2 || (sc0 && sh0), "sa1_fused_fwd: phase %d needs the BatchNorm constants of layer 0", phase);
  GADDPG_CHECK_ARG(phase < 3 || (sc1 && sh1 && gamma2 && ext && arg && part_ext && part_arg && seg_part), "sa1_fused_fwd: phase 3 arguments");
  if (M_max == 0) return GADDPG_OK;
  const int n0 = F_C * F_K0, n1 = F_C * F_C, n2 = F_C3 * F_C;
  const float *w0h = wsplit, *w0l = wsplit + n0, *w1h = wsplit + 2 * n0, *w1l = w1h + n1, *w2h = w1h + 2 * n1, *w2l = w2h + n2;
  CUtensorMap m0h, m0l, m1h, m1l, m2h, m2l, mY;
  bool ok = make_map(&m0h, w0h, F_C, F_K0, F_K0, F_C, F_K0, CU_TENSOR_MAP_SWIZZLE_64B) && make_map(&m0l, w0l, F_C, F_K0, F_K0, F_C, F_K0, CU_TENSOR_MAP_SWIZZLE_64B) &&
            make_map(&m1h, w1h, F_C, F_C, F_C, F_C, 32, CU_TENSOR_MAP_SWIZZLE_128B) && make_map(&m1l, w1l, F_C, F_C, F_C, F_C, 32, CU_TENSOR_MAP_SWIZZLE_128B) &&
            make_map(&m2h, w2h, F_C3, F_C, F_C, F_C3, 32, CU_TENSOR_MAP_SWIZZLE_128B) && make_map(&m2l, w2l, F_C3, F_C, F_C, F_C3, 32, CU_TENSOR_MAP_SWIZZLE_128B);
  const int CY = phase == 3 ? F_C3 : F_C;
  if (Ykeep) {
    GADDPG_CHECK_ARG(((uintptr_t)Ykeep & 15u) == 0, "sa1_fused_fwd: Ykeep must be 16-byte aligned");
    ok = ok && make_map(&mY, Ykeep, M_max, CY, CY, 32, 32, CU_TENSOR_MAP_SWIZZLE_128B);
  } else {
    mY = m1h;   // unused
  }
  if (!ok) {
    gaddpg_set_error("sa1_fused_fwd: cuTensorMapEncodeTiled failed or is unavailable");
    return GADDPG_ERR_CUDA;
  }
  Sa1FParams p;
  p.cloud = cloud; p.cloud_sb = cloud_sb; p.cloud_sc = cloud_sc; p.skip = skip; p.Cp = Cp; p.bc = bc; p.Cb = Cb; p.ctr = ctr;
  p.npoint = npoint; p.seg_off = seg_off; p.row_seg = row_seg; p.row_src = row_src; p.row_w = row_w; p.M_max = M_max; p.M_dev = M_dev;
  p.sc0 = sc0; p.sh0 = sh0; p.sc1 = sc1; p.sh1 = sh1; p.gamma2 = gamma2; p.stats = stats; p.ext = ext; p.arg = arg;
  p.part_ext = part_ext; p.part_arg = part_arg; p.seg_part = seg_part; p.keep = Ykeep ? 1 : 0;
  const int tiles = ceil_div(M_max, F_BM);
  const int grid = tiles < gaddpg_sm_count() ? tiles : gaddpg_sm_count();
  const size_t smem = f_layout(phase).total + 1024;
  cudaStream_t st = (cudaStream_t)stream;
#define F_LAUNCH(PH)                                                                                             \
  {                                                                                                              \
    auto kern = sa1_fused_kernel<PH>;                                                                            \
    GADDPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));             \
    kern<<<grid, F_THREADS, smem, st>>>(p, m0h, m0l, m1h, m1l, m2h, m2l, mY);                                    \
    GADDPG_CHECK_LAUNCH("sa1_fused_kernel");                                                                     \
  }
  if (phase == 1) F_LAUNCH(1) else if (phase == 2) F_LAUNCH(2) else F_LAUNCH(3)
#undef F_LAUNCH
  return GADDPG_OK;
}

int gaddpg_sa1_fused_grid_impl(int M_max) {
  const int tiles = ceil_div(M_max, F_BM);
  return tiles < gaddpg_sm_count() ? tiles : gaddpg_sm_count();
}

int gaddpg_sa1_pool_finalize_impl(const float* ext, const int32_t* arg, const float* part_ext, const int32_t* part_arg, int32_t* seg_part,
                                  const float* gamma, const float* scale, const float* shift, int S, float* out, int32_t* arg_out,
                                  void* stream) {
  GADDPG_CHECK_ARG(ext && arg && part_ext && part_arg && seg_part && gamma && scale && shift && out && S >= 1, "sa1_pool_finalize: bad argument");
  const int grid = S < 4 * gaddpg_sm_count() ? S : 4 * gaddpg_sm_count();
  sa1_pool_finalize_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(ext, arg, part_ext, part_arg, seg_part, gamma, scale, shift, S, out, arg_out);
  GADDPG_CHECK_LAUNCH("sa1_pool_finalize_kernel");
  return GADDPG_OK;
}
