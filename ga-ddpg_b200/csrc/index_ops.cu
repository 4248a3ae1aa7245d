// index_ops.cu — farthest-point sampling, ball query, gather/group index kernels (sm_100a).
//
// Replaces the `pointnet2_ops._ext` CUDA ops the reference reaches through
// pointnet2_ops.pointnet2_modules.PointnetSAModule (/root/reference/core/networks.py:66-81)
// and pointnet2_utils.{furthest_point_sample,gather_operation} (/root/reference/core/utils.py:795-796).
// Semantics: SURVEY.md §8 Spec S1 (FPS), S2 (ball query), S3 (grouping).  Results are bit-exact
// against oracle/pointnet2_cpu.c (which simulates the upstream launch shape literally).
//
// Design (B200): one CTA per cloud.  The cloud's xyz is staged ONCE into shared memory as SoA
// (3*N floats: 48 KB at N=4096, 96 KB at N=8192), every thread keeps its P points and their running
// min-distance in registers for all m-1 FPS iterations, the arg-max is a single 64-bit max-reduce on
// the packed key of Spec S1 (warp shuffles + one shared-memory stage, ONE __syncthreads per
// iteration), the gather of the new centroids is free (they are in shared memory), and the ball query
// runs in the same kernel as a warp-per-centroid ordered compaction (ballot + popc) over the staged
// points with coalesced int32 stores.  Nothing but the cloud is read from HBM and nothing but the
// indices, counts and centroids is written.
#include "common.cuh"

namespace {

struct XyzView {
  const float* p;  // element (b,k,c) = p[b*sb + k*sk + c*sc]
  long long sb;
  int sk;
  int sc;
};

__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
  float dx = ax - bx, dy = ay - by, dz = az - bz;
  return __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
}

// Spec S1 tie-break rank of point k under the upstream launch shape (bs virtual threads):
// smaller bit-reversed (k mod bs) first, then smaller (k div bs).
__device__ __forceinline__ unsigned fps_rank(int k, int bs, int log2bs, int per) {
  unsigned r = (unsigned)k & (unsigned)(bs - 1);
  unsigned q = (unsigned)k >> log2bs;
  unsigned br = log2bs ? (__brev(r) >> (32 - log2bs)) : 0u;
  return br * (unsigned)per + q;
}
__device__ __forceinline__ int fps_unrank(unsigned rank, int bs, int log2bs, int per) {
  unsigned br = rank / (unsigned)per, q = rank % (unsigned)per;
  unsigned r = log2bs ? (__brev(br) >> (32 - log2bs)) : 0u;
  return (int)(q * (unsigned)bs + r);
}

// warp-per-centroid ordered ball query over points staged in shared memory (Spec S2)
__device__ __forceinline__ void ball_query_warp(const float* sx, const float* sy, const float* sz, int N,
                                                float cx, float cy, float cz, float radius2, int nsample,
                                                int32_t* __restrict__ out, int32_t* __restrict__ cnt_out,
                                                int lane) {
  int cnt = 0, first = 0;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (int base = 0; base < N && cnt < nsample; base += 32) {
    int k = base + lane;
    bool hit = false;
    if (k < N) hit = sqdist3(cx, cy, cz, sx[k], sy[k], sz[k]) < radius2;
    unsigned mask = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      int pos = cnt + __popc(mask & lt_mask);
      if (pos < nsample) out[pos] = k;
    }
    if (cnt == 0 && mask) first = base + __ffs(mask) - 1;
    cnt += __popc(mask);
  }
  if (cnt > nsample) cnt = nsample;
  for (int l = cnt + lane; l < nsample; l += 32) out[l] = first;  // pad with the first hit (0 if none)
  if (lane == 0 && cnt_out) *cnt_out = cnt;
}

template <int T, int P>
__global__ void __launch_bounds__(T) fps_ballquery_kernel(XyzView xyz, int N, int m, int bs, int log2bs, int per,
                                                          int32_t* __restrict__ fps_idx,
                                                          float* __restrict__ new_xyz, int do_bq, float radius2,
                                                          int nsample, int32_t* __restrict__ bq_idx,
                                                          int32_t* __restrict__ bq_cnt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NW = T / 32;
  unsigned long long* s_slot = reinterpret_cast<unsigned long long*>(smem_raw);  // 2*NW keys (8-byte aligned)
  float* sx = reinterpret_cast<float*>(s_slot + 2 * NW);
  float* sy = sx + N;
  float* sz = sy + N;
  float* s_ctr = sz + N;  // m*3 centroids
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* base = xyz.p + (long long)b * xyz.sb;

  for (int k = tid; k < N; k += T) {
    sx[k] = base[(long long)k * xyz.sk];
    sy[k] = base[(long long)k * xyz.sk + xyz.sc];
    sz[k] = base[(long long)k * xyz.sk + 2 * xyz.sc];
  }
  __syncthreads();

  float px[P], py[P], pz[P], pt[P];
  unsigned valid = 0;
#pragma unroll
  for (int i = 0; i < P; ++i) {
    int k = tid + T * i;
    pt[i] = 1e10f;
    if (k < N) {
      px[i] = sx[k];
      py[i] = sy[k];
      pz[i] = sz[k];
      float mag = __fmaf_rn(pz[i], pz[i], __fmaf_rn(py[i], py[i], __fmul_rn(px[i], px[i])));
      if (!((double)mag <= 1e-3)) valid |= (1u << i);  // upstream skips |p|^2 <= 1e-3 (double literal)
    } else {
      px[i] = py[i] = pz[i] = 0.f;
    }
  }

  int old = 0;
  if (tid == 0) {
    fps_idx[(long long)b * m] = 0;
    s_ctr[0] = sx[0];
    s_ctr[1] = sy[0];
    s_ctr[2] = sz[0];
  }
  float ox = sx[0], oy = sy[0], oz = sz[0];
  for (int j = 1; j < m; ++j) {
    unsigned long long best = 0ull;
#pragma unroll
    for (int i = 0; i < P; ++i) {
      if (valid & (1u << i)) {
        float d = sqdist3(px[i], py[i], pz[i], ox, oy, oz);
        float t = fminf(d, pt[i]);
        pt[i] = t;
        unsigned low = 0xFFFFFFFFu - fps_rank(tid + T * i, bs, log2bs, per);
        unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | low;
        best = key > best ? key : best;
      }
    }
    best = warp_max_u64(best);
    unsigned long long* slot = s_slot + (j & 1) * NW;
    if (lane == 0) slot[warp] = best;
    __syncthreads();
    unsigned long long v = slot[lane & (NW - 1)];
    v = warp_max_u64(v);
    old = v ? fps_unrank(0xFFFFFFFFu - (unsigned)(v & 0xFFFFFFFFull), bs, log2bs, per) : 0;
    ox = sx[old];
    oy = sy[old];
    oz = sz[old];
    if (tid == 0) {
      fps_idx[(long long)b * m + j] = old;
      s_ctr[j * 3 + 0] = ox;
      s_ctr[j * 3 + 1] = oy;
      s_ctr[j * 3 + 2] = oz;
    }
  }
  __syncthreads();
  if (new_xyz) {
    for (int e = tid; e < m * 3; e += T) new_xyz[(long long)b * m * 3 + e] = s_ctr[e];
  }
  if (do_bq) {
    for (int j = warp; j < m; j += NW) {
      ball_query_warp(sx, sy, sz, N, s_ctr[j * 3], s_ctr[j * 3 + 1], s_ctr[j * 3 + 2], radius2, nsample,
                      bq_idx + ((long long)b * m + j) * nsample, bq_cnt ? bq_cnt + (long long)b * m + j : nullptr,
                      lane);
    }
  }
}

// stand-alone ball query (compat op): stage the cloud, then warp-per-centroid
__global__ void __launch_bounds__(512) ball_query_kernel(XyzView xyz, int N, int m, const float* __restrict__ new_xyz,
                                                         float radius2, int nsample, int32_t* __restrict__ bq_idx,
                                                         int32_t* __restrict__ bq_cnt) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sx = reinterpret_cast<float*>(smem_raw);
  float* sy = sx + N;
  float* sz = sy + N;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* base = xyz.p + (long long)b * xyz.sb;
  for (int k = tid; k < N; k += blockDim.x) {
    sx[k] = base[(long long)k * xyz.sk];
    sy[k] = base[(long long)k * xyz.sk + xyz.sc];
    sz[k] = base[(long long)k * xyz.sk + 2 * xyz.sc];
  }
  __syncthreads();
  const float* q = new_xyz + (long long)b * m * 3;
  for (int j = warp; j < m; j += blockDim.x / 32) {
    ball_query_warp(sx, sy, sz, N, q[j * 3], q[j * 3 + 1], q[j * 3 + 2], radius2, nsample,
                    bq_idx + ((long long)b * m + j) * nsample, bq_cnt ? bq_cnt + (long long)b * m + j : nullptr, lane);
  }
}

// ---- compat gather / group ops (not on the fused product path) -------------------------------
__global__ void gather_points_kernel(int B, int C, int N, int m, const float* __restrict__ pts,
                                     const int32_t* __restrict__ idx, float* __restrict__ out) {
  long long total = (long long)B * C * m;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int j = (int)(e % m);
    long long bc = e / m;
    int b = (int)(bc / C);
    out[e] = pts[bc * N + idx[(long long)b * m + j]];
  }
}
// deterministic: one thread per (b,c) walks the m entries in order
__global__ void gather_points_grad_kernel(int B, int C, int N, int m, const float* __restrict__ grad_out,
                                          const int32_t* __restrict__ idx, float* __restrict__ grad_pts) {
  long long bc = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (bc >= (long long)B * C) return;
  int b = (int)(bc / C);
  float* dst = grad_pts + bc * N;
  for (int j = 0; j < m; ++j) dst[idx[(long long)b * m + j]] += grad_out[bc * m + j];
}
__global__ void group_points_kernel(int B, int C, int N, int ms, const float* __restrict__ pts,
                                    const int32_t* __restrict__ idx, float* __restrict__ out) {
  long long total = (long long)B * C * ms;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int l = (int)(e % ms);
    long long bc = e / ms;
    int b = (int)(bc / C);
    out[e] = pts[bc * N + idx[(long long)b * ms + l]];
  }
}
__global__ void group_points_grad_kernel(int B, int C, int N, int ms, const float* __restrict__ grad_out,
                                         const int32_t* __restrict__ idx, float* __restrict__ grad_pts) {
  long long bc = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (bc >= (long long)B * C) return;
  int b = (int)(bc / C);
  float* dst = grad_pts + bc * N;
  const float* src = grad_out + bc * ms;
  const int32_t* ib = idx + (long long)b * ms;
  for (int l = 0; l < ms; ++l) dst[ib[l]] += src[l];
}

// ---- compact row table -------------------------------------------------------------------------
// A ball-query group (segment) with h hits expands to nsample slots of which only h' = max(h,1)
// are distinct: slots >= h repeat slot 0.  The shared MLP therefore runs on h' rows per segment,
// row 0 carrying multiplicity nsample-h'+1 (DESIGN.md "duplicate folding").
// seg_off[S+1]: exclusive scan of h' (single CTA, S <= 1024*64).
__global__ void __launch_bounds__(1024) seg_scan_kernel(int S, const int32_t* __restrict__ cnt,
                                                        int32_t* __restrict__ seg_off) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < S; base += 1024) {
    int i = base + tid;
    int v = 0;
    if (i < S) {
      v = cnt[i];
      v = v < 1 ? 1 : v;
    }
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += t;
      }
      s_warp[lane] = wi - w;  // exclusive warp offsets
    }
    __syncthreads();
    int carry = s_carry;
    int excl = carry + s_warp[warp] + incl - v;
    if (i < S) seg_off[i] = excl;
    __syncthreads();
    if (tid == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (tid == 0) seg_off[S] = s_carry;
}

// one warp per segment: row_seg / row_src (point index inside the sample) / row_w
__global__ void row_expand_kernel(int S, int nsample, const int32_t* __restrict__ cnt,
                                  const int32_t* __restrict__ idx, const int32_t* __restrict__ seg_off,
                                  int32_t* __restrict__ row_seg, int32_t* __restrict__ row_src,
                                  float* __restrict__ row_w) {
  int seg = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5);
  int lane = threadIdx.x & 31;
  if (seg >= S) return;
  int h = cnt[seg];
  h = h < 1 ? 1 : h;
  int off = seg_off[seg];
  for (int s = lane; s < h; s += 32) {
    row_seg[off + s] = seg;
    row_src[off + s] = idx[(long long)seg * nsample + s];
    row_w[off + s] = (s == 0) ? (float)(nsample - h + 1) : 1.0f;
  }
}

int ilog2_floor(int n) {
  int l = 0;
  while ((2 << l) <= n) ++l;
  return l;
}

// FPS for clouds too large for the register-resident kernel (N > 8192: the environment-side down-sampler
// regularize_pc_point_count, /root/reference/core/utils.py:784-812, sees raw depth clouds).  Same Spec S1 arithmetic and
// packed-key arg-max; the running min-distance lives in shared memory (4 B/point), xyz is re-read from L2 each round
// (12 B/point, 240 KB at N = 20 000) and the selected centroids go straight to global memory, so neither N nor m is
// bounded by the centroid stage.  One CTA of 1024 threads per cloud.
template <int T>
__global__ void __launch_bounds__(T) fps_large_kernel(XyzView xyz, int N, int m, int bs, int log2bs, int per,
                                                      int32_t* __restrict__ fps_idx, float* __restrict__ new_xyz) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int NW = T / 32;
  unsigned long long* s_slot = reinterpret_cast<unsigned long long*>(smem_raw);
  float* s_t = reinterpret_cast<float*>(s_slot + 2 * NW);
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* base = xyz.p + (long long)b * xyz.sb;
  for (int k = tid; k < N; k += T) s_t[k] = 1e10f;
  int old = 0;
  float ox = base[0], oy = base[xyz.sc], oz = base[2 * xyz.sc];
  if (tid == 0) {
    fps_idx[(long long)b * m] = 0;
    if (new_xyz) {
      float* c = new_xyz + (long long)b * m * 3;
      c[0] = ox, c[1] = oy, c[2] = oz;
    }
  }
  __syncthreads();
  for (int j = 1; j < m; ++j) {
    unsigned long long best = 0ull;
    for (int k = tid; k < N; k += T) {
      const float* q = base + (long long)k * xyz.sk;
      float x = __ldg(q), y = __ldg(q + xyz.sc), z = __ldg(q + 2 * xyz.sc);
      float mag = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
      if ((double)mag <= 1e-3) continue;
      float t = fminf(sqdist3(x, y, z, ox, oy, oz), s_t[k]);
      s_t[k] = t;
      unsigned low = 0xFFFFFFFFu - fps_rank(k, bs, log2bs, per);
      unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | low;
      best = key > best ? key : best;
    }
    best = warp_max_u64(best);
    unsigned long long* slot = s_slot + (j & 1) * NW;
    if (lane == 0) slot[warp] = best;
    __syncthreads();
    unsigned long long v = slot[lane & (NW - 1)];
    v = warp_max_u64(v);
    old = v ? fps_unrank(0xFFFFFFFFu - (unsigned)(v & 0xFFFFFFFFull), bs, log2bs, per) : 0;
    const float* q = base + (long long)old * xyz.sk;
    ox = __ldg(q), oy = __ldg(q + xyz.sc), oz = __ldg(q + 2 * xyz.sc);
    if (tid == 0) {
      fps_idx[(long long)b * m + j] = old;
      if (new_xyz) {
        float* c = new_xyz + ((long long)b * m + j) * 3;
        c[0] = ox, c[1] = oy, c[2] = oz;
      }
    }
  }
}

template <int T, int P>
int launch_fps(XyzView v, int B, int N, int m, int bs, int32_t* fps_idx, float* new_xyz, int do_bq, float radius,
               int nsample, int32_t* bq_idx, int32_t* bq_cnt, cudaStream_t st) {
  size_t smem = sizeof(float) * (3 * (size_t)N + (size_t)m * 3) + sizeof(unsigned long long) * 2 * (T / 32);
  auto kern = fps_ballquery_kernel<T, P>;
  if (smem > 48 * 1024) GADDPG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int log2bs = ilog2_floor(bs);
  int per = (N + bs - 1) / bs;
  kern<<<B, T, smem, st>>>(v, N, m, bs, log2bs, per, fps_idx, new_xyz, do_bq, radius * radius, nsample, bq_idx, bq_cnt);
  GADDPG_CHECK_LAUNCH("fps_ballquery_kernel");
  return GADDPG_OK;
}

}  // namespace

// upstream cuda_utils.h opt_n_threads(): the virtual block size that fixes FPS tie-breaking
extern "C" __attribute__((visibility("default"))) int gaddpg_opt_n_threads(int work_size) {
  int p = 1;
  if (work_size < 1) return 1;
  while ((p << 1) <= work_size) p <<= 1;
  return p > 512 ? 512 : p;
}

int gaddpg_fps_ballquery_impl(const float* xyz, long long sb, int sk, int sc, int B, int N, int m, int32_t* fps_idx,
                              float* new_xyz, int do_bq, float radius, int nsample, int32_t* bq_idx, int32_t* bq_cnt,
                              void* stream) {
  GADDPG_CHECK_ARG(xyz && fps_idx, "fps: null pointer");
  GADDPG_CHECK_ARG(B >= 0 && N >= 1 && m >= 1, "fps: bad shape B=%d N=%d m=%d", B, N, m);
  GADDPG_CHECK_ARG(!do_bq || (bq_idx && nsample >= 1), "fps_ballquery: ball query outputs missing");
  if (B == 0) return GADDPG_OK;
  XyzView v{xyz, sb, sk, sc};
  cudaStream_t st = (cudaStream_t)stream;
  int bs = gaddpg_opt_n_threads(N);
  if (N > 8192 || m * 3 > 8192) {
    if (do_bq) {
      gaddpg_set_error("fps_ballquery: N=%d / npoint=%d exceed the fused kernel (N <= 8192, npoint <= 2730); call fps alone", N, m);
      return GADDPG_ERR_UNSUPPORTED;
    }
    constexpr int T = 1024;
    size_t smem = sizeof(float) * (size_t)N + sizeof(unsigned long long) * 2 * (T / 32);
    if (smem > 200 * 1024) {
      gaddpg_set_error("fps: N=%d points per cloud exceed the shared-memory distance table (N <= 51072)", N);
      return GADDPG_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024)
      GADDPG_CUDA(cudaFuncSetAttribute(fps_large_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fps_large_kernel<T><<<B, T, smem, st>>>(v, N, m, bs, ilog2_floor(bs), (N + bs - 1) / bs, fps_idx, new_xyz);
    GADDPG_CHECK_LAUNCH("fps_large_kernel");
    return GADDPG_OK;
  }
#define L(T, P) return launch_fps<T, P>(v, B, N, m, bs, fps_idx, new_xyz, do_bq, radius, nsample, bq_idx, bq_cnt, st)
  if (N <= 128) L(128, 1);
  int p = (N + 511) / 512;
  if (p <= 1) L(512, 1);
  if (p <= 2) L(512, 2);
  if (p <= 4) L(512, 4);
  if (p <= 8) L(512, 8);
  L(512, 16);
#undef L
}

int gaddpg_ball_query_impl(const float* xyz, long long sb, int sk, int sc, int B, int N, int m, const float* new_xyz,
                           float radius, int nsample, int32_t* idx, int32_t* cnt, void* stream) {
  GADDPG_CHECK_ARG(xyz && new_xyz && idx, "ball_query: null pointer");
  GADDPG_CHECK_ARG(B >= 0 && N >= 1 && m >= 1 && nsample >= 1, "ball_query: bad shape");
  if (B == 0) return GADDPG_OK;
  size_t smem = sizeof(float) * 3 * (size_t)N;
  if (smem > 200 * 1024) {
    gaddpg_set_error("ball_query: N=%d does not fit the shared-memory stage", N);
    return GADDPG_ERR_UNSUPPORTED;
  }
  if (smem > 48 * 1024)
    GADDPG_CUDA(cudaFuncSetAttribute(ball_query_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  XyzView v{xyz, sb, sk, sc};
  ball_query_kernel<<<B, 512, smem, (cudaStream_t)stream>>>(v, N, m, new_xyz, radius * radius, nsample, idx, cnt);
  GADDPG_CHECK_LAUNCH("ball_query_kernel");
  return GADDPG_OK;
}

int gaddpg_gather_points_impl(int B, int C, int N, int m, const float* pts, const int32_t* idx, float* out, void* stream) {
  long long total = (long long)B * C * m;
  if (total == 0) return GADDPG_OK;
  int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  gather_points_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(B, C, N, m, pts, idx, out);
  GADDPG_CHECK_LAUNCH("gather_points_kernel");
  return GADDPG_OK;
}
int gaddpg_gather_points_grad_impl(int B, int C, int N, int m, const float* grad_out, const int32_t* idx,
                                   float* grad_pts, void* stream) {
  long long bc = (long long)B * C;
  if (bc == 0) return GADDPG_OK;
  GADDPG_CUDA(cudaMemsetAsync(grad_pts, 0, sizeof(float) * bc * N, (cudaStream_t)stream));
  gather_points_grad_kernel<<<(int)((bc + 127) / 128), 128, 0, (cudaStream_t)stream>>>(B, C, N, m, grad_out, idx, grad_pts);
  GADDPG_CHECK_LAUNCH("gather_points_grad_kernel");
  return GADDPG_OK;
}
int gaddpg_group_points_impl(int B, int C, int N, int m, int s, const float* pts, const int32_t* idx, float* out,
                             void* stream) {
  long long total = (long long)B * C * m * s;
  if (total == 0) return GADDPG_OK;
  int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  group_points_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(B, C, N, m * s, pts, idx, out);
  GADDPG_CHECK_LAUNCH("group_points_kernel");
  return GADDPG_OK;
}
int gaddpg_group_points_grad_impl(int B, int C, int N, int m, int s, const float* grad_out, const int32_t* idx,
                                  float* grad_pts, void* stream) {
  long long bc = (long long)B * C;
  if (bc == 0) return GADDPG_OK;
  GADDPG_CUDA(cudaMemsetAsync(grad_pts, 0, sizeof(float) * bc * N, (cudaStream_t)stream));
  group_points_grad_kernel<<<(int)((bc + 127) / 128), 128, 0, (cudaStream_t)stream>>>(B, C, N, m * s, grad_out, idx, grad_pts);
  GADDPG_CHECK_LAUNCH("group_points_grad_kernel");
  return GADDPG_OK;
}

int gaddpg_row_table_impl(int S, int nsample, const int32_t* cnt, const int32_t* idx, int32_t* seg_off,
                          int32_t* row_seg, int32_t* row_src, float* row_w, void* stream) {
  GADDPG_CHECK_ARG(S >= 1 && S <= 1024 * 64, "row_table: S=%d out of range", S);
  GADDPG_CHECK_ARG(cnt && idx && seg_off && row_seg && row_src && row_w, "row_table: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  seg_scan_kernel<<<1, 1024, 0, st>>>(S, cnt, seg_off);
  GADDPG_CHECK_LAUNCH("seg_scan_kernel");
  row_expand_kernel<<<ceil_div(S * 32, 256), 256, 0, st>>>(S, nsample, cnt, idx, seg_off, row_seg, row_src, row_w);
  GADDPG_CHECK_LAUNCH("row_expand_kernel");
  return GADDPG_OK;
}
