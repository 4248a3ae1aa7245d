// optim.cu — flat-arena optimiser kernels (sm_100a): Adam with L2 weight decay, gradient-norm clipping,
// Polyak target updates, abs-max statistics, derived weight layouts, input conversion.
//
// Replaces torch.optim.Adam.step (policy/critic: eps=1e-5, weight_decay=1e-5, /root/reference/core/utils.py:
// 969-970,993-994; encoders: lr=1e-3, utils.py:221-234), torch.nn.utils.clip_grad_norm_ (ddpg.py:141),
// soft_update / half_soft_update / half_hard_update (utils.py:750-770), module_max_param / module_max_gradient
// (utils.py:92-108) and torch.cuda.FloatTensor(ndarray) (agent.py:221-222).  These are the HBM-streaming
// kernels of the step: 16 B read + 12 B written per parameter for Adam (+8/4 for a fused Polyak target),
// float4 vectorised, grid-stride over 148 x 8 CTAs.
#include "common.cuh"
#include "impl.h"

namespace {

struct AdamArgs {
  float lr_over_bc1;    // step_size = lr / (1 - beta1^t)
  float bc2_sqrt;       // sqrt(1 - beta2^t)
  float beta1, beta2, one_minus_beta1, one_minus_beta2, eps, weight_decay;
  float grad_scale;     // 1/world_size for sample-sharded replicas
  const float* clip;    // device scalar: clip coefficient (NULL = 1)
  const float* dyn;     // device {step_size, bc2_sqrt}: overrides the two immediates (graph replay across steps)
  int write_back_grad;  // store the clipped/scaled gradient (clip_grad_norm_ semantics)
  float tau;            // Polyak factor when target != NULL
  float one_minus_tau;
};

__device__ __forceinline__ void adam_one(float& p, float& g, float& m, float& v, const AdamArgs& a, float clip) {
  g = g * a.grad_scale * clip;
  float gg = fmaf(a.weight_decay, p, g);             // grad.add(param, alpha=weight_decay)
  m = fmaf(gg - m, a.one_minus_beta1, m);            // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(a.one_minus_beta2 * gg, gg, a.beta2 * v); // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
  float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p = p - a.lr_over_bc1 * (m / denom);               // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, AdamArgs a, float* __restrict__ target) {
  const float clip = a.clip ? *a.clip : 1.f;
  if (a.dyn) {
    a.lr_over_bc1 = a.dyn[0];
    a.bc2_sqrt = a.dyn[1];
  }
  const long long n4 = n / 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 P = reinterpret_cast<float4*>(p)[i], G = reinterpret_cast<float4*>(g)[i], Mm = reinterpret_cast<float4*>(m)[i],
           V = reinterpret_cast<float4*>(v)[i];
    adam_one(P.x, G.x, Mm.x, V.x, a, clip);
    adam_one(P.y, G.y, Mm.y, V.y, a, clip);
    adam_one(P.z, G.z, Mm.z, V.z, a, clip);
    adam_one(P.w, G.w, Mm.w, V.w, a, clip);
    reinterpret_cast<float4*>(p)[i] = P;
    reinterpret_cast<float4*>(m)[i] = Mm;
    reinterpret_cast<float4*>(v)[i] = V;
    if (a.write_back_grad) reinterpret_cast<float4*>(g)[i] = G;
    if (target) {
      float4 T = reinterpret_cast<float4*>(target)[i];
      T.x = T.x * a.one_minus_tau + P.x * a.tau;
      T.y = T.y * a.one_minus_tau + P.y * a.tau;
      T.z = T.z * a.one_minus_tau + P.z * a.tau;
      T.w = T.w * a.one_minus_tau + P.w * a.tau;
      reinterpret_cast<float4*>(target)[i] = T;
    }
  }
  for (long long i = n4 * 4 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float P = p[i], G = g[i], Mm = m[i], V = v[i];
    adam_one(P, G, Mm, V, a, clip);
    p[i] = P;
    m[i] = Mm;
    v[i] = V;
    if (a.write_back_grad) g[i] = G;
    if (target) target[i] = target[i] * a.one_minus_tau + P * a.tau;
  }
}

// ---- one launch for a whole optimiser phase -----------------------------------------------------------------------------
// A job table (device memory, written once) lists element ranges with what happens to them: Adam (optionally followed by the
// Polyak update of a target copy and the abs-max of the updated parameters), Polyak only (scalar or per-element tau) or abs-max
// only.  Chunks of OPT_CHUNK elements of all jobs form one grid-stride space, so the five Adam launches, three two-stage
// abs-max reductions and two or three Polyak launches of a DDPG step (agent.py:192-209,242-259) are two launches.  Same
// arithmetic per element as adam_kernel / polyak_kernel / polyak_vec_kernel.  abs-max is exact and order-independent:
// non-negative floats compare like their bit patterns, so one atomicMax per block is deterministic.
constexpr int OPT_CHUNK = 4096;
struct OptJob {
  float* p; float* g; float* m; float* v; float* target; const float* tau_vec; const float* dyn; const float* clip;
  float* absmax_p; float* absmax_g;
  long long n, chunk0;
  float eps, weight_decay, tau;
  int kind, write_back, pad;   // kind 0: Adam; 1: Polyak (source = p); 2: statistics only
};
static_assert(sizeof(OptJob) == 120, "OptJob layout is mirrored in agent.py");

__global__ void __launch_bounds__(256) optim_multi_kernel(const OptJob* __restrict__ jobs, int njobs, int total_chunks) {
  __shared__ float red[2][8];
  for (int cid = blockIdx.x; cid < total_chunks; cid += gridDim.x) {   // block-uniform
    int j = 0;
    while (j + 1 < njobs && jobs[j + 1].chunk0 <= cid) ++j;
    const OptJob J = jobs[j];
    const long long e0 = (long long)(cid - J.chunk0) * OPT_CHUNK;
    long long e1 = e0 + OPT_CHUNK;
    e1 = e1 < J.n ? e1 : J.n;
    float mxp = 0.f, mxg = 0.f;
    if (J.kind == 0) {
      AdamArgs a;
      a.lr_over_bc1 = J.dyn[0];
      a.bc2_sqrt = J.dyn[1];
      a.beta1 = 0.9f;
      a.beta2 = 0.999f;
      a.one_minus_beta1 = (float)(1.0 - 0.9);
      a.one_minus_beta2 = (float)(1.0 - 0.999);
      a.eps = J.eps;
      a.weight_decay = J.weight_decay;
      a.grad_scale = 1.f;
      const float clip = J.clip ? *J.clip : 1.f;
      const float omt = (float)(1.0 - (double)J.tau);
      for (long long i = e0 + threadIdx.x * 4; i < e1; i += 256 * 4) {   // job starts are 16-byte aligned; the last <= 3 elements go scalar
        if (i + 4 > e1) {
          for (long long k = i; k < e1; ++k) {
            float P = J.p[k], G = J.g[k], Mm = J.m[k], V = J.v[k];
            adam_one(P, G, Mm, V, a, clip);
            J.p[k] = P;
            J.m[k] = Mm;
            J.v[k] = V;
            if (J.write_back) J.g[k] = G;
            if (J.target) J.target[k] = J.target[k] * omt + P * J.tau;
            mxp = fmaxf(mxp, fabsf(P));
            mxg = fmaxf(mxg, fabsf(G));
          }
          break;
        }
        float4 P = *reinterpret_cast<float4*>(J.p + i), G = *reinterpret_cast<float4*>(J.g + i), Mm = *reinterpret_cast<float4*>(J.m + i),
               V = *reinterpret_cast<float4*>(J.v + i);
        adam_one(P.x, G.x, Mm.x, V.x, a, clip);
        adam_one(P.y, G.y, Mm.y, V.y, a, clip);
        adam_one(P.z, G.z, Mm.z, V.z, a, clip);
        adam_one(P.w, G.w, Mm.w, V.w, a, clip);
        *reinterpret_cast<float4*>(J.p + i) = P;
        *reinterpret_cast<float4*>(J.m + i) = Mm;
        *reinterpret_cast<float4*>(J.v + i) = V;
        if (J.write_back) *reinterpret_cast<float4*>(J.g + i) = G;
        if (J.target) {
          float4 T = *reinterpret_cast<float4*>(J.target + i);
          T.x = T.x * omt + P.x * J.tau;
          T.y = T.y * omt + P.y * J.tau;
          T.z = T.z * omt + P.z * J.tau;
          T.w = T.w * omt + P.w * J.tau;
          *reinterpret_cast<float4*>(J.target + i) = T;
        }
        mxp = fmaxf(fmaxf(mxp, fmaxf(fabsf(P.x), fabsf(P.y))), fmaxf(fabsf(P.z), fabsf(P.w)));
        mxg = fmaxf(fmaxf(mxg, fmaxf(fabsf(G.x), fabsf(G.y))), fmaxf(fabsf(G.z), fabsf(G.w)));
      }
    } else {
      const float omt = (float)(1.0 - (double)J.tau);
      for (long long i = e0 + threadIdx.x; i < e1; i += 256) {
        const float P = J.p[i];
        if (J.kind == 1) {
          if (J.tau_vec) {
            const float tau = J.tau_vec[i];
            if (tau != 0.f) J.target[i] = J.target[i] * (1.0f - tau) + P * tau;
          } else {
            J.target[i] = J.target[i] * omt + P * J.tau;
          }
        }
        mxp = fmaxf(mxp, fabsf(P));
        if (J.absmax_g) mxg = fmaxf(mxg, fabsf(J.g[i]));
      }
    }
    if (J.absmax_p || J.absmax_g) {   // block-uniform
      mxp = warp_max(mxp);
      mxg = warp_max(mxg);
      __syncthreads();
      if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = mxp;
        red[1][threadIdx.x >> 5] = mxg;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        float a = red[0][0], b = red[1][0];
        for (int w = 1; w < 8; ++w) {
          a = fmaxf(a, red[0][w]);
          b = fmaxf(b, red[1][w]);
        }
        if (J.absmax_p) atomicMax(reinterpret_cast<unsigned int*>(J.absmax_p), __float_as_uint(a));
        if (J.absmax_g) atomicMax(reinterpret_cast<unsigned int*>(J.absmax_g), __float_as_uint(b));
      }
    }
  }
}

// target = target*(1-tau) + source*tau   (utils.py:750-754)
__global__ void polyak_kernel(float* __restrict__ t, const float* __restrict__ s, long long n, float tau, float one_minus_tau) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    t[i] = t[i] * one_minus_tau + s[i] * tau;
}

__global__ void polyak_vec_kernel(float* __restrict__ t, const float* __restrict__ s, const float* __restrict__ tv, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float tau = tv[i];
    if (tau != 0.f) t[i] = t[i] * (1.0f - tau) + s[i] * tau;
  }
}

// two-stage deterministic reductions: per-CTA partial -> fixed-order final
__global__ void __launch_bounds__(256) reduce_partial_kernel(const float* __restrict__ x, long long n, int op,
                                                             float* __restrict__ partial) {
  __shared__ float red[8];
  float a = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = x[i];
    a = op == 0 ? fmaxf(a, fabsf(v)) : fmaf(v, v, a);
  }
  a = op == 0 ? warp_max(a) : warp_sum(a);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = red[0];
    for (int w = 1; w < 8; ++w) r = op == 0 ? fmaxf(r, red[w]) : r + red[w];
    partial[blockIdx.x] = r;
  }
}
// op 0: out = max(partials)   op 1: out = min(1, max_norm / (sqrt(sum partials) + 1e-6))  [clip coefficient]
// one CTA of 256 threads: thread t folds partials t, t+256, ... and thread 0 folds the 256 thread results, both in a
// fixed order (deterministic); the single-thread loop this replaces took ~28 us for 1184 dependent loads.
__global__ void __launch_bounds__(256) reduce_final_kernel(const float* __restrict__ partial, int nblk, int op, float max_norm,
                                                           float* __restrict__ out, float* __restrict__ norm_out) {
  __shared__ double red[256];
  const int t = threadIdx.x;
  double a = 0.0;
  for (int i = t; i < nblk; i += 256) a = op == 0 ? fmax(a, (double)partial[i]) : a + (double)partial[i];
  red[t] = a;
  __syncthreads();
#pragma unroll
  for (int w = 128; w >= 1; w >>= 1) {
    if (t < w) red[t] = op == 0 ? fmax(red[t], red[t + w]) : red[t] + red[t + w];
    __syncthreads();
  }
  if (t != 0) return;
  if (op == 0) {
    *out = (float)red[0];
  } else {
    float nrm = (float)sqrt(red[0]);
    float c = max_norm / (nrm + 1e-6f);
    *out = c < 1.f ? c : 1.f;
    if (norm_out) *norm_out = nrm;
  }
}

// derived weight layouts: Wp[n][kp] = W[n][(kp+rot) % K] for kp < K else 0 (ldp >= K, multiple of 4);
// WT[kp][n] = Wp[n][kp] with row length ldt >= N (multiple of 4), rows kp < ldp
__global__ void wprep_kernel(const float* __restrict__ W, int N, int K, int rot, float* __restrict__ Wp, int ldp,
                             float* __restrict__ WT, int ldt) {
  long long total = (long long)(Wp ? N : 0) * ldp;
  long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  for (long long e = i0; e < total; e += stride) {
    int n = (int)(e / ldp), kp = (int)(e % ldp);
    float v = 0.f;
    if (kp < K) {
      int k = kp + rot;
      if (k >= K) k -= K;
      v = W[(long long)n * K + k];
    }
    Wp[e] = v;
  }
  if (WT) {
    long long tt = (long long)ldp * ldt;
    for (long long e = i0; e < tt; e += stride) {
      int kp = (int)(e / ldt), n = (int)(e % ldt);
      float v = 0.f;
      if (n < N && kp < K) {
        int k = kp + rot;
        if (k >= K) k -= K;
        v = W[(long long)n * K + k];
      }
      WT[e] = v;
    }
  }
}

// batched form: jobs[j] = {W, N, K, rot, Wp, ldp, WT, ldt} as 8 x int64 in device memory (pointers are static).
// 32x32 tiles through shared memory: W is read once, coalesced along k; Wp is written from the same registers
// (coalesced along kp) and WT from the transposed tile (coalesced along n).
__global__ void __launch_bounds__(256) wprep_batched_kernel(const long long* __restrict__ jobs) {
  __shared__ float tile[32][33];
  const long long* J = jobs + (long long)blockIdx.y * 8;
  const float* W = reinterpret_cast<const float*>(J[0]);
  const int N = (int)J[1], K = (int)J[2], rot = (int)J[3];
  float* Wp = reinterpret_cast<float*>(J[4]);
  const int ldp = (int)J[5];
  float* WT = reinterpret_cast<float*>(J[6]);
  const int ldt = (int)J[7];
  if (rot < 0) {   // hi/lo TF32 split job (fused SA1 chain): Wp = hi[N][ldp], WT = lo[N][ldp], columns >= K zero
    for (int e = blockIdx.x * 256 + threadIdx.x; e < N * ldp; e += gridDim.x * 256) {
      const int n = e / ldp, k = e % ldp;
      const float x = k < K ? W[(long long)n * K + k] : 0.f;
      const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
      Wp[e] = h;
      WT[e] = x - h;
    }
    return;
  }
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int nmax = (WT && ldt > N) ? ldt : N;
  const int tiles_n = (nmax + 31) >> 5, tiles_k = (ldp + 31) >> 5;
  for (int t = blockIdx.x; t < tiles_n * tiles_k; t += gridDim.x) {
    const int n0 = (t / tiles_k) << 5, k0 = (t % tiles_k) << 5;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + ty + 8 * j, kp = k0 + tx;
      float v = 0.f;
      if (n < N && kp < K) {
        int k = kp + rot;
        if (k >= K) k -= K;
        v = W[(long long)n * K + k];
      }
      tile[ty + 8 * j][tx] = v;
      if (Wp && n < N && kp < ldp) Wp[(long long)n * ldp + kp] = v;
    }
    __syncthreads();
    if (WT) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kp = k0 + ty + 8 * j, n = n0 + tx;
        if (kp < ldp && n < ldt) WT[(long long)kp * ldt + n] = tile[tx][ty + 8 * j];
      }
    }
    __syncthreads();
  }
}

__global__ void f64_to_f32_kernel(const double* __restrict__ src, float* __restrict__ dst, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = (float)src[i];
}

int stream_grid(long long n) {
  long long g = (n + 255) / 256;
  long long cap = (long long)gaddpg_sm_count() * 8;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace

int gaddpg_adam_step_impl(float* p, float* g, float* m, float* v, long long n, double lr, double beta1, double beta2, double eps,
                          double weight_decay, long long step, const float* dyn, double grad_scale, const float* clip,
                          int write_back_grad, float* target, double tau, void* stream) {
  GADDPG_CHECK_ARG(p && g && m && v && n >= 0 && (step >= 1 || dyn), "adam_step: bad argument");
  GADDPG_CHECK_ARG(((uintptr_t)p % 16) == 0 && ((uintptr_t)g % 16) == 0 && ((uintptr_t)m % 16) == 0 && ((uintptr_t)v % 16) == 0 &&
                       (!target || ((uintptr_t)target % 16) == 0),
                   "adam_step: arena segments must be 16-byte aligned");
  if (n == 0) return GADDPG_OK;
  AdamArgs a;
  if (step < 1) step = 1;
  double bc1 = 1.0 - pow(beta1, (double)step), bc2 = 1.0 - pow(beta2, (double)step);
  a.dyn = dyn;
  a.lr_over_bc1 = (float)(lr / bc1);
  a.bc2_sqrt = (float)sqrt(bc2);
  a.beta1 = (float)beta1;
  a.beta2 = (float)beta2;
  a.one_minus_beta1 = (float)(1.0 - beta1);
  a.one_minus_beta2 = (float)(1.0 - beta2);
  a.eps = (float)eps;
  a.weight_decay = (float)weight_decay;
  a.grad_scale = (float)grad_scale;
  a.clip = clip;
  a.write_back_grad = write_back_grad;
  a.tau = (float)tau;
  a.one_minus_tau = (float)(1.0 - tau);
  adam_kernel<<<stream_grid(n / 4 + 1), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, a, target);
  GADDPG_CHECK_LAUNCH("adam_kernel");
  return GADDPG_OK;
}

int gaddpg_optim_multi_impl(const void* jobs_dev, int njobs, int total_chunks, void* stream) {
  GADDPG_CHECK_ARG(jobs_dev && njobs >= 1 && njobs <= 64 && total_chunks >= 1, "optim_multi: bad argument");
  const int cap = gaddpg_sm_count() * 8;
  optim_multi_kernel<<<total_chunks < cap ? total_chunks : cap, 256, 0, (cudaStream_t)stream>>>((const OptJob*)jobs_dev, njobs, total_chunks);
  GADDPG_CHECK_LAUNCH("optim_multi_kernel");
  return GADDPG_OK;
}

int gaddpg_polyak_impl(float* target, const float* source, long long n, double tau, void* stream) {
  GADDPG_CHECK_ARG(target && source && n >= 0, "polyak: bad argument");
  if (n == 0) return GADDPG_OK;
  polyak_kernel<<<stream_grid(n), 256, 0, (cudaStream_t)stream>>>(target, source, n, (float)tau, (float)(1.0 - tau));
  GADDPG_CHECK_LAUNCH("polyak_kernel");
  return GADDPG_OK;
}

int gaddpg_polyak_vec_impl(float* target, const float* source, const float* tau_vec, long long n, void* stream) {
  GADDPG_CHECK_ARG(target && source && tau_vec && n >= 0, "polyak_vec: bad argument");
  if (n == 0) return GADDPG_OK;
  polyak_vec_kernel<<<stream_grid(n), 256, 0, (cudaStream_t)stream>>>(target, source, tau_vec, n);
  GADDPG_CHECK_LAUNCH("polyak_vec_kernel");
  return GADDPG_OK;
}

// ws: >= 1184 floats
int gaddpg_absmax_impl(const float* x, long long n, float* out, float* ws, void* stream) {
  GADDPG_CHECK_ARG(x && out && ws && n >= 1, "absmax: bad argument");
  int grid = stream_grid(n);
  reduce_partial_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, n, 0, ws);
  GADDPG_CHECK_LAUNCH("reduce_partial_kernel");
  reduce_final_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(ws, grid, 0, 0.f, out, nullptr);
  GADDPG_CHECK_LAUNCH("reduce_final_kernel");
  return GADDPG_OK;
}

int gaddpg_clip_coef_impl(const float* g, long long n, float max_norm, float* coef_out, float* norm_out, float* ws, void* stream) {
  GADDPG_CHECK_ARG(g && coef_out && ws && n >= 1, "clip_coef: bad argument");
  int grid = stream_grid(n);
  reduce_partial_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, n, 1, ws);
  GADDPG_CHECK_LAUNCH("reduce_partial_kernel");
  reduce_final_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(ws, grid, 1, max_norm, coef_out, norm_out);
  GADDPG_CHECK_LAUNCH("reduce_final_kernel");
  return GADDPG_OK;
}

int gaddpg_wprep_impl(const float* W, int N, int K, int rot, float* Wp, int ldp, float* WT, int ldt, void* stream) {
  GADDPG_CHECK_ARG(W && N >= 1 && K >= 1 && rot >= 0 && rot < K, "wprep: bad argument");
  GADDPG_CHECK_ARG(ldp >= K && (ldp % 4) == 0 && (!WT || (ldt >= N && (ldt % 4) == 0)), "wprep: bad leading dimensions");
  long long work = (long long)(N > ldt ? N : ldt) * ldp;
  wprep_kernel<<<stream_grid(work), 256, 0, (cudaStream_t)stream>>>(W, N, K, rot, Wp, ldp, WT, ldt);
  GADDPG_CHECK_LAUNCH("wprep_kernel");
  return GADDPG_OK;
}

int gaddpg_f64_to_f32_impl(const double* src, float* dst, long long n, void* stream) {
  GADDPG_CHECK_ARG(src && dst && n >= 0, "f64_to_f32: bad argument");
  if (n == 0) return GADDPG_OK;
  f64_to_f32_kernel<<<stream_grid(n), 256, 0, (cudaStream_t)stream>>>(src, dst, n);
  GADDPG_CHECK_LAUNCH("f64_to_f32_kernel");
  return GADDPG_OK;
}

int gaddpg_wprep_batched_impl(const long long* jobs_dev, int njobs, void* stream) {
  GADDPG_CHECK_ARG(jobs_dev && njobs >= 1 && njobs <= 65535, "wprep_batched: bad argument");
  wprep_batched_kernel<<<dim3(64, njobs), 256, 0, (cudaStream_t)stream>>>(jobs_dev);
  GADDPG_CHECK_LAUNCH("wprep_batched_kernel");
  return GADDPG_OK;
}
