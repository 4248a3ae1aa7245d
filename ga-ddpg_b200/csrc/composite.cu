// composite.cu — per-level entry points of the C ABI (sa_forward_train / sa_forward_eval / sa_backward,
// adam_fused_step, *_workspace_bytes): fixed launch sequences over the kernels of this library, written in C++ so that a
// non-Python host can drive one set-abstraction level — upstream PointnetSAModule.forward + its autograd backward, reached
// from /root/reference/core/networks.py:66-81 — without re-implementing the sequencing that ga-ddpg_b200/engine.py holds for
// the fused agents.  No kernel of its own; bit-identical to the engine path (same kernels, same order).
//
// A level = shared MLP of three 1x1 convs with train-/eval-mode BatchNorm + ReLU over the compact rows, then the max-pool
// over each ball group:
//   generic level (SA2, SA3): input rows G (M_max, K0p) built by gaddpg_gather_rows
//   first level (SA1)       : input straight from the channel-major cloud (gaddpg_sa1_l1_fwd / _bwd), widths 64, 64, 128
// Workspace (caller-owned, gaddpg_sa_level_workspace_bytes): [ statistic slots | TN split partials | SA1 first-layer scratch ].
#include "common.cuh"
#include "gemm_rows.cuh"
#include "impl.h"
#include "../../include/gaddpg_b200.h"

namespace {

constexpr float kEps = 1e-5f, kMomentum = 0.1f;   // torch.nn.BatchNorm2d defaults, as instantiated by upstream build_shared_mlp

size_t stats_floats() { return (size_t)GADDPG_STAT_SLOTS * 2 * 1024 + 4; }   // + the ticket word of the BatchNorm tails (16 bytes)

struct Ws {
  float* stats;
  unsigned int* counter;   // ticket of the fused BatchNorm tails (zeroed at the start of every composite call)
  float* tn;
  size_t tn_bytes;
  float* sa1;
  size_t sa1_bytes;
};

size_t sa1_scratch_bytes(int B, int M_max) {
  const int per = ceil_div(M_max, B > 0 ? B : 1);
  const int jmax = ceil_div(per, 256);
  return ((size_t)GADDPG_STAT_SLOTS * 64 * 16 + (size_t)B * 64 * (1 + jmax) + (size_t)B * 64) * sizeof(float);
}

bool carve(void* ws, long long ws_bytes, int B, int M_max, Ws* out) {
  const size_t need = stats_floats() * sizeof(float) + gaddpg_gemm_tn_workspace_bytes_impl() + sa1_scratch_bytes(B, M_max);
  if (!ws || ws_bytes < (long long)need || ((uintptr_t)ws & 15u)) return false;
  char* p = (char*)ws;
  out->stats = (float*)p;
  out->counter = (unsigned int*)(out->stats + (size_t)GADDPG_STAT_SLOTS * 2 * 1024);
  p += stats_floats() * sizeof(float);
  out->tn = (float*)p;
  out->tn_bytes = gaddpg_gemm_tn_workspace_bytes_impl();
  p += out->tn_bytes;
  out->sa1 = (float*)p;
  out->sa1_bytes = sa1_scratch_bytes(B, M_max);
  return true;
}

Operand plain(const float* x, int ld) {
  Operand o = {};
  o.X = x;
  o.ldx = ld;
  return o;
}
Operand bnrelu(const float* y, int ld, const gaddpg_sa_layer& L) {
  Operand o = {};
  o.X = y;
  o.ldx = ld;
  o.c0 = L.scale;
  o.c1 = L.shift;
  return o;
}
Operand bnbwd(const float* d, int ldd, const float* y, int ldy, const gaddpg_sa_layer& L, const float* rw) {
  Operand o = {};
  o.X = d;
  o.ldx = ldd;
  o.Y = y;
  o.ldy = ldy;
  o.rw = rw;
  o.c0 = L.bw_g;
  o.c1 = L.bw_m1;
  o.c2 = L.bw_m2;
  o.c3 = L.mean;
  o.c4 = L.rstd;
  return o;
}

// eval mode: scale / shift from the running statistics (training passes finalize in the tail of the producing kernel)
int finalize_eval(const Ws& w, const gaddpg_sa_layer& L, double count, void* st) {
  return gaddpg_bn_finalize_fwd_impl(w.stats, L.N, count, L.gamma, L.beta, kEps, kMomentum, L.running_mean, L.running_var,
                                     L.num_batches_tracked, 0, L.scale, L.shift, L.mean, L.rstd, st);
}
// BatchNorm finalize descriptors handed to the kernel that produces the statistics (gaddpg_bn_tail)
gaddpg_bn_tail tail_fwd(const Ws& w, const gaddpg_sa_layer& L, double count, int training) {
  gaddpg_bn_tail t = {};
  if (!training) return t;
  t.kind = 1;
  t.count = count;
  t.a = L.gamma;
  t.b = L.beta;
  t.eps = kEps;
  t.momentum = kMomentum;
  t.running_mean = L.running_mean;
  t.running_var = L.running_var;
  t.num_batches_tracked = L.num_batches_tracked;
  t.o0 = L.scale;
  t.o1 = L.shift;
  t.o2 = L.mean;
  t.o3 = L.rstd;
  t.counter = w.counter;
  return t;
}
gaddpg_bn_tail tail_bwd(const Ws& w, const gaddpg_sa_layer& L, double count, int want_grads, int accumulate) {
  gaddpg_bn_tail t = {};
  t.kind = 2;
  t.accumulate = accumulate;
  t.count = count;
  t.a = L.gamma;
  t.b = L.rstd;
  t.o0 = L.bw_g;
  t.o1 = L.bw_m1;
  t.o2 = L.bw_m2;
  t.dgamma = want_grads ? L.dgamma : nullptr;
  t.dbeta = want_grads ? L.dbeta : nullptr;
  t.counter = w.counter;
  return t;
}

int nt1(const Operand& A, const float* Bw, int ldb, float* C, int ldc, int M_max, const int* M_dev, int N, int K, float* stats,
        const float* srw, const float* Yprev, int ldyp, const gaddpg_sa_layer* pbn, int amode, int emode, void* st,
        const gaddpg_bn_tail* tail = nullptr) {
  NTGroup g = {};
  NTProblem& p = g.p[0];
  if (tail && stats) p.tail = *tail;
  p.A = A;
  p.Bw = Bw;
  p.ldb = ldb;
  p.C = C;
  p.ldc = ldc;
  p.M_max = M_max;
  p.M_dev = M_dev;
  p.N = N;
  p.K = K;
  p.stats = stats;
  p.srw = srw;
  p.Yprev = Yprev;
  p.ldyp = ldyp;
  if (pbn) {
    p.psc = pbn->scale;
    p.psh = pbn->shift;
    p.pmean = pbn->mean;
    p.prstd = pbn->rstd;
  }
  return gaddpg_gemm_nt_impl(&g, 1, amode, emode, st);
}

#define TRY(x)              \
  do {                      \
    int rc__ = (x);         \
    if (rc__) return rc__;  \
  } while (0)

bool check_level(const gaddpg_sa_level* d) {
  if (!d || d->M_max < 0 || d->S < 1 || d->count < 1.0) return false;
  for (int l = 0; l < 3; ++l) {
    const gaddpg_sa_layer& L = d->layer[l];
    const bool raw0 = (l == 0 && d->cloud != nullptr);   // the first level reads the raw (N, K) parameter
    if (!L.W || !L.gamma || !L.beta || !L.scale || !L.shift || !L.mean || !L.rstd || !L.Y || L.N < 4 || (L.N % 4) || L.K < 1 || L.Kp < L.K ||
        (!raw0 && (L.Kp % 4)))
      return false;
    if (l > 0 && L.Kp != d->layer[l - 1].N) return false;
  }
  return d->out && d->arg;
}

}  // namespace

extern "C" {

int gaddpg_sa_struct_sizes(int* sa_layer, int* sa_level) {
  if (sa_layer) *sa_layer = (int)sizeof(gaddpg_sa_layer);
  if (sa_level) *sa_level = (int)sizeof(gaddpg_sa_level);
  return GADDPG_OK;
}

long long gaddpg_sa_level_workspace_bytes(int B, int M_max) {
  return (long long)(stats_floats() * sizeof(float) + gaddpg_gemm_tn_workspace_bytes_impl() + sa1_scratch_bytes(B, M_max));
}

// forward of one level; training != 0: batch statistics (+ running-statistic update), else the running statistics
int gaddpg_sa_forward(const gaddpg_sa_level* d, int training, void* ws, long long ws_bytes, void* stream) {
  GADDPG_CHECK_ARG(check_level(d), "sa_forward: bad level descriptor");
  Ws w;
  GADDPG_CHECK_ARG(carve(ws, ws_bytes, d->B, d->M_max, &w), "sa_forward: workspace too small or misaligned (need %lld bytes)",
                   gaddpg_sa_level_workspace_bytes(d->B, d->M_max));
  if (d->M_max == 0) return GADDPG_OK;
  const gaddpg_sa_layer* L = d->layer;
  float* stats = training ? w.stats : nullptr;
  GADDPG_CUDA(cudaMemsetAsync(w.counter, 0, 16, (cudaStream_t)stream));
  gaddpg_bn_tail t0 = tail_fwd(w, L[0], d->count, training);
  if (d->cloud) {   // first level: conv0 straight from the cloud (grouped input never built)
    GADDPG_CHECK_ARG(L[0].N == 64 && d->row_seg && d->row_src && d->row_w && d->seg_off && d->ctr, "sa_forward: first-level arguments");
    float* bcbias = w.sa1 + (w.sa1_bytes / sizeof(float) - (size_t)d->B * 64);
    TRY(gaddpg_sa1_l1_fwd_impl(d->cloud, d->cloud_stride_b, d->cloud_stride_c, d->skip, d->Cp, d->bc, d->Cb, d->B, d->ctr, d->npoint,
                               d->seg_off, d->row_seg, d->row_src, d->row_w, d->M_max, d->M_dev, L[0].W, L[0].Kp, bcbias, L[0].Y, stats,
                               &t0, stream));
  } else {
    GADDPG_CHECK_ARG(d->G && d->ldg >= L[0].Kp, "sa_forward: input rows");
    TRY(nt1(plain(d->G, d->ldg), L[0].W, L[0].Kp, L[0].Y, L[0].N, d->M_max, d->M_dev, L[0].N, L[0].Kp, stats, d->row_w, nullptr, 0, nullptr,
            OP_PLAIN, EPI_STORE, stream, &t0));
  }
  if (!training) TRY(finalize_eval(w, L[0], d->count, stream));
  for (int l = 1; l < 3; ++l) {
    const gaddpg_bn_tail tl = tail_fwd(w, L[l], d->count, training);
    TRY(nt1(bnrelu(L[l - 1].Y, L[l - 1].N, L[l - 1]), L[l].W, L[l].Kp, L[l].Y, L[l].N, d->M_max, d->M_dev, L[l].N, L[l].Kp, stats, d->row_w,
            nullptr, 0, nullptr, OP_BNRELU, EPI_STORE, stream, &tl));
    if (!training) TRY(finalize_eval(w, L[l], d->count, stream));
  }
  return gaddpg_pool_fwd_impl(L[2].Y, L[2].N, L[2].scale, L[2].shift, d->seg_off, d->fixed_len, d->S, d->out, d->arg, stream);
}

// backward of the same level: dOut (S, ld_dout) -> parameter gradients (want_dw) and dG (generic level) / dbc (first level)
int gaddpg_sa_backward(const gaddpg_sa_level* d, const float* dOut, int ld_dout, int want_dw, int accumulate, float* dG,
                                  int lddg, float* dbc, void* ws, long long ws_bytes, void* stream) {
  GADDPG_CHECK_ARG(check_level(d) && dOut && ld_dout >= d->layer[2].N, "sa_backward: bad arguments");
  Ws w;
  GADDPG_CHECK_ARG(carve(ws, ws_bytes, d->B, d->M_max, &w), "sa_backward: workspace too small or misaligned (need %lld bytes)",
                   gaddpg_sa_level_workspace_bytes(d->B, d->M_max));
  if (d->M_max == 0) return GADDPG_OK;
  const gaddpg_sa_layer* L = d->layer;
  for (int l = 0; l < 3; ++l)
    GADDPG_CHECK_ARG(L[l].D && L[l].bw_g && L[l].bw_m1 && L[l].bw_m2 && (!want_dw || (L[l].dW && L[l].dgamma && L[l].dbeta)) && (l == 0 || L[l].WT),
                     "sa_backward: layer %d lacks backward buffers", l);
  GADDPG_CUDA(cudaMemsetAsync(w.counter, 0, 16, (cudaStream_t)stream));
  const gaddpg_bn_tail t2 = tail_bwd(w, L[2], d->count, want_dw, accumulate);
  TRY(gaddpg_pool_bwd_impl(dOut, ld_dout, d->out, d->arg, L[2].Y, L[2].N, d->row_seg, d->fixed_len, d->M_max, d->M_dev, L[2].mean,
                           L[2].rstd, L[2].D, w.stats, &t2, stream));
  for (int l = 2; l >= 1; --l) {
    const gaddpg_bn_tail tp = tail_bwd(w, L[l - 1], d->count, want_dw, accumulate);   // statistics of the DMASK epilogue below
    const Operand dy = bnbwd(L[l].D, L[l].N, L[l].Y, L[l].N, L[l], d->row_w);
    if (want_dw) {
      TNProblem t = {};
      t.P = dy;
      t.Q = bnrelu(L[l - 1].Y, L[l - 1].N, L[l - 1]);
      t.M_max = d->M_max;
      t.M_dev = d->M_dev;
      t.N = L[l].N;
      t.K = L[l].Kp;
      TRY(gaddpg_gemm_tn_impl(&t, OP_BNBWD, OP_BNRELU, L[l].dW, L[l].K, L[l].N, L[l].K, 0, nullptr, accumulate, w.tn, w.tn_bytes, stream));
    }
    TRY(nt1(dy, L[l].WT, L[l].N, L[l - 1].D, L[l].Kp, d->M_max, d->M_dev, L[l].Kp, L[l].N, w.stats, nullptr, L[l - 1].Y, L[l].Kp, &L[l - 1],
            OP_BNBWD, EPI_DMASK, stream, &tp));
  }
  if (d->cloud) {
    return gaddpg_sa1_l1_bwd_impl(d->cloud, d->cloud_stride_b, d->cloud_stride_c, d->skip, d->Cp, d->bc, d->Cb, d->B, d->ctr, d->npoint,
                                  d->seg_off, d->row_seg, d->row_src, d->row_w, d->M_max, d->M_dev, L[0].D, L[0].Y, L[0].bw_g, L[0].bw_m1,
                                  L[0].bw_m2, L[0].mean, L[0].rstd, L[0].W, L[0].Kp, want_dw ? L[0].dW : nullptr, accumulate,
                                  (d->Cb > 0) ? dbc : nullptr, w.sa1, w.sa1_bytes - (size_t)d->B * 64 * sizeof(float), stream);
  }
  const Operand dy0 = bnbwd(L[0].D, L[0].N, L[0].Y, L[0].N, L[0], d->row_w);
  if (want_dw) {
    TNProblem t = {};
    t.P = dy0;
    t.Q = plain(d->G, d->ldg);
    t.M_max = d->M_max;
    t.M_dev = d->M_dev;
    t.N = L[0].N;
    t.K = L[0].Kp;
    TRY(gaddpg_gemm_tn_impl(&t, OP_BNBWD, OP_PLAIN, L[0].dW, L[0].K, L[0].N, L[0].K, d->rot, nullptr, accumulate, w.tn, w.tn_bytes, stream));
  }
  if (dG) {
    GADDPG_CHECK_ARG(L[0].WT && lddg >= L[0].Kp, "sa_backward: dG needs layer 0's transposed weights");
    TRY(nt1(dy0, L[0].WT, L[0].N, dG, lddg, d->M_max, d->M_dev, L[0].Kp, L[0].N, nullptr, nullptr, nullptr, 0, nullptr, OP_BNBWD, EPI_STORE,
            stream));
  }
  return GADDPG_OK;
}

// SURVEY.md §8(b) name for the fused multi-tensor Adam: L2 weight decay, bias correction from `step`, optional gradient-clip
// coefficient (device scalar), 1/world gradient scale and Polyak target in the same pass (= gaddpg_adam_step).
int gaddpg_adam_fused_step(float* p, float* g, float* m, float* v, long long n, double lr, double beta1, double beta2, double eps,
                                      double weight_decay, long long step, double grad_scale, const float* clip_coef, int write_back_grad,
                                      float* polyak_target, double tau, void* stream) {
  return gaddpg_adam_step_impl(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, nullptr, grad_scale, clip_coef, write_back_grad,
                               polyak_target, tau, stream);
}

}  // extern "C"
