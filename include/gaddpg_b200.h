/*
 * gaddpg_b200.h — C ABI of libgaddpg_b200.so: the B200 (sm_100a) replacement for the native operator
 * boundary under GA-DDPG's offline actor-critic update.
 *
 * What it replaces (reference file:line):
 *   - the pybind module `pointnet2_ops._ext` (third-party, un-vendored; reached through
 *     pointnet2_ops.pointnet2_modules.PointnetSAModule at /root/reference/core/networks.py:66-81 and
 *     pointnet2_utils.{furthest_point_sample,gather_operation} at /root/reference/core/utils.py:795-796):
 *       furthest_point_sampling, gather_points(+_grad), ball_query, group_points(+_grad);
 *   - the ATen/cuDNN/cuBLAS kernels PyTorch launches for the shared MLP, heads, losses, Adam and target
 *     updates of Agent.update_parameters (/root/reference/core/ddpg.py:146-185, agent.py:127-139,192-209,
 *     networks.py:65-92,280-300,339-371, loss.py:17-31, utils.py:750-770,960-1006).
 *
 * Conventions (all entry points):
 *   - plain C types only; every pointer is a DEVICE pointer owned by the caller (no allocation inside);
 *     tensors are dense row-major float32 / int32 unless a stride is an explicit argument;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it and graph-capturable;
 *   - return 0 on success, <0 on error (GADDPG_ERR_*); gaddpg_last_error() returns a thread-local message;
 *   - there is no CPU fallback: without a CUDA device every compute call fails with GADDPG_ERR_CUDA.
 *
 * Row-count convention: kernels that run over the compact (duplicate-folded) row list take `M_max`
 * (buffer capacity, used for the launch shape) and `M_dev`, a device pointer to the live row count
 * (seg_off[S] of gaddpg_row_table); pass M_dev = NULL when the row count is exactly M_max.
 */
#ifndef GADDPG_B200_H
#define GADDPG_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define GADDPG_API __attribute__((visibility("default")))
#else
#define GADDPG_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define GADDPG_OK 0
#define GADDPG_ERR_ARG (-1)
#define GADDPG_ERR_CUDA (-2)
#define GADDPG_ERR_UNSUPPORTED (-3)

/* library identity */
GADDPG_API int gaddpg_version(void);                 /* ABI version, bumped on any signature change */
GADDPG_API long long gaddpg_launch_count(void);      /* kernels launched by this library so far (this process) */
GADDPG_API const char* gaddpg_last_error(void);      /* message of the last failing call on this thread */
GADDPG_API const char* gaddpg_build_info(void);      /* "sm_100a nvcc x.y ..." */
GADDPG_API int gaddpg_device_info(int* sm_count, int* cc_major, int* cc_minor, long long* total_mem);

/* ---- index ops: replace pointnet2_ops._ext (SURVEY.md §8 Spec S1-S3) ------------------------------ */

/* upstream cuda_utils.h opt_n_threads(): the virtual block size that fixes FPS tie-breaking. */
GADDPG_API int gaddpg_opt_n_threads(int work_size);

/* _ext.furthest_point_sampling(xyz[B,N,3], m) -> idx[B,m].  Bit-exact Spec S1. */
GADDPG_API int gaddpg_fps(const float* xyz, int B, int N, int m, int32_t* idx, void* stream);

/* _ext.ball_query(new_xyz[B,m,3], xyz[B,N,3], radius, nsample) -> idx[B,m,nsample]; cnt[B,m] optional
 * (hits found, capped at nsample). Bit-exact Spec S2. */
GADDPG_API int gaddpg_ball_query(const float* new_xyz, const float* xyz, int B, int N, int m, float radius, int nsample,
                      int32_t* idx, int32_t* cnt, void* stream);

/* Fused FPS + centroid gather + ball query for one set-abstraction level: what
 * _PointnetSAModuleBase.forward + QueryAndGroup do with three ops and two transposes.
 * xyz element (b,k,c) is read at xyz[b*stride_b + k*stride_k + c*stride_c], so both the (B,N,3) layout of
 * the extension and the reference's channel-major cloud (B,C,N+6) (xyz + 6, stride_k=1, stride_c=N+6) are
 * consumed in place.  Outputs: fps_idx[B,m], new_xyz[B,m,3], bq_idx[B,m,nsample], bq_cnt[B,m]. */
GADDPG_API int gaddpg_fps_ballquery(const float* xyz, long long stride_b, int stride_k, int stride_c, int B, int N, int m,
                         float radius, int nsample, int32_t* fps_idx, float* new_xyz, int32_t* bq_idx,
                         int32_t* bq_cnt, void* stream);

/* _ext.gather_points / gather_points_grad: pts[B,C,N], idx[B,m] <-> out[B,C,m] (grad is deterministic). */
GADDPG_API int gaddpg_gather_points(const float* pts, const int32_t* idx, int B, int C, int N, int m, float* out, void* stream);
GADDPG_API int gaddpg_gather_points_grad(const float* grad_out, const int32_t* idx, int B, int C, int N, int m, float* grad_pts,
                              void* stream);
/* _ext.group_points / group_points_grad: pts[B,C,N], idx[B,m,s] <-> out[B,C,m,s] (grad is deterministic). */
GADDPG_API int gaddpg_group_points(const float* pts, const int32_t* idx, int B, int C, int N, int m, int s, float* out,
                        void* stream);
GADDPG_API int gaddpg_group_points_grad(const float* grad_out, const int32_t* idx, int B, int C, int N, int m, int s,
                             float* grad_pts, void* stream);

/* Compact row table of one SA level (duplicate folding): S = B*m segments; a segment with h hits owns
 * h' = max(h,1) rows; row 0 carries weight nsample-h'+1.  seg_off[S+1] (seg_off[S] = live row count),
 * row_seg[M], row_src[M] (point index inside the sample), row_w[M]; capacity M >= S*min(nsample, N). */
GADDPG_API int gaddpg_row_table(const int32_t* bq_cnt, const int32_t* bq_idx, int S, int nsample, int32_t* seg_off,
                     int32_t* row_seg, int32_t* row_src, float* row_w, void* stream);

/* ---- row-GEMMs with fused BatchNorm prologues / epilogues -------------------------------------------
 * Replace the cuDNN 1x1 Conv2d + BatchNorm2d + ReLU chain of upstream build_shared_mlp (SA modules built at
 * /root/reference/core/networks.py:65-92), nn.Linear + BatchNorm1d of the FC head (networks.py:84-91) and the
 * nn.Linear layers of QNetwork / GaussianPolicy (networks.py:265-300,315-351), forward and backward.
 *   NT:  C[M,N]  = epi( pro(A)[M,K] . B[N,K]^T )         TN:  dW[N,K] = sum_r pro1(P)[r,:]^T pro2(Q)[r,:]
 * pro: GADDPG_OP_PLAIN x | GADDPG_OP_BNRELU relu(x*c0[c]+c1[c]) |
 *      GADDPG_OP_BNBWD  c0[c]*(X - rw[r]*(c1[c] + (Y-c3[c])*c4[c]*c2[c]))   (c0=gamma*rstd, c1=m1, c2=m2, c3=mean, c4=rstd)
 * epi: GADDPG_EPI_STORE  C = [relu](acc + bias), optional weighted (sum, sum^2) column statistics;
 *      GADDPG_EPI_DMASK  C = acc * [Yprev*psc+psh > 0], optional (sum, sum*xhat_prev) statistics.
 * Statistics land in GADDPG_STAT_SLOTS per-CTA slots ([slots][2][N] floats) and are summed in a fixed order by
 * the finalize calls below (deterministic; no float atomics). */
#define GADDPG_STAT_SLOTS 296
#define GADDPG_MAX_GROUP 4
enum { GADDPG_OP_PLAIN = 0, GADDPG_OP_BNRELU = 1, GADDPG_OP_BNBWD = 2, GADDPG_OP_BNBWD_POOL = 3 };
enum { GADDPG_EPI_STORE = 0, GADDPG_EPI_DMASK = 1 };

typedef struct gaddpg_operand {
  const float* X; int ldx;
  const float* Y; int ldy;
  const float* rw;
  const float* c0; const float* c1; const float* c2; const float* c3; const float* c4;
  /* GADDPG_OP_BNBWD_POOL only: the max-pool gradient is never materialised.  X = E (S, ldx): pooled gradient already
   * masked by the pooled ReLU, pmask (M, ldx/32) uint32: bit c of row r set iff r is the arg-max row of (segment(r), c)
   * (both written by gaddpg_pool_bwd_sparse), pseg (M): segment of every row; the operand is BNBWD with
   * D[r][c] = bit(pmask[r], c) ? E[pseg[r]][c] : 0.  Taken by the tcgen05 kernels only (NT with K = 128, TN); other
   * shapes must use the dense pool_bwd path. */
  const uint32_t* pmask; const int32_t* pseg;
} gaddpg_operand;

/* BatchNorm finalize as the TAIL of the kernel that produced the statistics (kind != 0): the CTA that finishes last (one
 * atomic ticket per CTA on *counter) sums the statistic slots in the fixed slot order and writes what
 * gaddpg_bn_finalize_fwd (kind 1) / gaddpg_bn_finalize_bwd (kind 2) would have written — same arithmetic, one launch fewer per
 * BatchNorm layer and pass.  Replaces the batch-statistics half of torch.nn.BatchNorm2d / BatchNorm1d (training mode) behind
 * upstream build_shared_mlp and /root/reference/core/networks.py:84-91.
 *   kind 1: a = gamma, b = beta, o0..o3 = scale, shift, mean, rstd (o2 / o3 may be NULL); running_mean / running_var / nbt as in
 *           gaddpg_bn_finalize_fwd (momentum == 1: staging mode);
 *   kind 2: a = gamma, b = rstd (forward pass), o0..o2 = g, m1, m2; dgamma / dbeta (may be NULL) (+)= per `accumulate`.
 * counter: one zero-initialised device word per statistics buffer (stream-ordered reuse); the kernel leaves it zero.
 * Kernels without a fused tail run the separate finalize kernel after the producer, so the call has the same effect on
 * every path. */
typedef struct gaddpg_bn_tail {
  int kind; int accumulate;
  double count;
  const float* a; const float* b;
  float eps, momentum;
  float* running_mean; float* running_var; long long* num_batches_tracked;
  float* o0; float* o1; float* o2; float* o3;
  float* dgamma; float* dbeta;
  unsigned int* counter;
} gaddpg_bn_tail;

/* The other statistics producers (gaddpg_sa1_l1_fwd, gaddpg_pool_bwd, gaddpg_pool_bwd_sparse, gaddpg_dmask_stats) take the same
 * descriptor as a trailing `const gaddpg_bn_tail* tail` argument (a HOST pointer, read during the call; NULL = no tail). */

typedef struct gaddpg_nt_problem {
  gaddpg_operand A;
  const float* Bw; int ldb;
  const float* bias;
  float* C; int ldc;
  int M_max; const int* M_dev;
  int N, K;
  int relu;
  float* stats; const float* srw;
  const float* Yprev; int ldyp;
  const float* psc; const float* psh; const float* pmean; const float* prstd;
  /* GADDPG_EPI_STORE on the tcgen05 kernels only — max-pool fused into the last shared-MLP layer (upstream
   * F.max_pool2d(kernel=[1,nsample]) after build_shared_mlp): when pool_keys != NULL the epilogue also reduces the raw
   * conv output per (segment, column) to its extreme value and the row that attains it, as one 64-bit key
   * pool_keys[seg][col] (atomicMax on order-preserving keys: deterministic; ties keep the lowest row).  BatchNorm is
   * monotone per channel and the sign of its scale is the sign of pool_gamma[col] (the BatchNorm weight), so
   * relu(bn(max)) (gamma >= 0) or relu(bn(min)) (gamma < 0) IS the pooled activation: gaddpg_pool_keys_finalize turns the
   * keys into out / arg once the batch statistics are known and re-zeroes them.
   * pool_seg (M): segment of every row.  no_store: do not write C at all (passes that never run backward). */
  unsigned long long* pool_keys; const int32_t* pool_seg; const float* pool_gamma; int no_store;
  /* optional: dense (N, K) images of Bw split for 3xTF32 (hi = x & 0xFFFFE000, lo = x - hi; gaddpg_wprep_batched split jobs).
   * When both are given the tcgen05 whole-K kernel fetches the weights with TMA (cp.async.bulk.tensor) instead of
   * splitting Bw in every CTA. */
  const float* Bw_hi; const float* Bw_lo;
  /* optional BatchNorm finalize of `stats` as the tail of this launch (see gaddpg_bn_tail; kind 0 = none) */
  gaddpg_bn_tail tail;
} gaddpg_nt_problem;

typedef struct gaddpg_nt_group { gaddpg_nt_problem p[GADDPG_MAX_GROUP]; } gaddpg_nt_group;

typedef struct gaddpg_tn_problem {
  gaddpg_operand P; gaddpg_operand Q;
  int M_max; const int* M_dev;
  int N, K;
} gaddpg_tn_problem;

/* sizeof() of the three structs above, for binding self-checks */
GADDPG_API int gaddpg_struct_sizes(int* operand, int* nt_problem, int* tn_problem);

/* Wide layers (M >= 8192 rows, K in {32,64,96,128}, N <= 256) run on tcgen05 tensor cores with a 3xTF32 split
 * (FP32-level accuracy); 0 forces every NT product onto the FP32 FFMA kernel (also: env GADDPG_TC=0). */
GADDPG_API int gaddpg_set_tensor_core(int enable);
GADDPG_API int gaddpg_get_tensor_core(void);
/* up to GADDPG_MAX_GROUP independent NT problems in one launch (blockIdx.y = problem) */
GADDPG_API int gaddpg_gemm_nt(const gaddpg_nt_group* group, int nprob, int amode, int emode, void* stream);
/* which kernel gaddpg_gemm_nt would launch for this group: 0 = FP32 FFMA (gemm_nt_kernel), 1 = tcgen05 whole-K
 * (tc_gemm_nt_kernel, the SA1 layers), 2 = tcgen05 K-chunked (tc_nt_kc_kernel), 3 = mma.sync 3xTF32 small-M kernel
 * (skinny_nt_kernel: FC head, actor / critic layers); < 0 = argument error.  Host-only query
 * used by bench.py to attribute measured time to kernels. */
GADDPG_API int gaddpg_gemm_nt_path(const gaddpg_nt_group* group, int nprob, int amode, int emode);
/* dW[n][(k+rot) % Ktrue] (+)= TN product for n < Ntrue (rows >= Ntrue and columns >= Ktrue are padding);
 * dbias[n] (+)= column sums of P */
GADDPG_API long long gaddpg_gemm_tn_workspace_bytes(void);
GADDPG_API int gaddpg_gemm_tn(const gaddpg_tn_problem* prob, int pmode, int qmode, float* dW, int ldd, int Ntrue, int Ktrue,
                              int rot, float* dbias, int accumulate, float* ws, long long ws_bytes, void* stream);

/* BatchNorm finalize.  fwd: slots -> batch mean / biased var -> scale, shift, mean, rstd and PyTorch's running-stat
 * update (momentum, unbiased var, num_batches_tracked += 1); training=0 uses the running stats (eval mode).
 * bwd: slots -> g = gamma*rstd, m1, m2 and dgamma / dbeta. */
GADDPG_API int gaddpg_bn_finalize_fwd(const float* stats, int C, double count, const float* gamma, const float* beta, float eps,
                                      float momentum, float* running_mean, float* running_var, long long* num_batches_tracked,
                                      int training, float* scale, float* shift, float* mean_out, float* rstd_out, void* stream);
GADDPG_API int gaddpg_bn_finalize_bwd(const float* stats, int C, double count, const float* gamma, const float* rstd, float* g,
                                      float* m1, float* m2, float* dgamma, float* dbeta, int accumulate, void* stream);
/* Deferred running-statistics update for encoder passes that run concurrently on several streams: each pass stages its
 * batch mean / unbiased variance (gaddpg_bn_finalize_fwd with momentum = 1 writes them verbatim into a staging copy of the
 * running-stat arena), and this applies  running = (1-momentum)*running + momentum*staged  over the whole arena (n floats)
 * plus num_batches_tracked[0..n_layers) += 1 — called once per pass, in the reference's pass order
 * (torch.nn.BatchNorm2d / BatchNorm1d training-mode forward, reached from core/networks.py:65-92). */
GADDPG_API int gaddpg_bn_running_update(float* running, const float* staged, long long n, float momentum,
                                        long long* num_batches_tracked, int n_layers, void* stream);

/* ---- set-abstraction glue (SURVEY.md §8 Spec S3; upstream QueryAndGroup / GroupAll / F.max_pool2d) ------------- */
/* SA1 first layer straight from the channel-major cloud (B, *, skip+N): W[64][3+Cp+Cb] = [dxyz | per-point | broadcast];
 * bc (B,Cb) are per-sample constant channels (the action, utils.py:291-297); bcbias_ws (B,64) scratch;
 * seg_off[B*npoint+1] of gaddpg_row_table (rows of one sample are contiguous: work is split per sample). */
GADDPG_API int gaddpg_sa1_l1_fwd(const float* cloud, long long cloud_stride_b, int cloud_stride_c, int skip, int Cp, const float* bc,
                                 int Cb, int B, const float* ctr, int npoint, const int32_t* seg_off, const int32_t* row_seg,
                                 const int32_t* row_src, const float* row_w, int M_max, const int* M_dev, const float* W, int ldw,
                                 float* bcbias_ws, float* Y, float* stats, const gaddpg_bn_tail* tail, void* stream);
/* backward of the same layer: dW (may be NULL), dbc (B,Cb) (may be NULL); dY1 is never materialised.
 * ws: (GADDPG_STAT_SLOTS*64*16 + B*64*(1 + ceil(M_max/B/256))) floats */
GADDPG_API int gaddpg_sa1_l1_bwd(const float* cloud, long long cloud_stride_b, int cloud_stride_c, int skip, int Cp, const float* bc,
                                 int Cb, int B, const float* ctr, int npoint, const int32_t* seg_off, const int32_t* row_seg,
                                 const int32_t* row_src, const float* row_w, int M_max, const int* M_dev, const float* D,
                                 const float* Y, const float* g, const float* m1, const float* m2, const float* mean,
                                 const float* rstd, const float* W, int ldw, float* dW, int accumulate, float* dbc, float* ws,
                                 long long ws_bytes, void* stream);
/* ---- SA1 shared MLP as one TMA-fed, recompute-instead-of-store tcgen05 chain (csrc/sa1_fused.cu) -------------------------
 * Replaces QueryAndGroup + build_shared_mlp (3 x [Conv2d 1x1, train-mode BatchNorm2d, ReLU]) + F.max_pool2d of the first
 * set-abstraction module (/root/reference/core/networks.py:66-71) for the compact row list.  Three phases = three calls; phase
 * k recomputes layers < k from the cloud and writes the batch-statistic slots of layer k-1's output (finalise each with
 * gaddpg_bn_finalize_fwd before the next phase).  Phase 3 also writes, per (ball group s, channel c), the extreme of the last
 * pre-BN output (max for gamma2[c] >= 0, else min) and its row; gaddpg_sa1_pool_finalize turns them into
 * out = relu(bn(extreme)) / arg.  Ykeep: NULL for forward-only passes (no activation touches HBM) or the (M_max, 64 | 64 | 128)
 * buffer that receives this phase's pre-BN output for the backward kernels (TMA stores).
 * wsplit: gaddpg_sa1f_wsplit_floats() floats written by gaddpg_sa1f_wprep (hi/lo TF32 split of the three conv weights; redo
 * after every optimiser step).  part_ext/part_arg: (ceil(M_max / 128) + 1, 128), one row per 128-row tile + a spare; seg_part: (S) int32, all -1 on entry
 * (the finalize call restores that).  Input channels: [dxyz(3) | cloud rows 0..Cp-1 | bc (B,Cb) broadcast], 3+Cp+Cb <= 16. */
GADDPG_API long long gaddpg_sa1f_wsplit_floats(void);
GADDPG_API int gaddpg_sa1_fused_grid(int M_max);
GADDPG_API int gaddpg_sa1f_wprep(const float* W0, int ld0, int K1, const float* W1, const float* W2, float* wsplit, void* stream);
GADDPG_API int gaddpg_sa1_fused_fwd(int phase, const float* cloud, long long cloud_stride_b, int cloud_stride_c, int skip, int Cp,
                                    const float* bc, int Cb, const float* ctr, int npoint, const int32_t* seg_off,
                                    const int32_t* row_seg, const int32_t* row_src, const float* row_w, int M_max, const int* M_dev,
                                    const float* wsplit, const float* sc0, const float* sh0, const float* sc1, const float* sh1,
                                    const float* gamma2, float* stats, float* Ykeep, float* ext, int32_t* arg, float* part_ext,
                                    int32_t* part_arg, int32_t* seg_part, void* stream);
GADDPG_API int gaddpg_sa1_pool_finalize(const float* ext, const int32_t* arg, const float* part_ext, const int32_t* part_arg,
                                        int32_t* seg_part, const float* gamma, const float* scale, const float* shift, int S,
                                        float* out, int32_t* arg_out, void* stream);
/* G[r] = [feats[src] (C) | xyz[src]-ctr[seg] (3) | 0-pad]; row tables NULL: identity rows, absolute xyz (GroupAll) */
GADDPG_API int gaddpg_gather_rows(const float* feats, int C, const float* xyz, int n_src, const float* ctr, int npoint,
                                  const int32_t* row_seg, const int32_t* row_src, int M_max, const int* M_dev, float* G, int ldg,
                                  void* stream);
/* deterministic group_points_grad over the compact rows: dfeats (B, n_src, C) */
GADDPG_API int gaddpg_scatter_rows(const float* dG, int ldg, int C, int B, int n_src, int npoint, const int32_t* seg_off,
                                   const int32_t* row_src, float* dfeats, void* stream);
/* out[seg][c] = max_r relu(Y[r][c]*scale[c]+shift[c]) over the segment, arg = first arg-max row */
GADDPG_API int gaddpg_pool_fwd(const float* Y, int C, const float* scale, const float* shift, const int32_t* seg_off, int fixed_len,
                               int S, float* out, int32_t* arg, void* stream);
GADDPG_API int gaddpg_pool_bwd(const float* dOut, int ld_dout, const float* out, const int32_t* arg, const float* Y, int C,
                               const int32_t* row_seg, int fixed_len, int M_max, const int* M_dev, const float* mean,
                               const float* rstd, float* D, float* stats, const gaddpg_bn_tail* tail, void* stream);
/* Sparse form of the max-pool backward for levels whose layer-3 operand is consumed as GADDPG_OP_BNBWD_POOL:
 * E[s][c] = out[s][c] > 0 ? dOut[s][c] : 0, the arg-max bit mask (M_max, C/32) (zeroed here, then one bit per (s, c)) and
 * the BatchNorm-backward sums (sum D, sum D*xhat) of the implied dense D, gathered from the S*C arg-max elements of Y only
 * (the dense kernel reads all of Y and writes all of D).  C in {64,128,256} (mask needs C % 32 == 0).
 * Replaces the autograd backward of F.max_pool2d(kernel=[1,nsample]) + ReLU in upstream _PointnetSAModuleBase.forward. */
GADDPG_API int gaddpg_pool_bwd_sparse(const float* dOut, int ldo, const float* out, const int32_t* arg, const float* Y, int C,
                                      int S, const float* mean, const float* rstd, float* E, uint32_t* mask, int M_max,
                                      float* stats, const gaddpg_bn_tail* tail, void* stream);
/* Second half of the fused max-pool (see gaddpg_nt_problem.pool_keys): out[s][c] = relu(v*scale[c]+shift[c]) with
 * v = the extreme value held by keys[s][c] (max for gamma[c] >= 0, min otherwise), arg[s][c] = the row attaining it;
 * every key is reset to 0 (= empty) for the next pass.  keys must be zero-initialised once by the caller. */
GADDPG_API int gaddpg_pool_keys_finalize(unsigned long long* keys, int S, int C, const float* gamma, const float* scale,
                                         const float* shift, float* out, int32_t* arg, void* stream);
/* feat[b] = [relu(Y*scale+shift) (C) | time[b]+time_offset | 0-pad to ld]  (ddpg.py:56-57 appends the time column) */
GADDPG_API int gaddpg_feat_finish(const float* Y, int C, const float* scale, const float* shift, const float* time,
                                  float time_offset, int B, float* feat, int ld, void* stream);

/* ---- per-level composites (sa_forward_train / sa_forward_eval / sa_backward, adam_fused_step) -----------
 * One set-abstraction level = upstream PointnetSAModule.forward (shared MLP of three [Conv2d 1x1, BatchNorm2d, ReLU] over the
 * grouped points, then F.max_pool2d over each ball group; /root/reference/core/networks.py:66-81) and its backward, as ONE call
 * each: fixed launch sequences over the kernels above, written in C++ (csrc/composite.cu) so that a non-Python host does not
 * have to re-implement the sequencing of ga-ddpg_b200/engine.py.  Everything is caller-owned device memory.
 *
 * gaddpg_sa_layer: one conv + BatchNorm.  W: forward operand (N, Kp) row-major, Kp = K padded to a multiple of 4 (the first
 * level reads the raw (N, K) parameter instead: Kp = K there); WT: (Kp, N) transposed copy for dX (layers 1, 2; layer 0 of a
 * generic level when dG is wanted) — gaddpg_wprep_batched writes both layouts.  Y: (M_max, N) pre-BatchNorm output kept for the
 * backward.  scale/shift/mean/rstd: (N) BatchNorm constants of the pass (outputs of the forward).  D, bw_*: backward scratch
 * ((M_max, N) and (N) x 3); dW: (N, K) gradient, dgamma/dbeta: (N).
 * gaddpg_sa_level: generic level: G (M_max, ldg) input rows [feats | rel. xyz | 0-pad] from gaddpg_gather_rows, rot = column
 * rotation of conv0's weight relative to G (3 when G is [feats | xyz]).  First level: cloud != NULL (see gaddpg_sa1_l1_fwd for
 * cloud / bc / ctr), layer widths 64, 64, 128.  Row tables from gaddpg_row_table (NULL + fixed_len for GroupAll: SA3).
 * count = B * npoint * nsample (rows of the dense grouped tensor the BatchNorm statistics are taken over). */
typedef struct gaddpg_sa_layer {
  const float* W; const float* WT; int N, K, Kp;
  const float* gamma; const float* beta; float* running_mean; float* running_var; long long* num_batches_tracked;
  float* scale; float* shift; float* mean; float* rstd;
  float* Y;
  float* D; float* bw_g; float* bw_m1; float* bw_m2;
  float* dW; float* dgamma; float* dbeta;
} gaddpg_sa_layer;

typedef struct gaddpg_sa_level {
  gaddpg_sa_layer layer[3];
  int B, S, M_max; const int* M_dev; double count;
  const int32_t* seg_off; const int32_t* row_seg; const int32_t* row_src; const float* row_w; int fixed_len;
  const float* G; int ldg; int rot;
  const float* cloud; long long cloud_stride_b; int cloud_stride_c, skip, Cp; const float* bc; int Cb; const float* ctr; int npoint;
  float* out; int32_t* arg;
} gaddpg_sa_level;

GADDPG_API int gaddpg_sa_struct_sizes(int* sa_layer, int* sa_level);   /* sizeof() of the two structs above, for binding self-checks */
GADDPG_API long long gaddpg_sa_level_workspace_bytes(int B, int M_max);
/* training != 0: batch statistics + PyTorch's running-statistic update (momentum 0.1, unbiased variance); 0: eval mode */
GADDPG_API int gaddpg_sa_forward(const gaddpg_sa_level* level, int training, void* ws, long long ws_bytes, void* stream);
/* dOut (S, ld_dout): gradient w.r.t. the pooled output.  want_dw: parameter gradients (accumulate: += instead of =); dG
 * (M_max, lddg): gradient w.r.t. the input rows of a generic level (may be NULL); dbc (B, Cb): gradient w.r.t. the broadcast
 * channels of the first level (may be NULL). */
GADDPG_API int gaddpg_sa_backward(const gaddpg_sa_level* level, const float* dOut, int ld_dout, int want_dw, int accumulate, float* dG,
                                  int lddg, float* dbc, void* ws, long long ws_bytes, void* stream);
/* §8(b) adam_fused_step: flat-arena Adam with L2 weight decay, bias correction from `step`, optional gradient-clip coefficient
 * (device scalar), 1/world gradient scale and Polyak target update in the same pass (replaces torch.optim.Adam.step +
 * clip_grad_norm_'s scaling + soft_update, utils.py:750-754,969-970; ddpg.py:141-143). */
GADDPG_API int gaddpg_adam_fused_step(float* p, float* g, float* m, float* v, long long n, double lr, double beta1, double beta2,
                                      double eps, double weight_decay, long long step, double grad_scale, const float* clip_coef,
                                      int write_back_grad, float* polyak_target, double tau, void* stream);

/* ---- heads, TD3 target, losses (networks.py:339-371; ddpg.py:61-88,119-130,170-177; agent.py:127-139; loss.py:17-31) -- */
GADDPG_API int gaddpg_heads_init(const float* act_scale, const float* act_bias, const float* cp_rotz); /* HOST pointers */
GADDPG_API int gaddpg_policy_head_fwd(const float* raw, int ldr, int B, float* pi, void* stream);
GADDPG_API int gaddpg_td3_next_action(const float* raw_t, int ldr, const float* u, float noise_scale, int B, float* next_action,
                                      void* stream);
GADDPG_API int gaddpg_td3_target(const float* qa, int ldq, int oq2, const float* reward, const float* done, float gamma, int B,
                                 float* y, void* stream); /* y = r + (1-done)*gamma*min(qa[:,0], qa[:,oq2]) */
/* qa/dqa [B,ldq]: q1 at column 0, q2 at column oq2, aux_raw(7) at column oaux;
 * out[0]=critic_loss out[1]=critic_grasp_aux_loss out[2]=#(return>0) */
GADDPG_API int gaddpg_critic_loss(const float* qa, int ldq, int oq2, int oaux, const float* y, const float* perturb_flag, const float* ret,
                                  const float* goal, int use_aux, int B, float grad_scale, float* dqa, float* out, void* stream);
/* praw/dpraw [B,ldr] = [mean(6) | extra | ...]; out[0]=bc_loss*bc_weight out[1]=policy_grasp_aux_loss out[2]=#(return>0) */
GADDPG_API int gaddpg_actor_loss(const float* praw, int ldr, const float* pi, const float* expert_action, const float* expert_flag,
                                 const float* ret, const float* goal, int use_aux, float bc_weight, const float* dpi_ac, int B,
                                 float grad_scale, float* dpraw, int n_head, float* out, void* stream);
GADDPG_API int gaddpg_actor_critic_loss(const float* qa, int ldq, int oq2, const float* ret, const float* expert_flag, float mix, int B,
                                        float grad_scale, int n_head, float* dqa, float* out, void* stream);
GADDPG_API int gaddpg_quat_head(const float* raw, int ldr, int B, float* out7, void* stream);
GADDPG_API int gaddpg_policy_sample(const float* raw, int ldr, int off_logstd, const float* eps, int B, float* action, float* logp,
                                    void* stream);

/* ---- optimiser / target networks / statistics (utils.py:750-770,960-1006,92-108; ddpg.py:141; agent.py:221-222) ------ */
/* torch.optim.Adam (L2 weight decay, not AdamW) on one arena segment; clip: device scalar multiplied into the gradient
 * (clip_grad_norm_), grad_scale: 1/world_size; target != NULL fuses target = target*(1-tau) + p*tau. */
GADDPG_API int gaddpg_adam_step(float* p, float* g, float* m, float* v, long long n, double lr, double beta1, double beta2,
                                double eps, double weight_decay, long long step, const float* dyn, double grad_scale,
                                const float* clip, int write_back_grad, float* target, double tau, void* stream);
/* dyn (optional, device): {lr/(1-beta1^t), sqrt(1-beta2^t)} read at run time, so one captured graph serves every step */
GADDPG_API int gaddpg_polyak(float* target, const float* source, long long n, double tau, void* stream);
/* target = target*(1-tau_vec[i]) + source*tau_vec[i]: half_soft_update / half_hard_update as one masked launch */
GADDPG_API int gaddpg_polyak_vec(float* target, const float* source, const float* tau_vec, long long n, void* stream);
GADDPG_API int gaddpg_absmax(const float* x, long long n, float* out, float* ws, void* stream);        /* ws: 1184 floats */
GADDPG_API int gaddpg_clip_coef(const float* g, long long n, float max_norm, float* coef_out, float* norm_out, float* ws,
                                void* stream);
/* derived layouts of a weight W[N][K]: Wp[N][ldp] (columns rotated by rot, zero padded), WT[ldp][ldt] = Wp^T */
GADDPG_API int gaddpg_wprep(const float* W, int N, int K, int rot, float* Wp, int ldp, float* WT, int ldt, void* stream);
/* all derived layouts of a network in one launch: jobs_dev[j] = {W, N, K, rot, Wp, ldp, WT, ldt} as 8 x int64 (device) */
GADDPG_API int gaddpg_wprep_batched(const long long* jobs_dev, int njobs, void* stream);
/* One launch for a whole optimiser phase (replaces the per-network torch.optim.Adam.step / soft_update / module_max_* calls of
 * /root/reference/core/agent.py:192-209,242-259 and ddpg.py:141-143): jobs_dev = njobs records of 120 bytes
 *   { float *p, *g, *m, *v, *target; const float *tau_vec, *dyn, *clip; float *absmax_p, *absmax_g; long long n, chunk0;
 *     float eps, weight_decay, tau; int kind, write_back, pad; }
 * kind 0: Adam with L2 weight decay on [p, p+n) (dyn = {lr/(1-beta1^t), sqrt(1-beta2^t)} on the device, clip = optional device
 * scalar multiplied into the gradient, write_back = store the clipped gradient), then target = target*(1-tau) + p*tau if target;
 * kind 1: that Polyak update only (per-element tau_vec if given); kind 2: statistics only.  absmax_p / absmax_g (optional):
 * device floats that receive max |p| (after the update) / max |g| through an order-independent atomic max of the bit patterns
 * (they must hold a non-negative value, normally 0, on entry).  chunk0 = first 4096-element chunk of the job in the launch's
 * chunk space (running sum of ceil(n / 4096)); total_chunks = that sum over all jobs.  n and every pointer 16-byte aligned. */
GADDPG_API int gaddpg_optim_multi(const void* jobs_dev, int njobs, int total_chunks, void* stream);
/* stand-alone EPI_DMASK: D = dX*[Yprev*psc+psh > 0] plus its BN-backward sums (for gradients arriving from autograd) */
GADDPG_API int gaddpg_dmask_stats(const float* dX, int ldx, const float* Yprev, int C, int M, const float* psc, const float* psh,
                                  const float* pmean, const float* prstd, float* D, float* stats, const gaddpg_bn_tail* tail,
                                  void* stream);
GADDPG_API int gaddpg_f64_to_f32(const double* src, float* dst, long long n, void* stream);

/* ---- device-resident replay buffer -------------------------------------------------------
 * Replaces the producer of update_parameters' input dict when the buffer lives in HBM: the numpy fancy indexing of
 * BaseMemory.__getitem__ / post_process_batch (/root/reference/core/replay_memory.py:109-127,251-272) and the
 * float64->float32 + H2D staging of Agent.prepare_data (/root/reference/core/agent.py:211-240).
 *   cloud_store[capacity][row_floats]  one stored cloud per row ((C, N+6) flattened, float32 of the reference's float64)
 *   rec_store[capacity][rec_width]     the small per-transition fields packed as one float32 record (<= 32 columns);
 *                                      ts_col is the `timestep` column; may be NULL (clouds only)
 *   episode_map[capacity]              index of the last transition of the episode a slot belongs to (uint32 bits)
 *   idx[B]                             sampled slots (batch_idx); out-of-range values are clamped into the store
 * Outputs: state_out[B][row_floats] = cloud_store[idx]; next_out[B][row_floats] = cloud_store[inc] with
 * inc = min(episode_map[idx], idx + 1); rec_out[B][2*rec_width] = [rec_store[idx] | rec_store[inc]] where the timestep
 * column of the first half holds the remaining time (timestep[episode_map[idx]] + 1) - timestep[idx]; inc_out[B]
 * (optional) = inc.  soa_map (optional; 2*rec_width device ints base[c], stride[c], base < 0 = skip) additionally
 * scatters column c of the CURRENT record of sample b to soa_out[base[c] + b*stride[c]] — the field-major vectors
 * Agent.prepare_data keeps on the device — so a minibatch lands in the update's input buffers with no per-field copies
 * (rec_out may then be NULL).  Bit-exact against the reference (byte movement plus one float32 add/sub). */
GADDPG_API int gaddpg_replay_gather(const float* cloud_store, long long row_floats, const float* rec_store, int rec_width, int ts_col,
                                    const int32_t* episode_map, long long capacity, const int32_t* idx, int B, float* state_out,
                                    float* next_out, float* rec_out, int32_t* inc_out, const int32_t* soa_map, float* soa_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GADDPG_B200_H */
