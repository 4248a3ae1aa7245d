// capi.cu — extern "C" surface of libgaddpg_b200.so (declared in include/gaddpg_b200.h).
#include <stdarg.h>
#include <string.h>

#include "../../include/gaddpg_b200.h"
#include "common.cuh"
#include "impl.h"

#define GADDPG_ABI_VERSION 4

static thread_local char g_err[512] = "";
long long g_gaddpg_launches = 0;

void gaddpg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" {

int gaddpg_version(void) { return GADDPG_ABI_VERSION; }
long long gaddpg_launch_count(void) { return g_gaddpg_launches; }
const char* gaddpg_last_error(void) { return g_err; }
const char* gaddpg_build_info(void) {
  return "gaddpg_b200 sm_100a, nvcc " GADDPG_STR(__CUDACC_VER_MAJOR__) "." GADDPG_STR(__CUDACC_VER_MINOR__);
}
int gaddpg_device_info(int* sm_count, int* cc_major, int* cc_minor, long long* total_mem) {
  int dev = 0;
  GADDPG_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  GADDPG_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (total_mem) *total_mem = (long long)p.totalGlobalMem;
  return GADDPG_OK;
}

int gaddpg_fps(const float* xyz, int B, int N, int m, int32_t* idx, void* stream) {
  return gaddpg_fps_ballquery_impl(xyz, (long long)N * 3, 3, 1, B, N, m, idx, nullptr, 0, 0.f, 0, nullptr, nullptr, stream);
}
int gaddpg_ball_query(const float* new_xyz, const float* xyz, int B, int N, int m, float radius, int nsample,
                      int32_t* idx, int32_t* cnt, void* stream) {
  return gaddpg_ball_query_impl(xyz, (long long)N * 3, 3, 1, B, N, m, new_xyz, radius, nsample, idx, cnt, stream);
}
int gaddpg_fps_ballquery(const float* xyz, long long stride_b, int stride_k, int stride_c, int B, int N, int m,
                         float radius, int nsample, int32_t* fps_idx, float* new_xyz, int32_t* bq_idx,
                         int32_t* bq_cnt, void* stream) {
  return gaddpg_fps_ballquery_impl(xyz, stride_b, stride_k, stride_c, B, N, m, fps_idx, new_xyz, 1, radius, nsample,
                                   bq_idx, bq_cnt, stream);
}
int gaddpg_gather_points(const float* pts, const int32_t* idx, int B, int C, int N, int m, float* out, void* stream) {
  return gaddpg_gather_points_impl(B, C, N, m, pts, idx, out, stream);
}
int gaddpg_gather_points_grad(const float* grad_out, const int32_t* idx, int B, int C, int N, int m, float* grad_pts,
                              void* stream) {
  return gaddpg_gather_points_grad_impl(B, C, N, m, grad_out, idx, grad_pts, stream);
}
int gaddpg_group_points(const float* pts, const int32_t* idx, int B, int C, int N, int m, int s, float* out,
                        void* stream) {
  return gaddpg_group_points_impl(B, C, N, m, s, pts, idx, out, stream);
}
int gaddpg_group_points_grad(const float* grad_out, const int32_t* idx, int B, int C, int N, int m, int s,
                             float* grad_pts, void* stream) {
  return gaddpg_group_points_grad_impl(B, C, N, m, s, grad_out, idx, grad_pts, stream);
}
int gaddpg_row_table(const int32_t* bq_cnt, const int32_t* bq_idx, int S, int nsample, int32_t* seg_off,
                     int32_t* row_seg, int32_t* row_src, float* row_w, void* stream) {
  return gaddpg_row_table_impl(S, nsample, bq_cnt, bq_idx, seg_off, row_seg, row_src, row_w, stream);
}

int gaddpg_struct_sizes(int* operand, int* nt_problem, int* tn_problem) {
  if (operand) *operand = (int)sizeof(gaddpg_operand);
  if (nt_problem) *nt_problem = (int)sizeof(gaddpg_nt_problem);
  if (tn_problem) *tn_problem = (int)sizeof(gaddpg_tn_problem);
  return GADDPG_OK;
}
int gaddpg_set_tensor_core(int enable) {
  gaddpg_set_tensor_core_impl(enable);
  return GADDPG_OK;
}
int gaddpg_get_tensor_core(void) { return gaddpg_get_tensor_core_impl(); }
int gaddpg_gemm_nt(const gaddpg_nt_group* group, int nprob, int amode, int emode, void* stream) {
  return gaddpg_gemm_nt_impl(group, nprob, amode, emode, stream);
}
int gaddpg_gemm_nt_path(const gaddpg_nt_group* group, int nprob, int amode, int emode) {
  if (!group || nprob < 1 || nprob > GADDPG_MAX_GROUP) return GADDPG_ERR_ARG;
  const int tc = gaddpg_get_tensor_core_impl();
  if (nprob == 1 && tc >= 1 && tc <= 3 && gaddpg_tc_gemm_supported(group->p[0], amode, emode)) return 1;
  if (nprob == 1 && tc >= 2 && gaddpg_tc_nt_kc_supported(group->p[0], amode, emode)) return 2;
  if (gaddpg_skinny_enabled() && gaddpg_skinny_supported(*group, nprob, amode, emode)) return 3;
  return 0;
}
long long gaddpg_gemm_tn_workspace_bytes(void) { return (long long)gaddpg_gemm_tn_workspace_bytes_impl(); }
int gaddpg_gemm_tn(const gaddpg_tn_problem* prob, int pmode, int qmode, float* dW, int ldd, int Ntrue, int Ktrue, int rot,
                   float* dbias, int accumulate, float* ws, long long ws_bytes, void* stream) {
  return gaddpg_gemm_tn_impl(prob, pmode, qmode, dW, ldd, Ntrue, Ktrue, rot, dbias, accumulate, ws, (size_t)ws_bytes, stream);
}
int gaddpg_bn_finalize_fwd(const float* stats, int C, double count, const float* gamma, const float* beta, float eps,
                           float momentum, float* running_mean, float* running_var, long long* num_batches_tracked, int training,
                           float* scale, float* shift, float* mean_out, float* rstd_out, void* stream) {
  return gaddpg_bn_finalize_fwd_impl(stats, C, count, gamma, beta, eps, momentum, running_mean, running_var,
                                     num_batches_tracked, training, scale, shift, mean_out, rstd_out, stream);
}
int gaddpg_bn_finalize_bwd(const float* stats, int C, double count, const float* gamma, const float* rstd, float* g, float* m1,
                           float* m2, float* dgamma, float* dbeta, int accumulate, void* stream) {
  return gaddpg_bn_finalize_bwd_impl(stats, C, count, gamma, rstd, g, m1, m2, dgamma, dbeta, accumulate, stream);
}
int gaddpg_bn_running_update(float* running, const float* staged, long long n, float momentum, long long* num_batches_tracked,
                             int n_layers, void* stream) {
  return gaddpg_bn_running_update_impl(running, staged, n, momentum, num_batches_tracked, n_layers, stream);
}
int gaddpg_sa1_l1_fwd(const float* cloud, long long cloud_stride_b, int cloud_stride_c, int skip, int Cp, const float* bc, int Cb,
                      int B, const float* ctr, int npoint, const int32_t* seg_off, const int32_t* row_seg, const int32_t* row_src,
                      const float* row_w, int M_max, const int* M_dev, const float* W, int ldw, float* bcbias_ws, float* Y,
                      float* stats, const gaddpg_bn_tail* tail, void* stream) {
  return gaddpg_sa1_l1_fwd_impl(cloud, cloud_stride_b, cloud_stride_c, skip, Cp, bc, Cb, B, ctr, npoint, seg_off, row_seg, row_src,
                                row_w, M_max, M_dev, W, ldw, bcbias_ws, Y, stats, tail, stream);
}
int gaddpg_sa1_l1_bwd(const float* cloud, long long cloud_stride_b, int cloud_stride_c, int skip, int Cp, const float* bc, int Cb,
                      int B, const float* ctr, int npoint, const int32_t* seg_off, const int32_t* row_seg, const int32_t* row_src,
                      const float* row_w, int M_max, const int* M_dev, const float* D, const float* Y, const float* g,
                      const float* m1, const float* m2, const float* mean, const float* rstd, const float* W, int ldw, float* dW,
                      int accumulate, float* dbc, float* ws, long long ws_bytes, void* stream) {
  return gaddpg_sa1_l1_bwd_impl(cloud, cloud_stride_b, cloud_stride_c, skip, Cp, bc, Cb, B, ctr, npoint, seg_off, row_seg, row_src,
                                row_w, M_max, M_dev, D, Y, g, m1, m2, mean, rstd, W, ldw, dW, accumulate, dbc, ws,
                                (size_t)ws_bytes, stream);
}
long long gaddpg_sa1f_wsplit_floats(void) { return gaddpg_sa1f_wsplit_floats_impl(); }
int gaddpg_sa1_fused_grid(int M_max) { return gaddpg_sa1_fused_grid_impl(M_max); }
int gaddpg_sa1f_wprep(const float* W0, int ld0, int K1, const float* W1, const float* W2, float* wsplit, void* stream) {
  return gaddpg_sa1f_wprep_impl(W0, ld0, K1, W1, W2, wsplit, stream);
}
int gaddpg_sa1_fused_fwd(int phase, const float* cloud, long long cloud_stride_b, int cloud_stride_c, int skip, int Cp, const float* bc,
                         int Cb, const float* ctr, int npoint, const int32_t* seg_off, const int32_t* row_seg, const int32_t* row_src,
                         const float* row_w, int M_max, const int* M_dev, const float* wsplit, const float* sc0, const float* sh0,
                         const float* sc1, const float* sh1, const float* gamma2, float* stats, float* Ykeep, float* ext, int32_t* arg,
                         float* part_ext, int32_t* part_arg, int32_t* seg_part, void* stream) {
  return gaddpg_sa1_fused_fwd_impl(phase, cloud, cloud_stride_b, cloud_stride_c, skip, Cp, bc, Cb, ctr, npoint, seg_off, row_seg, row_src,
                                   row_w, M_max, M_dev, wsplit, sc0, sh0, sc1, sh1, gamma2, stats, Ykeep, ext, arg, part_ext, part_arg,
                                   seg_part, stream);
}
int gaddpg_sa1_pool_finalize(const float* ext, const int32_t* arg, const float* part_ext, const int32_t* part_arg, int32_t* seg_part,
                             const float* gamma, const float* scale, const float* shift, int S, float* out, int32_t* arg_out,
                             void* stream) {
  return gaddpg_sa1_pool_finalize_impl(ext, arg, part_ext, part_arg, seg_part, gamma, scale, shift, S, out, arg_out, stream);
}
int gaddpg_gather_rows(const float* feats, int C, const float* xyz, int n_src, const float* ctr, int npoint,
                       const int32_t* row_seg, const int32_t* row_src, int M_max, const int* M_dev, float* G, int ldg,
                       void* stream) {
  return gaddpg_gather_rows_impl(feats, C, xyz, n_src, ctr, npoint, row_seg, row_src, M_max, M_dev, G, ldg, stream);
}
int gaddpg_scatter_rows(const float* dG, int ldg, int C, int B, int n_src, int npoint, const int32_t* seg_off,
                        const int32_t* row_src, float* dfeats, void* stream) {
  return gaddpg_scatter_rows_impl(dG, ldg, C, B, n_src, npoint, seg_off, row_src, dfeats, stream);
}
int gaddpg_pool_fwd(const float* Y, int C, const float* scale, const float* shift, const int32_t* seg_off, int fixed_len, int S,
                    float* out, int32_t* arg, void* stream) {
  return gaddpg_pool_fwd_impl(Y, C, scale, shift, seg_off, fixed_len, S, out, arg, stream);
}
int gaddpg_pool_bwd(const float* dOut, int ld_dout, const float* out, const int32_t* arg, const float* Y, int C,
                    const int32_t* row_seg, int fixed_len, int M_max, const int* M_dev, const float* mean, const float* rstd,
                    float* D, float* stats, const gaddpg_bn_tail* tail, void* stream) {
  return gaddpg_pool_bwd_impl(dOut, ld_dout, out, arg, Y, C, row_seg, fixed_len, M_max, M_dev, mean, rstd, D, stats, tail, stream);
}
int gaddpg_pool_bwd_sparse(const float* dOut, int ldo, const float* out, const int32_t* arg, const float* Y, int C, int S,
                           const float* mean, const float* rstd, float* E, uint32_t* mask, int M_max, float* stats,
                           const gaddpg_bn_tail* tail, void* stream) {
  return gaddpg_pool_bwd_sparse_impl(dOut, ldo, out, arg, Y, C, S, mean, rstd, E, mask, M_max, stats, tail, stream);
}
int gaddpg_pool_keys_finalize(unsigned long long* keys, int S, int C, const float* gamma, const float* scale, const float* shift, float* out,
                              int32_t* arg, void* stream) {
  return gaddpg_pool_keys_finalize_impl(keys, S, C, gamma, scale, shift, out, arg, stream);
}
int gaddpg_feat_finish(const float* Y, int C, const float* scale, const float* shift, const float* time, float time_offset, int B,
                       float* feat, int ld, void* stream) {
  return gaddpg_feat_finish_impl(Y, C, scale, shift, time, time_offset, B, feat, ld, stream);
}
int gaddpg_heads_init(const float* act_scale, const float* act_bias, const float* cp_rotz) {
  return gaddpg_heads_init_impl(act_scale, act_bias, cp_rotz);
}
int gaddpg_policy_head_fwd(const float* raw, int ldr, int B, float* pi, void* stream) {
  return gaddpg_policy_head_fwd_impl(raw, ldr, B, pi, stream);
}
int gaddpg_td3_next_action(const float* raw_t, int ldr, const float* u, float noise_scale, int B, float* next_action,
                           void* stream) {
  return gaddpg_td3_next_action_impl(raw_t, ldr, u, noise_scale, B, next_action, stream);
}
int gaddpg_td3_target(const float* qa, int ldq, int oq2, const float* reward, const float* done, float gamma, int B, float* y,
                      void* stream) {
  return gaddpg_td3_target_impl(qa, ldq, oq2, reward, done, gamma, B, y, stream);
}
int gaddpg_critic_loss(const float* qa, int ldq, int oq2, int oaux, const float* y, const float* perturb_flag, const float* ret,
                       const float* goal, int use_aux, int B, float grad_scale, float* dqa, float* out, void* stream) {
  return gaddpg_critic_loss_impl(qa, ldq, oq2, oaux, y, perturb_flag, ret, goal, use_aux, B, grad_scale, dqa, out, stream);
}
int gaddpg_actor_loss(const float* praw, int ldr, const float* pi, const float* expert_action, const float* expert_flag,
                      const float* ret, const float* goal, int use_aux, float bc_weight, const float* dpi_ac, int B,
                      float grad_scale, float* dpraw, int n_head, float* out, void* stream) {
  return gaddpg_actor_loss_impl(praw, ldr, pi, expert_action, expert_flag, ret, goal, use_aux, bc_weight, dpi_ac, B, grad_scale,
                                dpraw, n_head, out, stream);
}
int gaddpg_actor_critic_loss(const float* qa, int ldq, int oq2, const float* ret, const float* expert_flag, float mix, int B,
                             float grad_scale, int n_head, float* dqa, float* out, void* stream) {
  return gaddpg_actor_critic_loss_impl(qa, ldq, oq2, ret, expert_flag, mix, B, grad_scale, n_head, dqa, out, stream);
}
int gaddpg_quat_head(const float* raw, int ldr, int B, float* out7, void* stream) {
  return gaddpg_quat_head_impl(raw, ldr, B, out7, stream);
}
int gaddpg_policy_sample(const float* raw, int ldr, int off_logstd, const float* eps, int B, float* action, float* logp,
                         void* stream) {
  return gaddpg_policy_sample_impl(raw, ldr, off_logstd, eps, B, action, logp, stream);
}
int gaddpg_adam_step(float* p, float* g, float* m, float* v, long long n, double lr, double beta1, double beta2, double eps,
                     double weight_decay, long long step, const float* dyn, double grad_scale, const float* clip,
                     int write_back_grad, float* target, double tau, void* stream) {
  return gaddpg_adam_step_impl(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, step, dyn, grad_scale, clip, write_back_grad,
                               target, tau, stream);
}
int gaddpg_optim_multi(const void* jobs_dev, int njobs, int total_chunks, void* stream) {
  return gaddpg_optim_multi_impl(jobs_dev, njobs, total_chunks, stream);
}
int gaddpg_wprep_batched(const long long* jobs_dev, int njobs, void* stream) {
  return gaddpg_wprep_batched_impl(jobs_dev, njobs, stream);
}
int gaddpg_dmask_stats(const float* dX, int ldx, const float* Yprev, int C, int M, const float* psc, const float* psh,
                       const float* pmean, const float* prstd, float* D, float* stats, const gaddpg_bn_tail* tail, void* stream) {
  return gaddpg_dmask_stats_impl(dX, ldx, Yprev, C, M, psc, psh, pmean, prstd, D, stats, tail, stream);
}
int gaddpg_polyak(float* target, const float* source, long long n, double tau, void* stream) {
  return gaddpg_polyak_impl(target, source, n, tau, stream);
}
int gaddpg_polyak_vec(float* target, const float* source, const float* tau_vec, long long n, void* stream) {
  return gaddpg_polyak_vec_impl(target, source, tau_vec, n, stream);
}
int gaddpg_absmax(const float* x, long long n, float* out, float* ws, void* stream) {
  return gaddpg_absmax_impl(x, n, out, ws, stream);
}
int gaddpg_clip_coef(const float* g, long long n, float max_norm, float* coef_out, float* norm_out, float* ws, void* stream) {
  return gaddpg_clip_coef_impl(g, n, max_norm, coef_out, norm_out, ws, stream);
}
int gaddpg_wprep(const float* W, int N, int K, int rot, float* Wp, int ldp, float* WT, int ldt, void* stream) {
  return gaddpg_wprep_impl(W, N, K, rot, Wp, ldp, WT, ldt, stream);
}
int gaddpg_f64_to_f32(const double* src, float* dst, long long n, void* stream) {
  return gaddpg_f64_to_f32_impl(src, dst, n, stream);
}
int gaddpg_replay_gather(const float* cloud_store, long long row_floats, const float* rec_store, int rec_width, int ts_col,
                         const int32_t* episode_map, long long capacity, const int32_t* idx, int B, float* state_out, float* next_out,
                         float* rec_out, int32_t* inc_out, const int32_t* soa_map, float* soa_out, void* stream) {
  return gaddpg_replay_gather_impl(cloud_store, row_floats, rec_store, rec_width, ts_col, episode_map, capacity, idx, B, state_out,
                                   next_out, rec_out, inc_out, soa_map, soa_out, stream);
}

}  // extern "C"
