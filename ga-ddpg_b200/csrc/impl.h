// impl.h — internal (non-ABI) entry points implemented across the .cu files.
#pragma once
#include <stdint.h>

#define GADDPG_STR_(x) #x
#define GADDPG_STR(x) GADDPG_STR_(x)

// index_ops.cu
int gaddpg_fps_ballquery_impl(const float* xyz, long long sb, int sk, int sc, int B, int N, int m, int32_t* fps_idx,
                              float* new_xyz, int do_bq, float radius, int nsample, int32_t* bq_idx, int32_t* bq_cnt,
                              void* stream);
int gaddpg_ball_query_impl(const float* xyz, long long sb, int sk, int sc, int B, int N, int m, const float* new_xyz,
                           float radius, int nsample, int32_t* idx, int32_t* cnt, void* stream);
int gaddpg_gather_points_impl(int B, int C, int N, int m, const float* pts, const int32_t* idx, float* out, void* stream);
int gaddpg_gather_points_grad_impl(int B, int C, int N, int m, const float* grad_out, const int32_t* idx,
                                   float* grad_pts, void* stream);
int gaddpg_group_points_impl(int B, int C, int N, int m, int s, const float* pts, const int32_t* idx, float* out,
                             void* stream);
int gaddpg_group_points_grad_impl(int B, int C, int N, int m, int s, const float* grad_out, const int32_t* idx,
                                  float* grad_pts, void* stream);
int gaddpg_row_table_impl(int S, int nsample, const int32_t* cnt, const int32_t* idx, int32_t* seg_off,
                          int32_t* row_seg, int32_t* row_src, float* row_w, void* stream);

// gemm_rows.cu
#include "gemm_rows.cuh"
#include <stddef.h>
int gaddpg_gemm_nt_impl(const NTGroup* g, int nprob, int amode, int emode, void* stream);
size_t gaddpg_gemm_tn_workspace_bytes_impl();
int gaddpg_gemm_tn_impl(const TNProblem* p, int pmode, int qmode, float* dW, int ldd, int Ntrue, int Ktrue, int rot,
                        float* dbias, int accumulate, float* ws, size_t ws_bytes, void* stream);
int gaddpg_bn_finalize_fwd_impl(const float* stats, int C, double count, const float* gamma, const float* beta, float eps,
                                float momentum, float* running_mean, float* running_var, long long* nbt, int training,
                                float* scale, float* shift, float* mean_out, float* rstd_out, void* stream);
int gaddpg_bn_finalize_bwd_impl(const float* stats, int C, double count, const float* gamma, const float* rstd, float* g,
                                float* m1, float* m2, float* dgamma, float* dbeta, int accumulate, void* stream);
int gaddpg_bn_running_update_impl(float* running, const float* staged, long long n, float momentum, long long* nbt, int n_layers,
                                  void* stream);
// sa_ops.cu
int gaddpg_sa1_l1_fwd_impl(const float* cloud, long long cloud_sb, int cloud_sc, int skip, int Cp, const float* bc, int Cb,
                           int B, const float* ctr, int npoint, const int32_t* seg_off, const int32_t* row_seg,
                           const int32_t* row_src, const float* row_w, int M_max, const int* M_dev, const float* W, int ldw,
                           float* bcbias_ws, float* Y, float* stats, const gaddpg_bn_tail* tail, void* stream);
int gaddpg_sa1_l1_bwd_impl(const float* cloud, long long cloud_sb, int cloud_sc, int skip, int Cp, const float* bc, int Cb,
                           int B, const float* ctr, int npoint, const int32_t* seg_off, const int32_t* row_seg,
                           const int32_t* row_src, const float* row_w, int M_max, const int* M_dev, const float* D,
                           const float* Y, const float* g, const float* m1, const float* m2, const float* mean,
                           const float* rstd, const float* W, int ldw, float* dW, int accumulate, float* dbc, float* ws,
                           size_t ws_bytes, void* stream);
int gaddpg_gather_rows_impl(const float* feats, int C, const float* xyz, int n_src, const float* ctr, int npoint,
                            const int32_t* row_seg, const int32_t* row_src, int M_max, const int* M_dev, float* G, int ldg,
                            void* stream);
int gaddpg_scatter_rows_impl(const float* dG, int ldg, int C, int B, int n_src, int npoint, const int32_t* seg_off,
                             const int32_t* row_src, float* dfeats, void* stream);
int gaddpg_pool_fwd_impl(const float* Y, int C, const float* scale, const float* shift, const int32_t* seg_off, int fixed_len,
                         int S, float* out, int32_t* arg, void* stream);
int gaddpg_pool_bwd_impl(const float* dOut, int ldo, const float* out, const int32_t* arg, const float* Y, int C,
                         const int32_t* row_seg, int fixed_len, int M_max, const int* M_dev, const float* mean,
                         const float* rstd, float* D, float* stats, const gaddpg_bn_tail* tail, void* stream);
int gaddpg_pool_bwd_sparse_impl(const float* dOut, int ldo, const float* out, const int32_t* arg, const float* Y, int C, int S,
                                const float* mean, const float* rstd, float* E, uint32_t* mask, int M_max, float* stats,
                                const gaddpg_bn_tail* tail, void* stream);
int gaddpg_pool_keys_finalize_impl(unsigned long long* keys, int S, int C, const float* gamma, const float* scale, const float* shift, float* out,
                                   int32_t* arg, void* stream);
int gaddpg_feat_finish_impl(const float* Y, int C, const float* scale, const float* shift, const float* time,
                            float time_offset, int B, float* feat, int ld, void* stream);
// sa1_fused.cu
long long gaddpg_sa1f_wsplit_floats_impl();
int gaddpg_sa1_fused_grid_impl(int M_max);
int gaddpg_sa1f_wprep_impl(const float* W0, int ld0, int K1, const float* W1, const float* W2, float* wsplit, void* stream);
int gaddpg_sa1_fused_fwd_impl(int phase, const float* cloud, long long cloud_sb, int cloud_sc, int skip, int Cp, const float* bc, int Cb,
                              const float* ctr, int npoint, const int32_t* seg_off, const int32_t* row_seg, const int32_t* row_src,
                              const float* row_w, int M_max, const int* M_dev, const float* wsplit, const float* sc0, const float* sh0,
                              const float* sc1, const float* sh1, const float* gamma2, float* stats, float* Ykeep, float* ext,
                              int32_t* arg, float* part_ext, int32_t* part_arg, int32_t* seg_part, void* stream);
int gaddpg_sa1_pool_finalize_impl(const float* ext, const int32_t* arg, const float* part_ext, const int32_t* part_arg, int32_t* seg_part,
                                  const float* gamma, const float* scale, const float* shift, int S, float* out, int32_t* arg_out,
                                  void* stream);
// head_ops.cu
int gaddpg_heads_init_impl(const float* act_scale, const float* act_bias, const float* cp_rotz);
int gaddpg_policy_head_fwd_impl(const float* raw, int ldr, int B, float* pi, void* stream);
int gaddpg_td3_next_action_impl(const float* raw_t, int ldr, const float* u, float noise_scale, int B, float* next_action,
                                void* stream);
int gaddpg_td3_target_impl(const float* qa, int ldq, int oq2, const float* reward, const float* done, float gamma, int B,
                           float* y, void* stream);
int gaddpg_critic_loss_impl(const float* qa, int ldq, int oq2, int oaux, const float* y, const float* perturb_flag, const float* ret,
                            const float* goal, int use_aux, int B, float grad_scale, float* dqa, float* out, void* stream);
int gaddpg_actor_loss_impl(const float* praw, int ldr, const float* pi, const float* expert_action, const float* expert_flag,
                           const float* ret, const float* goal, int use_aux, float bc_weight, const float* dpi_ac, int B,
                           float grad_scale, float* dpraw, int n_head, float* out, void* stream);
int gaddpg_actor_critic_loss_impl(const float* qa, int ldq, int oq2, const float* ret, const float* expert_flag, float mix, int B,
                                  float grad_scale, int n_head, float* dqa, float* out, void* stream);
int gaddpg_quat_head_impl(const float* raw, int ldr, int B, float* out7, void* stream);
int gaddpg_policy_sample_impl(const float* raw, int ldr, int off_logstd, const float* eps, int B, float* action, float* logp,
                              void* stream);
// optim.cu
int gaddpg_adam_step_impl(float* p, float* g, float* m, float* v, long long n, double lr, double beta1, double beta2, double eps,
                          double weight_decay, long long step, const float* dyn, double grad_scale, const float* clip,
                          int write_back_grad, float* target, double tau, void* stream);
int gaddpg_optim_multi_impl(const void* jobs_dev, int njobs, int total_chunks, void* stream);
int gaddpg_wprep_batched_impl(const long long* jobs_dev, int njobs, void* stream);
int gaddpg_dmask_stats_impl(const float* dX, int ldx, const float* Yprev, int C, int M, const float* psc, const float* psh,
                            const float* pmean, const float* prstd, float* D, float* stats, const gaddpg_bn_tail* tail, void* stream);
int gaddpg_polyak_impl(float* target, const float* source, long long n, double tau, void* stream);
int gaddpg_polyak_vec_impl(float* target, const float* source, const float* tau_vec, long long n, void* stream);
int gaddpg_absmax_impl(const float* x, long long n, float* out, float* ws, void* stream);
int gaddpg_clip_coef_impl(const float* g, long long n, float max_norm, float* coef_out, float* norm_out, float* ws, void* stream);
int gaddpg_wprep_impl(const float* W, int N, int K, int rot, float* Wp, int ldp, float* WT, int ldt, void* stream);
int gaddpg_f64_to_f32_impl(const double* src, float* dst, long long n, void* stream);
int gaddpg_replay_gather_impl(const float* cloud_store, long long row_floats, const float* rec_store, int rec_width, int ts_col,
                              const int32_t* episode_map, long long capacity, const int32_t* idx, int B, float* state_out,
                              float* next_out, float* rec_out, int32_t* inc_out, const int32_t* soa_map, float* soa_out, void* stream);
// tc_gemm.cu
bool gaddpg_tc_gemm_supported(const NTProblem& p, int amode, int emode);
int gaddpg_tc_gemm_nt_impl(const NTProblem* p, int amode, int emode, void* stream);
// tc_gemm_kc.cu
bool gaddpg_tc_nt_kc_supported(const NTProblem& p, int amode, int emode);
int gaddpg_tc_nt_kc_impl(const NTProblem* p, int amode, int emode, void* stream);
bool gaddpg_tc_tn_supported(const TNProblem& p, int pmode, int qmode);
int gaddpg_tc_tn_impl(const TNProblem* p, int pmode, int qmode, float* ws, size_t ws_floats, float* ws_bias, int* splits_out,
                      void* stream);
// skinny_gemm.cu
bool gaddpg_skinny_supported(const NTGroup& g, int nprob, int amode, int emode);
int gaddpg_skinny_nt_impl(const NTGroup* g, int nprob, int amode, int emode, void* stream);
int gaddpg_skinny_enabled();
void gaddpg_set_tensor_core_impl(int enable);
int gaddpg_get_tensor_core_impl();
