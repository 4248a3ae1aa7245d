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

#ifdef __cplusplus
}
#endif
#endif /* GADDPG_B200_H */
