// capi.cu — extern "C" surface of libgaddpg_b200.so (declared in include/gaddpg_b200.h).
#include <stdarg.h>
#include <string.h>

#include "../../include/gaddpg_b200.h"
#include "common.cuh"
#include "impl.h"

#define GADDPG_ABI_VERSION 1

static thread_local char g_err[512] = "";

void gaddpg_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" {

int gaddpg_version(void) { return GADDPG_ABI_VERSION; }
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

}  // extern "C"
