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
