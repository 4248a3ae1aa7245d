// replay.cu — device-resident replay buffer: minibatch gather.
//
// Replaces, for a buffer that lives in HBM, the producer of the update step's input dict:
//   BaseMemory.__getitem__ / post_process_batch (/root/reference/core/replay_memory.py:109-127,251-272): numpy fancy
//   indexing of every array at batch_idx and at increment_idx = min(episode_map[batch_idx], batch_idx + 1), the
//   remaining-time remap time = timestep[episode_map[idx]] + 1 - timestep[idx];
//   Agent.prepare_data (/root/reference/core/agent.py:211-240): float64 -> float32 conversion and the H2D copy.
// Pure byte movement: HBM-bound, 2 x B cloud rows read once and written once with 16-byte accesses.
#include "common.cuh"
#include "impl.h"

namespace {

// One (chunk, row) per CTA: rows 0..B-1 are the state clouds (store row idx[b]), rows B..2B-1 the next-state clouds
// (store row inc[b]).  Each thread keeps UNROLL independent 16-byte loads in flight before the first store.
template <int UNROLL>
__global__ void __launch_bounds__(256) replay_gather_cloud_kernel(const float4* __restrict__ store, long long row_vec,
                                                                  const int32_t* __restrict__ episode_map, long long capacity,
                                                                  const int32_t* __restrict__ idx, int B, float4* __restrict__ state_out,
                                                                  float4* __restrict__ next_out, int32_t* __restrict__ inc_out) {
  const int r = blockIdx.y;
  const int b = r < B ? r : r - B;
  long long i = idx[b];
  i = i < 0 ? 0 : (i >= capacity ? capacity - 1 : i);
  long long src = i;
  if (r >= B) {
    long long end = (long long)(uint32_t)episode_map[i];   // uint32 in the reference (replay_memory.py:381)
    src = end < i + 1 ? end : i + 1;                       // np.minimum(episode_map[idx], idx + 1)
    src = src >= capacity ? capacity - 1 : src;
    if (blockIdx.x == 0 && threadIdx.x == 0 && inc_out) inc_out[b] = (int32_t)src;
  }
  const float4* __restrict__ in = store + src * row_vec;
  float4* __restrict__ out = (r < B ? state_out : next_out) + (long long)b * row_vec;
  const long long base = (long long)blockIdx.x * (256 * UNROLL) + threadIdx.x;
  float4 v[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    long long k = base + u * 256;
    if (k < row_vec) v[u] = __ldg(in + k);
  }
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    long long k = base + u * 256;
    if (k < row_vec) out[k] = v[u];
  }
}

// scalar variant for rows whose float count is not a multiple of 4 (never the case for 4-channel clouds)
__global__ void __launch_bounds__(256) replay_gather_cloud_scalar_kernel(const float* __restrict__ store, long long row_floats,
                                                                         const int32_t* __restrict__ episode_map, long long capacity,
                                                                         const int32_t* __restrict__ idx, int B, float* __restrict__ state_out,
                                                                         float* __restrict__ next_out, int32_t* __restrict__ inc_out) {
  const int r = blockIdx.y;
  const int b = r < B ? r : r - B;
  long long i = idx[b];
  i = i < 0 ? 0 : (i >= capacity ? capacity - 1 : i);
  long long src = i;
  if (r >= B) {
    long long end = (long long)(uint32_t)episode_map[i];
    src = end < i + 1 ? end : i + 1;
    src = src >= capacity ? capacity - 1 : src;
    if (blockIdx.x == 0 && threadIdx.x == 0 && inc_out) inc_out[b] = (int32_t)src;
  }
  const float* __restrict__ in = store + src * row_floats;
  float* __restrict__ out = (r < B ? state_out : next_out) + (long long)b * row_floats;
  for (long long k = (long long)blockIdx.x * 256 + threadIdx.x; k < row_floats; k += (long long)gridDim.x * 256) out[k] = __ldg(in + k);
}

// One warp per sample: lane c moves record column c (rec_width <= 32).  rec_out[b] = [record(idx) | record(inc)], with
// the timestep column of the first half replaced by the remaining time (timestep[episode_end] + 1) - timestep[idx].
// soa_map (optional, 2*W ints: base[c], stride[c]; base < 0 = skip) additionally scatters the current record field-major
// into soa_out — the layout Agent.prepare_data's per-field device vectors have, so no per-field copies are needed.
__global__ void __launch_bounds__(256) replay_gather_records_kernel(const float* __restrict__ rec, int W, int ts_col,
                                                                    const int32_t* __restrict__ episode_map, long long capacity,
                                                                    const int32_t* __restrict__ idx, int B, float* __restrict__ rec_out,
                                                                    const int32_t* __restrict__ soa_map, float* __restrict__ soa_out) {
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  long long i = idx[b];
  i = i < 0 ? 0 : (i >= capacity ? capacity - 1 : i);
  long long end = (long long)(uint32_t)episode_map[i];
  end = end >= capacity ? capacity - 1 : end;
  long long inc = end < i + 1 ? end : i + 1;
  if (lane < W) {
    float cur = __ldg(rec + i * W + lane);
    float nxt = __ldg(rec + inc * W + lane);
    if (lane == ts_col) {
      float t_end = __ldg(rec + end * W + ts_col);
      cur = __fsub_rn(__fadd_rn(t_end, 1.0f), cur);  // float32(timestep[episode_map[idx]]) + 1 - time_batch
    }
    if (rec_out) {
      rec_out[(long long)b * 2 * W + lane] = cur;
      rec_out[(long long)b * 2 * W + W + lane] = nxt;
    }
    if (soa_map) {  // field-major copy of the current record: column c of sample b -> soa_out[base[c] + b * stride[c]]
      int base = soa_map[lane], stride = soa_map[W + lane];
      if (base >= 0) soa_out[base + (long long)b * stride] = cur;
    }
  }
}

}  // namespace

int gaddpg_replay_gather_impl(const float* cloud_store, long long row_floats, const float* rec_store, int rec_width, int ts_col,
                              const int32_t* episode_map, long long capacity, const int32_t* idx, int B, float* state_out,
                              float* next_out, float* rec_out, int32_t* inc_out, const int32_t* soa_map, float* soa_out,
                              void* stream) {
  GADDPG_CHECK_ARG(cloud_store && episode_map && idx && state_out && next_out, "replay_gather: null pointer");
  GADDPG_CHECK_ARG(B >= 0 && capacity >= 1 && row_floats >= 1, "replay_gather: bad size");
  GADDPG_CHECK_ARG(!rec_store || ((rec_out || soa_map) && rec_width >= 1 && rec_width <= 32 && ts_col >= 0 && ts_col < rec_width),
                   "replay_gather: record table needs rec_out or soa_map, 1 <= rec_width <= 32 and a timestep column");
  GADDPG_CHECK_ARG(!soa_map || (rec_store && soa_out), "replay_gather: soa_map needs rec_store and soa_out");
  if (B == 0) return GADDPG_OK;
  GADDPG_CHECK_ARG(2 * (long long)B <= 65535, "replay_gather: batch too large for one launch (2B <= 65535)");
  cudaStream_t s = (cudaStream_t)stream;
  bool vec = (row_floats % 4 == 0) && ((uintptr_t)cloud_store % 16 == 0) && ((uintptr_t)state_out % 16 == 0) &&
             ((uintptr_t)next_out % 16 == 0);
  if (vec) {
    constexpr int UNROLL = 8;
    long long row_vec = row_floats / 4;
    int chunks = (int)((row_vec + 256 * UNROLL - 1) / (256 * UNROLL));
    replay_gather_cloud_kernel<UNROLL><<<dim3(chunks, 2 * B), 256, 0, s>>>((const float4*)cloud_store, row_vec, episode_map, capacity, idx,
                                                                         B, (float4*)state_out, (float4*)next_out, inc_out);
    GADDPG_CHECK_LAUNCH("replay_gather_cloud_kernel");
  } else {
    int chunks = (int)((row_floats + 256 * 8 - 1) / (256 * 8));
    replay_gather_cloud_scalar_kernel<<<dim3(chunks, 2 * B), 256, 0, s>>>(cloud_store, row_floats, episode_map, capacity, idx, B, state_out,
                                                                        next_out, inc_out);
    GADDPG_CHECK_LAUNCH("replay_gather_cloud_scalar_kernel");
  }
  if (rec_store) {
    replay_gather_records_kernel<<<(B + 7) / 8, 256, 0, s>>>(rec_store, rec_width, ts_col, episode_map, capacity, idx, B, rec_out,
                                                             soa_map, soa_out);
    GADDPG_CHECK_LAUNCH("replay_gather_records_kernel");
  }
  return GADDPG_OK;
}
