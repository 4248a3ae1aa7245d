// head_ops.cu — actor/critic output transforms, TD3 target and the losses with their gradients (sm_100a).
//
// Restates on the device, with hand-derived backward passes, what the reference computes with ATen ops and
// autograd (B <= a few thousand rows, so each kernel is one CTA with a fixed-order block reduction):
//   policy head      tanh(mean)*scale+bias, quaternion-normalised aux     networks.py:339-371
//   TD3 target       noise, clamp, y = r + (1-done)*gamma*min(q1,q2)      ddpg.py:61-88, utils.py:568-584
//   critic losses    smooth-L1 on perturb_flag<1 rows + goal aux loss     ddpg.py:119-130, loss.py:17-23
//   actor losses     BC control-point loss, goal aux loss, -mix*mean(minQ) agent.py:127-139, ddpg.py:170-177,
//                    loss.py:25-31, utils.py:814-958
// Empty masks give 0/0 = NaN exactly as torch.mean of an empty tensor does.
#include "common.cuh"
#include "impl.h"

namespace {

// control points (utils.py:819-824); ROTZ = right-multiplied by Rz(pi/2) in float64 then cast (utils.py:826-829)
__constant__ float c_cp[6][3] = {{0.f, 0.f, 0.f},       {0.f, 0.f, 0.f},        {0.053f, -0.f, 0.075f},
                                 {-0.053f, 0.f, 0.075f}, {0.053f, -0.f, 0.105f}, {-0.053f, 0.f, 0.105f}};
__constant__ float c_cp_rotz[6][3];

__constant__ float c_act_scale[6];
__constant__ float c_act_bias[6];

struct V3 {
  float x, y, z;
};
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float sgn(float v) { return (v > 0.f) - (v < 0.f); }

// qrot (utils.py:940-958): v + 2*(w*(u x v) + u x (u x v))
__device__ __forceinline__ V3 qrot(float w, V3 u, V3 v) {
  V3 uv = cross(u, v), uuv = cross(u, uv);
  return {v.x + 2.f * (w * uv.x + uuv.x), v.y + 2.f * (w * uv.y + uuv.y), v.z + 2.f * (w * uv.z + uuv.z)};
}

__device__ float block_sum(float v, float* red) {  // fixed-order: warp shuffles, then warp 0 over the warp sums
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (warp == 0) t = warp_sum(t);
  if (threadIdx.x == 0) red[32] = t;
  __syncthreads();
  return red[32];
}

// Goal aux loss for one row: pred = [normalize(raw[0:4]), raw[4:7]] vs goal (quat used as given).
// Returns sum_k sum_xyz |cp_pred - cp_gt| and, scaled by gscale, the gradient w.r.t. raw[0:7].
__device__ float goal_row(const float* raw, const float* goal, float gscale, float* draw) {
  float qn = sqrtf(raw[0] * raw[0] + raw[1] * raw[1] + raw[2] * raw[2] + raw[3] * raw[3]);
  float den = fmaxf(qn, 1e-12f);  // F.normalize eps
  float w = raw[0] / den;
  V3 u = {raw[1] / den, raw[2] / den, raw[3] / den};
  V3 t = {raw[4], raw[5], raw[6]};
  float gw = goal[0];
  V3 gu = {goal[1], goal[2], goal[3]}, gt = {goal[4], goal[5], goal[6]};
  float loss = 0.f, dw = 0.f;
  V3 du = {0.f, 0.f, 0.f}, dt = {0.f, 0.f, 0.f};
  for (int k = 0; k < 6; ++k) {
    V3 v = {c_cp_rotz[k][0], c_cp_rotz[k][1], c_cp_rotz[k][2]};
    V3 p = qrot(w, u, v), q = qrot(gw, gu, v);
    V3 d = {p.x + t.x - (q.x + gt.x), p.y + t.y - (q.y + gt.y), p.z + t.z - (q.z + gt.z)};
    loss += fabsf(d.x) + fabsf(d.y) + fabsf(d.z);
    V3 g = {sgn(d.x), sgn(d.y), sgn(d.z)};
    dt.x += g.x;
    dt.y += g.y;
    dt.z += g.z;
    V3 uv = cross(u, v);
    dw += 2.f * dot(g, uv);
    V3 vxg = cross(v, g);
    float uvd = dot(u, v), gu_ = dot(g, u), gv = dot(g, v);
    du.x += 2.f * w * vxg.x + 2.f * (g.x * uvd + v.x * gu_ - 2.f * u.x * gv);
    du.y += 2.f * w * vxg.y + 2.f * (g.y * uvd + v.y * gu_ - 2.f * u.y * gv);
    du.z += 2.f * w * vxg.z + 2.f * (g.z * uvd + v.z * gu_ - 2.f * u.z * gv);
  }
  if (draw) {
    // through x / max(|x|, eps): d raw = (dn - n (n . dn)) / |x|   (the clamp branch has zero measure)
    float dn[4] = {dw, du.x, du.y, du.z}, n[4] = {w, u.x, u.y, u.z};
    float nd = n[0] * dn[0] + n[1] * dn[1] + n[2] * dn[2] + n[3] * dn[3];
    for (int i = 0; i < 4; ++i) draw[i] = gscale * (qn > 1e-12f ? (dn[i] - n[i] * nd) / den : dn[i] / den);
    draw[4] = gscale * dt.x;
    draw[5] = gscale * dt.y;
    draw[6] = gscale * dt.z;
  }
  return loss;
}

// BC loss for one row (loss.py:25-31): cp @ R^T + t with R = Rz(a5) Ry(a4) Rx(a3)  (utils.py:890-937)
__device__ float bc_row(const float* a, const float* e, float gscale, float* da) {
  auto points = [](const float* act, V3* out, V3* d3, V3* d4, V3* d5) {
    float cx = cosf(act[3]), sx = sinf(act[3]), cy = cosf(act[4]), sy = sinf(act[4]), cz = cosf(act[5]), sz = sinf(act[5]);
    for (int k = 0; k < 6; ++k) {
      V3 p = {c_cp[k][0], c_cp[k][1], c_cp[k][2]};
      V3 r1 = {p.x, cx * p.y - sx * p.z, sx * p.y + cx * p.z};              // Rx p
      V3 r1d = {0.f, -sx * p.y - cx * p.z, cx * p.y - sx * p.z};            // dRx/dax p
      V3 r2 = {cy * r1.x + sy * r1.z, r1.y, -sy * r1.x + cy * r1.z};        // Ry Rx p
      V3 r2d = {-sy * r1.x + cy * r1.z, 0.f, -cy * r1.x - sy * r1.z};       // dRy/day (Rx p)
      V3 r2x = {cy * r1d.x + sy * r1d.z, r1d.y, -sy * r1d.x + cy * r1d.z};  // Ry dRx p
      out[k] = {cz * r2.x - sz * r2.y + act[0], sz * r2.x + cz * r2.y + act[1], r2.z + act[2]};
      if (d3) {
        d3[k] = {cz * r2x.x - sz * r2x.y, sz * r2x.x + cz * r2x.y, r2x.z};
        d4[k] = {cz * r2d.x - sz * r2d.y, sz * r2d.x + cz * r2d.y, r2d.z};
        d5[k] = {-sz * r2.x - cz * r2.y, cz * r2.x - sz * r2.y, 0.f};
      }
    }
  };
  V3 pa[6], pe[6], d3[6], d4[6], d5[6];
  points(a, pa, d3, d4, d5);
  points(e, pe, nullptr, nullptr, nullptr);
  float loss = 0.f, g[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < 6; ++k) {
    V3 d = {pa[k].x - pe[k].x, pa[k].y - pe[k].y, pa[k].z - pe[k].z};
    loss += fabsf(d.x) + fabsf(d.y) + fabsf(d.z);
    V3 s = {sgn(d.x), sgn(d.y), sgn(d.z)};
    g[0] += s.x;
    g[1] += s.y;
    g[2] += s.z;
    g[3] += dot(s, d3[k]);
    g[4] += dot(s, d4[k]);
    g[5] += dot(s, d5[k]);
  }
  if (da)
    for (int i = 0; i < 6; ++i) da[i] = gscale * g[i];
  return loss;
}

// ---- policy head: raw [B, ldr] = [mean(6) | extra(E) | log_std(6)] -> pi = tanh(mean)*scale+bias -------------
__global__ void policy_head_fwd_kernel(const float* __restrict__ raw, int ldr, int B, float* __restrict__ pi) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * 6) return;
  int b = e / 6, c = e % 6;
  pi[e] = tanhf(raw[(long long)b * ldr + c]) * c_act_scale[c] + c_act_bias[c];
}

// TD3 target action (ddpg.py:77-82, utils.py:568-584): next = tanh(mean_t)*s+b + noise,
// noise = (u*3 - 6)*noise_scale, [:,3:] *= 5, [:, :3] clamped to +-0.01
__global__ void td3_next_action_kernel(const float* __restrict__ raw_t, int ldr, const float* __restrict__ u, float noise_scale,
                                       int B, float* __restrict__ next_action) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= B * 6) return;
  int b = e / 6, c = e % 6;
  float mean = tanhf(raw_t[(long long)b * ldr + c]) * c_act_scale[c] + c_act_bias[c];
  float nd = (u[e] * 3.f - 6.f) * noise_scale;
  if (c >= 3)
    nd *= 5.f;
  else
    nd = fminf(fmaxf(nd, -0.01f), 0.01f);
  next_action[e] = mean + nd;
}

// y = r + (1-done)*gamma*min(q1t,q2t)   (ddpg.py:86-87)
__global__ void td3_target_kernel(const float* __restrict__ qa, int ldq, int oq2, const float* __restrict__ reward,
                                  const float* __restrict__ done, float gamma, int B, float* __restrict__ y) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  y[b] = reward[b] + (1.f - done[b]) * gamma * fminf(qa[(long long)b * ldq], qa[(long long)b * ldq + oq2]);
}

// critic losses + gradients.  out[0]=critic_loss, out[1]=critic_grasp_aux_loss, out[2]=reward_mask_num.
// qa [B, ldq] = [q1 | q2 | aux_raw(7)] ; dqa same layout (d aux w.r.t. the RAW head output, through normalize)
__global__ void __launch_bounds__(256) critic_loss_kernel(const float* __restrict__ qa, int ldq, int oq2, int oaux, const float* __restrict__ y,
                                                          const float* __restrict__ perturb_flag,
                                                          const float* __restrict__ ret, const float* __restrict__ goal,
                                                          int use_aux, int B, float grad_scale, float* __restrict__ dqa,
                                                          float* __restrict__ out) {
  __shared__ float red[33];
  float nsel = 0.f, ngoal = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    nsel += (perturb_flag[b] < 1.f) ? 1.f : 0.f;
    ngoal += (ret[b] > 0.f) ? 1.f : 0.f;
  }
  nsel = block_sum(nsel, red);
  ngoal = block_sum(ngoal, red);
  float l1 = 0.f, l2 = 0.f, la = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float* r = qa + (long long)b * ldq;
    float* d = dqa + (long long)b * ldq;
    for (int c = 0; c < ldq; ++c) d[c] = 0.f;
    if (perturb_flag[b] < 1.f) {
      for (int c = 0; c < 2; ++c) {
        const int col = c ? oq2 : 0;
        float df = r[col] - y[b], ad = fabsf(df);
        float l = ad < 1.f ? 0.5f * df * df : ad - 0.5f;  // smooth_l1, beta = 1
        if (c == 0) l1 += l; else l2 += l;
        d[col] = grad_scale * (ad < 1.f ? df : sgn(df)) / nsel;
      }
    }
    if (use_aux && ret[b] > 0.f) la += goal_row(r + oaux, goal + (long long)b * 7, grad_scale / (6.f * ngoal), d + oaux);
  }
  l1 = block_sum(l1, red);
  l2 = block_sum(l2, red);
  la = block_sum(la, red);
  if (threadIdx.x == 0) {
    out[0] = l1 / nsel + l2 / nsel;
    out[1] = use_aux ? la / (6.f * ngoal) : 0.f;
    out[2] = ngoal;
  }
}

// actor losses + gradients w.r.t. the RAW policy head output praw [B, ldr] = [mean(6) | extra(E)...].
// out[0]=bc_loss (already * bc_weight), out[1]=policy_grasp_aux_loss, out[2]=#(return>0).
// dpi_ac [B,6] (may be NULL): gradient of the actor-critic term w.r.t. pi, produced by the value-encoder backward.
__global__ void __launch_bounds__(256) actor_loss_kernel(const float* __restrict__ praw, int ldr, const float* __restrict__ pi,
                                                         const float* __restrict__ expert_action,
                                                         const float* __restrict__ expert_flag,
                                                         const float* __restrict__ ret, const float* __restrict__ goal,
                                                         int use_aux, float bc_weight, const float* __restrict__ dpi_ac, int B,
                                                         float grad_scale, float* __restrict__ dpraw, int n_head,
                                                         float* __restrict__ out) {
  __shared__ float red[33];
  float nexp = 0.f, ngoal = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    nexp += (expert_flag[b] >= 1.f) ? 1.f : 0.f;
    ngoal += (ret[b] > 0.f) ? 1.f : 0.f;
  }
  nexp = block_sum(nexp, red);
  ngoal = block_sum(ngoal, red);
  float lbc = 0.f, la = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float* r = praw + (long long)b * ldr;
    float* d = dpraw + (long long)b * ldr;
    for (int c = 0; c < n_head; ++c) d[c] = 0.f;
    float dpi[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (expert_flag[b] >= 1.f) lbc += bc_row(pi + (long long)b * 6, expert_action + (long long)b * 6, grad_scale * bc_weight / (6.f * nexp), dpi);
    if (dpi_ac)
      for (int c = 0; c < 6; ++c) dpi[c] += dpi_ac[(long long)b * 6 + c];
    for (int c = 0; c < 6; ++c) {
      float th = tanhf(r[c]);
      d[c] = dpi[c] * c_act_scale[c] * (1.f - th * th);
    }
    if (use_aux && ret[b] > 0.f) la += goal_row(r + 6, goal + (long long)b * 7, grad_scale / (6.f * ngoal), d + 6);
  }
  lbc = block_sum(lbc, red);
  la = block_sum(la, red);
  if (threadIdx.x == 0) {
    out[0] = bc_weight * lbc / (6.f * nexp);
    out[1] = use_aux ? la / (6.f * ngoal) : 0.f;
    out[2] = ngoal;
  }
}

// actor-critic term (ddpg.py:175-177): loss = -mix * mean(min(q1,q2)[sel]), sel = ~(return>0 & expert>=1).
// qa [B, ldq] = [q1 | q2 | ...]; writes dqa (zeros outside q1,q2) and out[0] = loss.
__global__ void __launch_bounds__(256) actor_critic_loss_kernel(const float* __restrict__ qa, int ldq, int oq2, const float* __restrict__ ret,
                                                                const float* __restrict__ expert_flag, float mix, int B,
                                                                float grad_scale, int n_head, float* __restrict__ dqa,
                                                                float* __restrict__ out) {
  __shared__ float red[33];
  float nsel = 0.f, s = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    bool sel = !((ret[b] > 0.f) && (expert_flag[b] >= 1.f));
    nsel += sel ? 1.f : 0.f;
  }
  nsel = block_sum(nsel, red);
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const float* r = qa + (long long)b * ldq;
    float* d = dqa + (long long)b * ldq;
    for (int c = 0; c < n_head; ++c) d[c] = 0.f;
    bool sel = !((ret[b] > 0.f) && (expert_flag[b] >= 1.f));
    if (sel) {
      float q1 = r[0], q2 = r[oq2];
      s += fminf(q1, q2);
      float gq = -mix * grad_scale / nsel;
      // torch.min(a,b) backward: ties split the gradient evenly
      d[0] = q1 < q2 ? gq : (q1 == q2 ? 0.5f * gq : 0.f);
      d[oq2] = q2 < q1 ? gq : (q1 == q2 ? 0.5f * gq : 0.f);
    }
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = -mix * (s / nsel);
}

// aux head post-processing for select_action / inspection: out7 = [normalize(raw[0:4]), raw[4:7]]
__global__ void quat_head_kernel(const float* __restrict__ raw, int ldr, int B, float* __restrict__ out7) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* r = raw + (long long)b * ldr;
  float den = fmaxf(sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2] + r[3] * r[3]), 1e-12f);
  for (int i = 0; i < 4; ++i) out7[b * 7 + i] = r[i] / den;
  for (int i = 4; i < 7; ++i) out7[b * 7 + i] = r[i];
}

// Gaussian policy sample + log-prob (networks.py:353-371), used by select_action:
// raw = [mean(6) | extra(E) | log_std(6)], eps ~ N(0,1) given; outputs action[6], logp
__global__ void policy_sample_kernel(const float* __restrict__ raw, int ldr, int off_logstd, const float* __restrict__ eps, int B,
                                     float* __restrict__ action, float* __restrict__ logp) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const float* r = raw + (long long)b * ldr;
  float lp = 0.f;
  for (int c = 0; c < 6; ++c) {
    float mean = r[c];
    float ls = fminf(fmaxf(r[off_logstd + c], -10.f), 2.f);
    float sd = expf(ls);
    float x = mean + sd * eps[b * 6 + c];
    float yv = tanhf(x);
    action[b * 6 + c] = yv * c_act_scale[c] + c_act_bias[c];
    float l = -((x - mean) * (x - mean)) / (2.f * sd * sd) - ls - 0.9189385332046727f;
    l -= logf(c_act_scale[c] * (1.f - yv * yv) + 1e-6f);
    lp += l;
  }
  logp[b] = lp;
}

// __constant__ symbols exist once per device: remember which devices were initialised (a second agent on
// another GPU of the same process must not run with zeroed action scale / control points)
bool g_consts_ready[64] = {};

int cur_device() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < 64) ? d : 0;
}

}  // namespace

int gaddpg_heads_init_impl(const float* act_scale, const float* act_bias, const float* cp_rotz) {
  GADDPG_CHECK_ARG(act_scale && act_bias && cp_rotz, "heads_init: null pointer");
  GADDPG_CUDA(cudaMemcpyToSymbol(c_act_scale, act_scale, 6 * sizeof(float)));
  GADDPG_CUDA(cudaMemcpyToSymbol(c_act_bias, act_bias, 6 * sizeof(float)));
  GADDPG_CUDA(cudaMemcpyToSymbol(c_cp_rotz, cp_rotz, 18 * sizeof(float)));
  g_consts_ready[cur_device()] = true;
  return GADDPG_OK;
}

#define NEED_CONSTS(name) GADDPG_CHECK_ARG(g_consts_ready[cur_device()], name ": call gaddpg_heads_init on this device first")

int gaddpg_policy_head_fwd_impl(const float* raw, int ldr, int B, float* pi, void* stream) {
  NEED_CONSTS("policy_head_fwd");
  GADDPG_CHECK_ARG(raw && pi && ldr >= 6, "policy_head_fwd: bad argument");
  if (B == 0) return GADDPG_OK;
  policy_head_fwd_kernel<<<ceil_div(B * 6, 128), 128, 0, (cudaStream_t)stream>>>(raw, ldr, B, pi);
  GADDPG_CHECK_LAUNCH("policy_head_fwd_kernel");
  return GADDPG_OK;
}
int gaddpg_td3_next_action_impl(const float* raw_t, int ldr, const float* u, float noise_scale, int B, float* next_action,
                                void* stream) {
  NEED_CONSTS("td3_next_action");
  GADDPG_CHECK_ARG(raw_t && u && next_action && ldr >= 6, "td3_next_action: bad argument");
  if (B == 0) return GADDPG_OK;
  td3_next_action_kernel<<<ceil_div(B * 6, 128), 128, 0, (cudaStream_t)stream>>>(raw_t, ldr, u, noise_scale, B, next_action);
  GADDPG_CHECK_LAUNCH("td3_next_action_kernel");
  return GADDPG_OK;
}
int gaddpg_td3_target_impl(const float* qa, int ldq, int oq2, const float* reward, const float* done, float gamma, int B,
                           float* y, void* stream) {
  GADDPG_CHECK_ARG(qa && reward && done && y && oq2 >= 1 && oq2 < ldq, "td3_target: bad argument");
  if (B == 0) return GADDPG_OK;
  td3_target_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(qa, ldq, oq2, reward, done, gamma, B, y);
  GADDPG_CHECK_LAUNCH("td3_target_kernel");
  return GADDPG_OK;
}
int gaddpg_critic_loss_impl(const float* qa, int ldq, int oq2, int oaux, const float* y, const float* perturb_flag, const float* ret,
                            const float* goal, int use_aux, int B, float grad_scale, float* dqa, float* out, void* stream) {
  NEED_CONSTS("critic_loss");
  GADDPG_CHECK_ARG(qa && y && perturb_flag && ret && goal && dqa && out && oq2 >= 1 && oq2 < ldq && (!use_aux || oaux + 7 <= ldq),
                   "critic_loss: bad argument");
  critic_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(qa, ldq, oq2, oaux, y, perturb_flag, ret, goal, use_aux, B, grad_scale, dqa, out);
  GADDPG_CHECK_LAUNCH("critic_loss_kernel");
  return GADDPG_OK;
}
int gaddpg_actor_loss_impl(const float* praw, int ldr, const float* pi, const float* expert_action, const float* expert_flag,
                           const float* ret, const float* goal, int use_aux, float bc_weight, const float* dpi_ac, int B,
                           float grad_scale, float* dpraw, int n_head, float* out, void* stream) {
  NEED_CONSTS("actor_loss");
  GADDPG_CHECK_ARG(praw && pi && expert_action && expert_flag && ret && goal && dpraw && out, "actor_loss: null pointer");
  GADDPG_CHECK_ARG(n_head >= (use_aux ? 13 : 6) && ldr >= n_head, "actor_loss: head too narrow");
  actor_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(praw, ldr, pi, expert_action, expert_flag, ret, goal, use_aux, bc_weight,
                                                         dpi_ac, B, grad_scale, dpraw, n_head, out);
  GADDPG_CHECK_LAUNCH("actor_loss_kernel");
  return GADDPG_OK;
}
int gaddpg_actor_critic_loss_impl(const float* qa, int ldq, int oq2, const float* ret, const float* expert_flag, float mix, int B,
                                  float grad_scale, int n_head, float* dqa, float* out, void* stream) {
  GADDPG_CHECK_ARG(qa && ret && expert_flag && dqa && out && oq2 >= 1 && n_head > oq2 && ldq >= n_head, "actor_critic_loss: bad argument");
  actor_critic_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(qa, ldq, oq2, ret, expert_flag, mix, B, grad_scale, n_head, dqa, out);
  GADDPG_CHECK_LAUNCH("actor_critic_loss_kernel");
  return GADDPG_OK;
}
int gaddpg_quat_head_impl(const float* raw, int ldr, int B, float* out7, void* stream) {
  GADDPG_CHECK_ARG(raw && out7 && ldr >= 7, "quat_head: bad argument");
  if (B == 0) return GADDPG_OK;
  quat_head_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(raw, ldr, B, out7);
  GADDPG_CHECK_LAUNCH("quat_head_kernel");
  return GADDPG_OK;
}
int gaddpg_policy_sample_impl(const float* raw, int ldr, int off_logstd, const float* eps, int B, float* action, float* logp,
                              void* stream) {
  NEED_CONSTS("policy_sample");
  GADDPG_CHECK_ARG(raw && eps && action && logp && off_logstd >= 6 && ldr >= off_logstd + 6, "policy_sample: bad argument");
  if (B == 0) return GADDPG_OK;
  policy_sample_kernel<<<ceil_div(B, 128), 128, 0, (cudaStream_t)stream>>>(raw, ldr, off_logstd, eps, B, action, logp);
  GADDPG_CHECK_LAUNCH("policy_sample_kernel");
  return GADDPG_OK;
}
