// Replay-buffer minibatch gather fused with VecNormalize and the policy's /255 input scaling.
//
// Restates [SB2] ReplayBuffer.sample(batch_size, env=vec_normalize) ->
// VecNormalize.normalize_obs / normalize_reward (configured at sb_helper.py:118-119: clip_obs=10,
// clip_reward=10, eps=1e-8; float64 arithmetic like numpy, cast to fp32 by the feed_dict) and
// observation_input(scale=True) ((x-0)/255 for the Box(0,255) of robot.py:224-228), and the
// channel split of custom_obs_policy.py:28-32 (last plane pixel [0,0] = direct feature).
// HBM-bound: one coalesced pass over 2*B observations; no intermediate copies.
#include "common.cuh"

namespace b2g {
namespace {

__global__ void __launch_bounds__(256) gather_kernel(GatherArgs g) {
  const int b = blockIdx.y;
  const int which = blockIdx.z;               // 0 = obs, 1 = next_obs
  const bool cnn = g.H > 0;
  const int E = cnn ? g.H * g.W * g.Cfull : g.W;
  const long long slot = g.indices ? (long long)g.indices[b] : (long long)b;
  const float* __restrict__ src = (which ? g.next_obs : g.obs) + (size_t)slot * E;
  const int cimg = g.Cfull - 1;
  const double ret_istd = g.normc[0], clip_obs = g.normc[1], clip_rew = g.normc[2];
  const bool norm_obs = g.normc[3] != 0.0, norm_rew = g.normc[4] != 0.0;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
    float y = src[e];
    if (norm_obs) {
      double d = ((double)y - g.mean[e]) * g.var[e];       // var[] holds 1/sqrt(var+eps) (set_norm_stats)
      d = fmin(fmax(d, -clip_obs), clip_obs);
      y = (float)d;
    }
    y = y / g.scale;
    if (cnn) {
      const int c = e % g.Cfull, pix = e / g.Cfull;
      if (c < cimg) {
        (which ? g.x_next : g.x_obs)[((size_t)b * g.H * g.W + pix) * cimg + c] = y;
      } else if (pix == 0) {
        if (which) g.F_t[(size_t)b * g.FS + g.feat_col] = y;
        else { g.F_pi[(size_t)b * g.FS + g.feat_col] = y; g.F_v[(size_t)b * g.FS + g.feat_col] = y; }
      }
    } else {
      if (which) g.F_t[(size_t)b * g.FS + e] = y;
      else { g.F_pi[(size_t)b * g.FS + e] = y; g.F_v[(size_t)b * g.FS + e] = y; }
    }
  }
  if (which == 0 && blockIdx.x == 0 && g.act) {
    const int feat_dim = cnn ? g.feat_col + 1 : g.W;
    if (threadIdx.x < g.n_act) g.F_v[(size_t)b * g.FS + feat_dim + threadIdx.x] = g.act[slot * g.n_act + threadIdx.x];
    if (threadIdx.x == 32) {
      float r = g.rew[slot];
      if (norm_rew) {
        double d = (double)r * ret_istd;
        d = fmin(fmax(d, -clip_rew), clip_rew);
        r = (float)d;
      }
      g.rew_out[b] = r;
      g.done_out[b] = g.done[slot];
    }
  }
}
}  // namespace

void gather_launch(const GatherArgs& a, cudaStream_t s) {
  const int E = a.H > 0 ? a.H * a.W * a.Cfull : a.W;
  int gx = (E + 256 * 4 - 1) / (256 * 4);
  if (gx < 1) gx = 1;
  dim3 grid(gx, a.B, a.next_obs ? 2 : 1);
  gather_kernel<<<grid, 256, 0, s>>>(a);
}

}  // namespace b2g
