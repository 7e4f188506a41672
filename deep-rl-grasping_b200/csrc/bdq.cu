// libb200grasp: branching dueling Q-network (BDQ) learner -- SURVEY.md section 8 row a11.
//
// The reference's BDQ lives in the absent `bdq_sb` fork (/root/reference/.gitmodules:1-3); call sites
// train_stable_baselines.py:103-104, sb_helper.py:202-226, hyper-parameters config/gripper_grasp.yaml:104-118.
// Algorithm restated in oracle/bdq_ref.py (Tavakoli et al., AAAI-18); variable names and shapes are the ones in
// trained_models/BDQ_8pads/BDQ_simple_8pads.zip.  PARITY UNPINNED (source absent).
//
// All layers are small dense contractions (<= 512 wide) and run on the fp32 gather-GEMM engine
// (gg_simt.cu) as grouped launches over the three network evaluations (online(s), online(s'), target(s'));
// the dueling aggregation, double-Q target, TD loss and backward seeds are one fused per-sample kernel.
// Output-layer weights are held with their row stride padded to 4 floats (n_bins = 33 -> 36) so every
// operand row is 16-byte aligned; get/set repack to the zip layout.
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "../../include/b200grasp.h"
#include "common.cuh"

using namespace b2g;

extern thread_local std::string g_b2g_err;     // sac.cu
static int bfail(int code, const std::string& msg) { g_b2g_err = msg; return code; }
#define BCK(call)                                                                                       \
  do {                                                                                                  \
    cudaError_t e_ = (call);                                                                            \
    if (e_ != cudaSuccess) return bfail(B2G_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

namespace {

struct BTensor {
  std::string name;
  int rows, cols, stride;   // zip shape [rows, cols] (cols = 1 for biases, rows = 1), device row stride
  int64_t off;              // float offset inside P (online) -- the target copy sits at off + n_train
  bool is_weight;
};

constexpr int BMET_LOSS = 0, BMET_MEANQ = 1, BMET_GN = MET_GN_PI;   // optim_kernel accumulates the squared norm at MET_GN_PI

struct BdqTailArgs {
  int B, D, n, NBS;            // batch, branches, bins, padded bin stride
  float gamma;
  const float* V[3];           // [B,4] value outputs of the 3 evaluations (col 0)
  const float* A[3][8];        // [B,NBS] advantages per evaluation / branch
  const float* act; int act_stride;   // action indices (as floats) inside the obs rows
  const float* rew; const float* done; const float* weights;
  float* dA[8];                // [B,NBS] gradient wrt advantages
  float* dV;                   // [B,4]
  float* td;                   // [B,D]
  float* metrics;
};

__global__ void bdq_tail_kernel(BdqTailArgs t) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  float loss_b = 0.f, q_b = 0.f;
  if (b < t.B) {
    const float invBD = 1.0f / (float)(t.B * t.D);
    float y = 0.f;
    for (int d = 0; d < t.D; ++d) {       // double-Q: online net picks, target net evaluates
      const float* a1 = t.A[1][d] + (size_t)b * t.NBS;
      int best = 0;
      float bv = a1[0];
      for (int k = 1; k < t.n; ++k) if (a1[k] > bv) { bv = a1[k]; best = k; }     // argmax_n (V + A - mean) = argmax_n A
      const float* a2 = t.A[2][d] + (size_t)b * t.NBS;
      float mean2 = 0.f;
      for (int k = 0; k < t.n; ++k) mean2 += a2[k];
      mean2 /= (float)t.n;
      y += t.V[2][(size_t)b * 4] + a2[best] - mean2;
    }
    y = t.rew[b] + t.gamma * (1.f - t.done[b]) * (y / (float)t.D);
    const float w = t.weights ? t.weights[b] : 1.f;
    float dv = 0.f;
    for (int d = 0; d < t.D; ++d) {
      const float* a0 = t.A[0][d] + (size_t)b * t.NBS;
      float mean0 = 0.f;
      for (int k = 0; k < t.n; ++k) mean0 += a0[k];
      mean0 /= (float)t.n;
      const int ai = (int)(t.act[(size_t)b * t.act_stride + d] + 0.5f);
      const float q = t.V[0][(size_t)b * 4] + a0[ai] - mean0;
      const float td = q - y;
      t.td[b * t.D + d] = td;
      loss_b += w * td * td;
      q_b += q;
      const float dq = 2.f * w * td * invBD;
      dv += dq;
      float* da = t.dA[d] + (size_t)b * t.NBS;
      for (int k = 0; k < t.NBS; ++k) da[k] = k < t.n ? dq * ((k == ai ? 1.f : 0.f) - 1.f / (float)t.n) : 0.f;
    }
    float* dvp = t.dV + (size_t)b * 4;
    dvp[0] = dv; dvp[1] = dvp[2] = dvp[3] = 0.f;
    loss_b *= invBD;
    q_b *= invBD;
  }
  for (int o = 16; o > 0; o >>= 1) {
    loss_b += __shfl_xor_sync(0xffffffffu, loss_b, o);
    q_b += __shfl_xor_sync(0xffffffffu, q_b, o);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(t.metrics + BMET_LOSS, loss_b); atomicAdd(t.metrics + BMET_MEANQ, q_b); }
}

__global__ void bdq_argmax_kernel(const float* const* A, int n_rows, int D, int n, int NBS, int* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_rows * D) return;
  const int b = i / D, d = i % D;
  const float* a = A[d] + (size_t)b * NBS;
  int best = 0;
  float bv = a[0];
  for (int k = 1; k < n; ++k) if (a[k] > bv) { bv = a[k]; best = k; }
  out[i] = best;
}

// ------------------------------------------------------------------------------------------------ prioritised replay
// Proportional prioritisation ([SB2] common/buffers.py PrioritizedReplayBuffer over common/segment_tree.py; Schaul et al.
// 2016) with the sum / min segment trees resident in HBM: leaves C..2C-1 (C = capacity rounded up to a power of two),
// node i = f(2i, 2i+1).  Sums are kept in float64 like the Python floats of the reference.
struct PerArgs {
  double* tsum; double* tmin; long long C;
  float* max_prio;                  // running max of the raw priorities (new transitions enter with it)
  const long long* counters;        // [4] rng step, [5] replay size
  unsigned long long seed;
  int B; float alpha, eps; const float* beta;
  int* indices; float* weights; float* prio_out;
  const float* td; int D;
};

// one CTA, B threads: draws B slots proportionally to priority (find_prefixsum_idx descent) and their IS weights
__global__ void per_sample_kernel(PerArgs a) {
  const int b = threadIdx.x;
  if (b >= a.B) return;
  const unsigned long long step = (unsigned long long)a.counters[4];
  const long long size = a.counters[5];
  const uint4 r = philox4x32_10(make_uint4((unsigned)step, (unsigned)(step >> 32), (unsigned)(b >> 2), 2u), make_uint2((unsigned)a.seed, (unsigned)(a.seed >> 32)));
  const unsigned v = (b & 3) == 0 ? r.x : (b & 3) == 1 ? r.y : (b & 3) == 2 ? r.z : r.w;
  const double total = a.tsum[1];
  double mass = ((double)v + 0.5) * (1.0 / 4294967296.0) * total;
  long long node = 1;
  while (node < a.C) {
    const double left = a.tsum[2 * node];
    if (left > mass) node = 2 * node;
    else { mass -= left; node = 2 * node + 1; }
  }
  long long idx = node - a.C;
  if (idx >= size) idx = size - 1;                       // (rounding at the right edge of the occupied range)
  const double beta = (double)a.beta[0];
  const double p_min = a.tmin[1] / total;
  const double max_w = pow(p_min * (double)size, -beta);
  const double p = a.tsum[a.C + idx] / total;
  a.indices[b] = (int)idx;
  a.weights[b] = (float)(pow(p * (double)size, -beta) / max_w);
}

// one CTA: writes `n` leaves and repairs their ancestors level by level (siblings recomputed redundantly: same values)
__global__ void per_write_kernel(PerArgs a, const int* __restrict__ slots, long long first_slot, long long cap, int n, int from_td) {
  const int i = threadIdx.x;
  long long leaf = 0;
  if (i < n) {
    const long long slot = slots ? (long long)slots[i] : (first_slot + i) % cap;
    float raw;
    if (from_td) {
      float s = 0.f;
      for (int d = 0; d < a.D; ++d) s += fabsf(a.td[i * a.D + d]);
      raw = s + a.eps;
      atomicMax(reinterpret_cast<int*>(a.max_prio), __float_as_int(raw));      // positive floats order like their bit patterns
      if (a.prio_out) a.prio_out[i] = raw;
    } else raw = a.max_prio[0];
    const double pr = pow((double)raw, (double)a.alpha);
    leaf = a.C + slot;
    a.tsum[leaf] = pr; a.tmin[leaf] = pr;
  }
  __syncthreads();
  for (long long span = a.C; span > 1; span >>= 1) {
    if (i < n) {
      leaf >>= 1;
      a.tsum[leaf] = a.tsum[2 * leaf] + a.tsum[2 * leaf + 1];
      a.tmin[leaf] = fmin(a.tmin[2 * leaf], a.tmin[2 * leaf + 1]);
    }
    __syncthreads();
  }
}

__global__ void per_init_kernel(double* tsum, double* tmin, long long n2, float* max_prio) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) { tsum[i] = 0.0; tmin[i] = INFINITY; }
  if (blockIdx.x == 0 && threadIdx.x == 0) max_prio[0] = 1.0f;
}

// hard target copy every `freq` updates, decided on the device so that the step can live in a CUDA graph
__global__ void bdq_target_copy_kernel(float* __restrict__ P, long long n_train, const long long* __restrict__ counters, int freq) {
  if (freq <= 0 || counters[3] % freq != 0) return;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_train; i += (long long)gridDim.x * blockDim.x) P[n_train + i] = P[i];
}

std::vector<int> iota_t(int n, int stride = 1, int base = 0) {
  std::vector<int> v(n);
  for (int i = 0; i < n; ++i) v[i] = base + i * stride;
  return v;
}
}  // namespace

struct b2g_bdq {
  b2g_bdq_cfg cfg{};
  int B = 0, D = 0, n = 0, NBS = 0, T0 = 0, T1 = 0, HB = 0, XS = 0, E = 0;
  std::vector<BTensor> tensors;
  std::map<std::string, int> tindex;
  int64_t n_train = 0;
  float *P = nullptr, *Mo = nullptr, *Vo = nullptr, *G = nullptr, *metrics = nullptr;
  float eps_value = 1.0f;      // bdq/eps (exploration epsilon variable of the zip)
  cudaStream_t stream = nullptr;
  std::vector<void*> allocs;
  float *r_obs = nullptr, *r_next = nullptr, *r_act = nullptr, *r_rew = nullptr, *r_done = nullptr;
  int64_t r_size = 0, r_pos = 0;
  double *d_mean = nullptr, *d_istd = nullptr, *d_normc = nullptr;
  float *X = nullptr, *Xn = nullptr, *Xscratch = nullptr;
  float *h1[3]{}, *h2[3]{}, *hb[3][8]{}, *Aout[3][8]{}, *hv[3]{}, *Vout[3]{};
  float *dA[8]{}, *dV = nullptr, *dcat = nullptr, *dh2 = nullptr, *dh1 = nullptr, *td = nullptr;
  float *rew_n = nullptr, *done_n = nullptr, *weights = nullptr, *eps_dummy = nullptr;
  float *s_obs = nullptr, *s_next = nullptr, *s_act = nullptr, *s_rew = nullptr, *s_done = nullptr;
  int* indices = nullptr;
  int* act_idx_out = nullptr;
  const float** d_Aptr = nullptr;
  long long* counters = nullptr;
  double* step_consts = nullptr;
  float* d_lr = nullptr;
  float cur_lr = -1.f;
  std::vector<GemmGroup> fwd, bwd, act;
  long long n_updates = 0;
  float* h_met = nullptr;
  void* nccl_comm = nullptr;
  // prioritised replay
  bool per = false;
  double *t_sum = nullptr, *t_min = nullptr;
  long long per_C = 0;
  float *max_prio = nullptr, *d_beta = nullptr, *prio_out = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  float* p(const std::string& nm) { return P + tensors[tindex.at(nm)].off; }
  float* g(const std::string& nm) { return G + tensors[tindex.at(nm)].off; }
  float* pt(const std::string& nm) { return P + n_train + tensors[tindex.at(nm)].off; }
};

namespace {
template <class T>
int balloc(b2g_bdq* h, T** ptr, size_t count) {
  void* q = nullptr;
  BCK(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
  BCK(cudaMemsetAsync(q, 0, std::max<size_t>(count, 1) * sizeof(T), h->stream));
  h->allocs.push_back(q);
  *ptr = (T*)q;
  return 0;
}
int btab(b2g_bdq* h, const std::vector<int>& v, const int** out) {
  int* d = nullptr;
  if (int rc = balloc(h, &d, v.size())) return rc;
  BCK(cudaMemcpyAsync(d, v.data(), v.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
  BCK(cudaStreamSynchronize(h->stream));
  *out = d;
  return 0;
}
std::string fcname(int i) { return i == 0 ? "fully_connected" : "fully_connected_" + std::to_string(i); }

void add_t(b2g_bdq* h, const std::string& name, int rows, int cols, bool w, int stride, int64_t& off) {
  BTensor t{name, rows, cols, stride, off, w};
  off += ((int64_t)(w ? rows * stride : stride) + 31) / 32 * 32;
  h->tindex[name] = (int)h->tensors.size();
  h->tensors.push_back(t);
}

GemmDesc mkd(const float* A, const int* aM, const int* aR, const float* B, const int* bR, const int* bN, float* C, const int* cM,
             const int* cN, int M, int N, int R, int flags) {
  GemmDesc d{};
  d.A = A; d.B = B; d.C = C; d.aM = aM; d.aR = aR; d.bR = bR; d.bN = bN; d.cM = cM; d.cN = cN;
  d.M = M; d.N = N; d.R = R; d.flags = flags; d.splitR = 1; d.alpha = 1.f;
  return d;
}
int fin_group(b2g_bdq* h, GemmGroup& g) {
  int start = 0;
  for (auto& d : g.host) {
    d.tiles_m = (d.M + GG_SIMT_BM - 1) / GG_SIMT_BM;
    d.tiles_n = (d.N + GG_SIMT_BN - 1) / GG_SIMT_BN;
    d.tile_start = start;
    d.tile_count = d.tiles_m * d.tiles_n;
    start += d.tile_count;
  }
  g.total_tiles = start;
  if (int rc = balloc(h, &g.dev, g.host.size())) return rc;
  BCK(cudaMemcpyAsync(g.dev, g.host.data(), g.host.size() * sizeof(GemmDesc), cudaMemcpyHostToDevice, h->stream));
  BCK(cudaStreamSynchronize(h->stream));
  return 0;
}

int build(b2g_bdq* h) {
  const int B = h->B, D = h->D, NBS = h->NBS, T0 = h->T0, T1 = h->T1, HB = h->HB, XS = h->XS, obs = h->cfg.obs_dim;
  const int* iT0; const int* iT1; const int* iHB; const int* iNBS; const int* i4; const int* iobs;
  const int* rXS; const int* rT0; const int* rT1; const int* rHB; const int* rNBS; const int* r4; const int* rcat;
  const int* kT0; const int* kT1; const int* kHB; const int* kNBS; const int* k4; const int* icat;
#define BT(var, vec) if (int rc = btab(h, (vec), &var)) return rc;
  BT(iT0, iota_t(T0)) BT(iT1, iota_t(T1)) BT(iHB, iota_t(HB)) BT(iNBS, iota_t(NBS)) BT(i4, iota_t(4)) BT(iobs, iota_t(XS))
  BT(rXS, iota_t(B, XS)) BT(rT0, iota_t(B, T0)) BT(rT1, iota_t(B, T1)) BT(rHB, iota_t(B, HB)) BT(rNBS, iota_t(B, NBS)) BT(r4, iota_t(B, 4))
  BT(rcat, iota_t(B, (D + 1) * HB)) BT(icat, iota_t((D + 1) * HB))
  BT(kT0, iota_t(std::max(obs, T0) + 8, T0)) BT(kT1, iota_t(std::max(T0, T1) + 8, T1)) BT(kHB, iota_t(T1 + 8, HB)) BT(kNBS, iota_t(HB + 8, NBS)) BT(k4, iota_t(HB + 8, 4))
  const std::string sc[3] = {"bdq/model", "bdq/model", "bdq/target_q_func/model"};
  auto W = [&](int e, const std::string& rel) { return e == 2 ? h->pt("bdq/model" + rel) : h->p("bdq/model" + rel); };
  // ---------------- forward (3 evaluations), also the policy-inference groups (evaluation 0 only)
  auto fwd_group = [&](const char* name, int layer) {
    GemmGroup g, a;
    g.name = name; a.name = std::string("act_") + name;
    for (int e = 0; e < 3; ++e) {
      const float* x = e == 0 ? h->X : h->Xn;
      std::vector<GemmDesc> ds;
      if (layer == 0) {
        GemmDesc d = mkd(x, rXS, iobs, W(e, "/common_net/" + fcname(0) + "/weights"), kT0, iT0, h->h1[e], rT0, iT0, B, T0, obs, GG_A_RVEC | GG_EPI_BIAS_RELU);
        d.bias = W(e, "/common_net/" + fcname(0) + "/biases"); ds.push_back(d);
      } else if (layer == 1) {
        GemmDesc d = mkd(h->h1[e], rT0, iT0, W(e, "/common_net/" + fcname(1) + "/weights"), kT1, iT1, h->h2[e], rT1, iT1, B, T1, T0, GG_A_RVEC | GG_EPI_BIAS_RELU);
        d.bias = W(e, "/common_net/" + fcname(1) + "/biases"); ds.push_back(d);
      } else if (layer == 2) {
        for (int q = 0; q <= D; ++q) {
          const std::string rel = q < D ? "/action_value/" + fcname(2 * q) : "/state_value/" + fcname(0);
          GemmDesc d = mkd(h->h2[e], rT1, iT1, W(e, rel + "/weights"), kHB, iHB, q < D ? h->hb[e][q] : h->hv[e], rHB, iHB, B, HB, T1, GG_A_RVEC | GG_EPI_BIAS_RELU);
          d.bias = W(e, rel + "/biases"); ds.push_back(d);
        }
      } else {
        for (int q = 0; q <= D; ++q) {
          const std::string rel = q < D ? "/action_value/" + fcname(2 * q + 1) : "/state_value/" + fcname(1);
          const int N = q < D ? NBS : 4;
          GemmDesc d = mkd(q < D ? h->hb[e][q] : h->hv[e], rHB, iHB, W(e, rel + "/weights"), q < D ? kNBS : k4, q < D ? iNBS : i4,
                           q < D ? h->Aout[e][q] : h->Vout[e], q < D ? rNBS : r4, q < D ? iNBS : i4, B, N, HB, GG_A_RVEC | GG_EPI_BIAS);
          d.bias = W(e, rel + "/biases"); ds.push_back(d);
        }
      }
      for (auto& d : ds) { g.host.push_back(d); if (e == 0) a.host.push_back(d); }
    }
    h->fwd.push_back(g); h->act.push_back(a);
  };
  fwd_group("bdq_trunk1", 0); fwd_group("bdq_trunk2", 1); fwd_group("bdq_hidden", 2); fwd_group("bdq_out", 3);
  // ---------------- backward (online evaluation 0)
  {
    GemmGroup g; g.name = "bdq_out_bwd";
    for (int q = 0; q <= D; ++q) {
      const std::string rel = q < D ? "bdq/model/action_value/" + fcname(2 * q + 1) : "bdq/model/state_value/" + fcname(1);
      const int N = q < D ? NBS : 4;
      const float* dz = q < D ? h->dA[q] : h->dV;
      const float* hin = q < D ? h->hb[0][q] : h->hv[0];
      GemmDesc w = mkd(hin, iHB, rHB, dz, q < D ? rNBS : r4, q < D ? iNBS : i4, h->g(rel + "/weights"), q < D ? kNBS : k4, q < D ? iNBS : i4, HB, N, B, GG_COLSUM);
      w.colsum = h->g(rel + "/biases");
      g.host.push_back(w);
      // d(hidden) = (dz . W_out^T) masked by relu, written into the concatenated buffer [B, (D+1) HB] at column q HB
      const int* ccol; BT(ccol, iota_t(HB, 1, q * HB))
      GemmDesc dg = mkd(dz, q < D ? rNBS : r4, q < D ? iNBS : i4, h->p(rel + "/weights"), q < D ? iNBS : i4, q < D ? kNBS : k4, h->dcat, rcat, ccol, B, HB, N,
                        GG_A_RVEC | GG_B_RVEC | GG_EPI_MASK);
      dg.mask = hin; dg.kM = rHB; dg.kN = iHB;
      g.host.push_back(dg);
    }
    h->bwd.push_back(g);
  }
  {
    GemmGroup g; g.name = "bdq_hidden_bwd";
    std::vector<int> br((D + 1) * HB);
    for (int q = 0; q <= D; ++q) {
      const std::string rel = q < D ? "bdq/model/action_value/" + fcname(2 * q) : "bdq/model/state_value/" + fcname(0);
      const int* dzcol; BT(dzcol, iota_t(B, (D + 1) * HB, q * HB))
      GemmDesc w = mkd(h->h2[0], iT1, rT1, h->dcat, dzcol, iHB, h->g(rel + "/weights"), kHB, iHB, T1, HB, B, GG_COLSUM);
      w.colsum = h->g(rel + "/biases");
      g.host.push_back(w);
      for (int r = 0; r < HB; ++r) br[q * HB + r] = (int)h->tensors[h->tindex.at(rel + "/weights")].off + r;
    }
    const int* brt; BT(brt, br)
    GemmDesc dg = mkd(h->dcat, rcat, icat, h->P, brt, kHB, h->dh2, rT1, iT1, B, T1, (D + 1) * HB, GG_A_RVEC | GG_B_RVEC | GG_EPI_MASK | GG_EPI_SCALE);
    dg.mask = h->h2[0]; dg.kM = rT1; dg.kN = iT1;
    dg.alpha = h->cfg.trunk_grad_rescale ? 1.0f / (float)(D + 1) : 1.0f;
    g.host.push_back(dg);
    h->bwd.push_back(g);
  }
  {
    GemmGroup g; g.name = "bdq_trunk2_bwd";
    GemmDesc w = mkd(h->h1[0], iT0, rT0, h->dh2, rT1, iT1, h->g("bdq/model/common_net/" + fcname(1) + "/weights"), kT1, iT1, T0, T1, B, GG_COLSUM);
    w.colsum = h->g("bdq/model/common_net/" + fcname(1) + "/biases");
    g.host.push_back(w);
    GemmDesc dg = mkd(h->dh2, rT1, iT1, h->p("bdq/model/common_net/" + fcname(1) + "/weights"), iT1, kT1, h->dh1, rT0, iT0, B, T0, T1, GG_A_RVEC | GG_B_RVEC | GG_EPI_MASK);
    dg.mask = h->h1[0]; dg.kM = rT0; dg.kN = iT0;
    g.host.push_back(dg);
    h->bwd.push_back(g);
  }
  {
    GemmGroup g; g.name = "bdq_trunk1_wgrad";
    GemmDesc w = mkd(h->X, iobs, rXS, h->dh1, rT0, iT0, h->g("bdq/model/common_net/" + fcname(0) + "/weights"), kT0, iT0, obs, T0, B, GG_COLSUM);
    w.colsum = h->g("bdq/model/common_net/" + fcname(0) + "/biases");
    g.host.push_back(w);
    h->bwd.push_back(g);
  }
  for (auto& g : h->fwd) if (int rc = fin_group(h, g)) return rc;
  for (auto& g : h->bwd) if (int rc = fin_group(h, g)) return rc;
  for (auto& g : h->act) if (int rc = fin_group(h, g)) return rc;
  (void)sc;
  return 0;
}

GatherArgs bgather(b2g_bdq* h, bool from_replay, bool with_next) {
  GatherArgs g{};
  g.obs = from_replay ? h->r_obs : h->s_obs;
  g.next_obs = with_next ? (from_replay ? h->r_next : h->s_next) : nullptr;
  g.act = with_next ? (from_replay ? h->r_act : h->s_act) : nullptr;
  g.rew = from_replay ? h->r_rew : h->s_rew;
  g.done = from_replay ? h->r_done : h->s_done;
  g.indices = from_replay ? h->indices : nullptr;
  g.mean = h->d_mean; g.var = h->d_istd; g.normc = h->d_normc;
  g.B = h->B; g.H = 0; g.W = h->cfg.obs_dim; g.Cfull = 1; g.scale = 1.f;
  g.F_pi = h->Xscratch; g.F_v = h->X; g.F_t = h->Xn; g.FS = h->XS; g.feat_col = 0;
  g.rew_out = h->rew_n; g.done_out = h->done_n; g.n_act = h->D;
  return g;
}

int bdq_issue(b2g_bdq* h, bool sampled, bool apply, const float* weights) {
  cudaStream_t s = h->stream;
  PrepArgs pa{};
  pa.counters = h->counters; pa.step_consts = h->step_consts; pa.lr = h->d_lr; pa.metrics = h->metrics;
  pa.indices = h->indices; pa.eps = h->eps_dummy; pa.B = h->B; pa.A = 1; pa.replay_size = nullptr;
  pa.seed = h->cfg.seed + 0x9E3779B97F4A7C15ull * (unsigned long long)h->cfg.rank; pa.gen = sampled ? 1 : 0; pa.apply = apply ? 1 : 0;
  prep_launch(pa, s);
  PerArgs pr{};
  if (h->per) {
    pr.tsum = h->t_sum; pr.tmin = h->t_min; pr.C = h->per_C; pr.max_prio = h->max_prio; pr.counters = h->counters; pr.seed = pa.seed;
    pr.B = h->B; pr.alpha = h->cfg.per_alpha; pr.eps = h->cfg.per_eps; pr.beta = h->d_beta; pr.indices = h->indices; pr.weights = h->weights;
    pr.prio_out = h->prio_out; pr.td = h->td; pr.D = h->D;
    if (sampled) { per_sample_kernel<<<1, ((h->B + 31) / 32) * 32, 0, s>>>(pr); weights = h->weights; }   // overwrites the uniform draw
  }
  gather_launch(bgather(h, sampled, true), s);
  BCK(cudaMemsetAsync(h->G, 0, (size_t)(h->n_train + MET_COUNT) * sizeof(float), s));
  for (auto& g : h->fwd) gg_simt_launch(g.dev, (int)g.host.size(), g.total_tiles, s);
  BdqTailArgs t{};
  t.B = h->B; t.D = h->D; t.n = h->n; t.NBS = h->NBS; t.gamma = h->cfg.gamma;
  for (int e = 0; e < 3; ++e) { t.V[e] = h->Vout[e]; for (int d = 0; d < h->D; ++d) t.A[e][d] = h->Aout[e][d]; }
  t.act = h->X + h->cfg.obs_dim; t.act_stride = h->XS;
  t.rew = h->rew_n; t.done = h->done_n; t.weights = weights;
  for (int d = 0; d < h->D; ++d) t.dA[d] = h->dA[d];
  t.dV = h->dV; t.td = h->td; t.metrics = h->metrics;
  bdq_tail_kernel<<<(h->B + 127) / 128, 128, 0, s>>>(t);
  for (auto& g : h->bwd) gg_simt_launch(g.dev, (int)g.host.size(), g.total_tiles, s);
  if (h->per && sampled) per_write_kernel<<<1, ((h->B + 31) / 32) * 32, 0, s>>>(pr, h->indices, 0, h->cfg.buffer_capacity, h->B, 1);   // update_priorities(|td| + eps)
  if (h->cfg.nranks > 1) {      // gradients + loss scalars averaged over the ranks (each rank sampled its own replay shard)
    BCK(cudaMemcpyAsync(h->G + h->n_train, h->metrics, 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (int rc = nccl_allreduce_sum_f32(h->nccl_comm, h->G, (size_t)(h->n_train + MET_COUNT), s)) return rc;
    BCK(cudaMemcpyAsync(h->metrics, h->G + h->n_train, 2 * sizeof(float), cudaMemcpyDeviceToDevice, s));
  }
  OptimArgs oa{};
  oa.P = h->P; oa.Mo = h->Mo; oa.Vo = h->Vo; oa.G = h->G; oa.T = h->P + h->n_train;
  oa.n_pi = (int)h->n_train; oa.n_values = 0; oa.n_ent = 0; oa.n_target = 0;
  oa.step_consts = h->step_consts; oa.tau = 0.f; oa.grad_scale = 1.0f / (float)h->cfg.nranks; oa.metrics = h->metrics; oa.apply = apply ? 1 : 0;
  optim_launch(oa, s);
  BCK(cudaGetLastError());
  if (apply) bdq_target_copy_kernel<<<64, 256, 0, s>>>(h->P, h->n_train, h->counters, h->cfg.target_update_freq);   // counters[3] = n_updates (prep)
  BCK(cudaGetLastError());
  return 0;
}

int bset_lr(b2g_bdq* h, float lr) {
  if (lr != h->cur_lr) {
    BCK(cudaStreamSynchronize(h->stream));
    BCK(cudaMemcpy(h->d_lr, &lr, sizeof(float), cudaMemcpyHostToDevice));
    h->cur_lr = lr;
  }
  return 0;
}
int bfetch(b2g_bdq* h, b2g_bdq_metrics* out) {
  BCK(cudaMemcpyAsync(h->h_met, h->metrics, MET_COUNT * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  BCK(cudaStreamSynchronize(h->stream));
  if (out) {
    const float inv = 1.0f / (float)h->cfg.nranks;
    out->loss = h->h_met[BMET_LOSS] * inv; out->mean_q = h->h_met[BMET_MEANQ] * inv; out->grad_norm = sqrtf(h->h_met[BMET_GN]);
    out->n_updates = h->n_updates;
  }
  return 0;
}
}  // namespace

extern "C" {

int b2g_bdq_destroy(b2g_bdq* h) {
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
  nccl_comm_destroy(h->nccl_comm);
  for (void* q : h->allocs) cudaFree(q);
  if (h->h_met) cudaFreeHost(h->h_met);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return 0;
}

int b2g_bdq_create(const b2g_bdq_cfg* cfg, b2g_bdq** out) {
  if (!cfg || !out) return bfail(B2G_EINVAL, "cfg/out is NULL");
  *out = nullptr;
  if (cfg->n_branches < 1 || cfg->n_branches > 8 || cfg->n_bins < 2 || cfg->n_bins > 64) return bfail(B2G_EINVAL, "n_branches in [1,8], n_bins in [2,64]");
  if (cfg->trunk0 % 4 || cfg->trunk1 % 4 || cfg->branch_hidden % 4 || cfg->trunk0 < 4 || cfg->trunk1 < 4 || cfg->branch_hidden < 4)
    return bfail(B2G_EINVAL, "layer widths must be positive multiples of 4");
  if (cfg->obs_dim < 1 || cfg->batch < 1 || cfg->buffer_capacity < 1) return bfail(B2G_EINVAL, "obs_dim, batch, buffer_capacity must be positive");
  if (cfg->nranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->nranks) return bfail(B2G_EINVAL, "bad rank/nranks");
  if (cfg->nranks > 1 && !cfg->nccl_id) return bfail(B2G_EINVAL, "nranks > 1 needs nccl_id");
  if (cfg->prioritized_replay && cfg->batch > 1024) return bfail(B2G_EINVAL, "prioritised replay supports batch <= 1024");
  int ndev = 0;
  BCK(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) return bfail(B2G_ECUDA, "no such CUDA device");
  BCK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop{};
  BCK(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) return bfail(B2G_ECUDA, std::string("libb200grasp is built for sm_100a only; found ") + prop.name);
  b2g_bdq* h = new b2g_bdq();
  h->cfg = *cfg;
  h->cfg.nccl_id = nullptr; h->cfg.nccl_lib = nullptr;
  h->per = cfg->prioritized_replay != 0;
  h->B = cfg->batch; h->D = cfg->n_branches; h->n = cfg->n_bins; h->NBS = (cfg->n_bins + 3) / 4 * 4;
  h->T0 = cfg->trunk0; h->T1 = cfg->trunk1; h->HB = cfg->branch_hidden; h->E = cfg->obs_dim;
  h->XS = (cfg->obs_dim + cfg->n_branches + 7) / 8 * 8;
  auto bail = [&](int rc) { std::string keep = g_b2g_err; b2g_bdq_destroy(h); g_b2g_err = keep; return rc; };
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(bfail(B2G_ECUDA, "stream"));
  // parameter inventory: same names as the zips (oracle/bdq_ref.py param_specs)
  int64_t off = 0;
  for (int d = 0; d < h->D; ++d) {
    add_t(h, "bdq/model/action_value/" + fcname(2 * d) + "/biases", 1, h->HB, false, h->HB, off);
    add_t(h, "bdq/model/action_value/" + fcname(2 * d) + "/weights", h->T1, h->HB, true, h->HB, off);
    add_t(h, "bdq/model/action_value/" + fcname(2 * d + 1) + "/biases", 1, h->n, false, h->NBS, off);
    add_t(h, "bdq/model/action_value/" + fcname(2 * d + 1) + "/weights", h->HB, h->n, true, h->NBS, off);
  }
  add_t(h, "bdq/model/common_net/" + fcname(0) + "/biases", 1, h->T0, false, h->T0, off);
  add_t(h, "bdq/model/common_net/" + fcname(0) + "/weights", cfg->obs_dim, h->T0, true, h->T0, off);
  add_t(h, "bdq/model/common_net/" + fcname(1) + "/biases", 1, h->T1, false, h->T1, off);
  add_t(h, "bdq/model/common_net/" + fcname(1) + "/weights", h->T0, h->T1, true, h->T1, off);
  add_t(h, "bdq/model/state_value/" + fcname(0) + "/biases", 1, h->HB, false, h->HB, off);
  add_t(h, "bdq/model/state_value/" + fcname(0) + "/weights", h->T1, h->HB, true, h->HB, off);
  add_t(h, "bdq/model/state_value/" + fcname(1) + "/biases", 1, 1, false, 4, off);
  add_t(h, "bdq/model/state_value/" + fcname(1) + "/weights", h->HB, 1, true, 4, off);
  h->n_train = off;
  int rc = 0;
  const int B = h->B, D = h->D;
#define BA(ptr, count) if ((rc = balloc(h, &(ptr), (size_t)(count)))) return bail(rc)
  BA(h->P, 2 * h->n_train); BA(h->Mo, h->n_train); BA(h->Vo, h->n_train); BA(h->G, h->n_train + MET_COUNT); BA(h->metrics, MET_COUNT);
  BA(h->counters, 8); BA(h->step_consts, 4); BA(h->d_lr, 1);
  const int64_t cap = cfg->buffer_capacity;
  BA(h->r_obs, cap * h->E); BA(h->r_next, cap * h->E); BA(h->r_act, cap * D); BA(h->r_rew, cap); BA(h->r_done, cap);
  BA(h->d_mean, h->E); BA(h->d_istd, h->E); BA(h->d_normc, 8);
  BA(h->X, (size_t)B * h->XS); BA(h->Xn, (size_t)B * h->XS); BA(h->Xscratch, (size_t)B * h->XS);
  for (int e = 0; e < 3; ++e) {
    BA(h->h1[e], B * h->T0); BA(h->h2[e], B * h->T1); BA(h->hv[e], B * h->HB); BA(h->Vout[e], B * 4);
    for (int d = 0; d < D; ++d) { BA(h->hb[e][d], B * h->HB); BA(h->Aout[e][d], B * h->NBS); }
  }
  for (int d = 0; d < D; ++d) BA(h->dA[d], B * h->NBS);
  BA(h->dV, B * 4); BA(h->dcat, (size_t)B * (D + 1) * h->HB); BA(h->dh2, B * h->T1); BA(h->dh1, B * h->T0); BA(h->td, B * D);
  BA(h->rew_n, B); BA(h->done_n, B); BA(h->weights, B); BA(h->eps_dummy, B + 8); BA(h->indices, B + 4); BA(h->act_idx_out, B * D);
  BA(h->s_obs, (size_t)B * h->E); BA(h->s_next, (size_t)B * h->E); BA(h->s_act, B * D); BA(h->s_rew, B); BA(h->s_done, B);
  BA(h->d_Aptr, 8);
  BA(h->d_beta, 1); BA(h->max_prio, 1); BA(h->prio_out, B);
  if (h->per) {
    h->per_C = 1;
    while (h->per_C < cap) h->per_C <<= 1;
    BA(h->t_sum, 2 * h->per_C); BA(h->t_min, 2 * h->per_C);
    per_init_kernel<<<256, 256, 0, h->stream>>>(h->t_sum, h->t_min, 2 * h->per_C, h->max_prio);
    const float beta0 = 0.4f;
    if (cudaMemcpyAsync(h->d_beta, &beta0, sizeof(float), cudaMemcpyHostToDevice, h->stream) != cudaSuccess) return bail(bfail(B2G_ECUDA, "per init"));
  }
#undef BA
  if (cudaMallocHost((void**)&h->h_met, MET_COUNT * sizeof(float)) != cudaSuccess) return bail(bfail(B2G_ECUDA, "cudaMallocHost"));
  {
    std::vector<double> ones(h->E, 1.0);
    const double nc[8] = {1.0, 10.0, 10.0, 0.0, 0.0, 0, 0, 0};
    const float* ap[8] = {};
    for (int d = 0; d < D; ++d) ap[d] = h->Aout[0][d];
    if (cudaMemcpyAsync(h->d_istd, ones.data(), h->E * sizeof(double), cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
        cudaMemcpyAsync(h->d_normc, nc, sizeof(nc), cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
        cudaMemcpyAsync(h->d_Aptr, ap, sizeof(ap), cudaMemcpyHostToDevice, h->stream) != cudaSuccess ||
        cudaStreamSynchronize(h->stream) != cudaSuccess)
      return bail(bfail(B2G_ECUDA, "init copies"));
  }
  if ((rc = build(h))) return bail(rc);
  if (cfg->nranks > 1) {
    if ((rc = nccl_comm_init(&h->nccl_comm, cfg->nranks, cfg->nccl_id, cfg->rank, cfg->nccl_lib))) return bail(rc);
    if ((rc = nccl_allreduce_sum_f32(h->nccl_comm, h->G, (size_t)(h->n_train + MET_COUNT), h->stream))) return bail(rc);   // warm-up outside capture
  }
  if (cudaStreamSynchronize(h->stream) != cudaSuccess) return bail(bfail(B2G_ECUDA, "create sync"));
  *out = h;
  return 0;
}

int b2g_bdq_param_count(const b2g_bdq* h) { return h ? 1 + 2 * (int)h->tensors.size() : 0; }

// index 0 = bdq/eps; 1..T = online tensors; T+1..2T = target tensors (names as in the zips)
int b2g_bdq_param_info(const b2g_bdq* h, int idx, char* name, size_t name_cap, int64_t* rows, int64_t* cols, int32_t* ndim) {
  if (!h || idx < 0 || idx >= b2g_bdq_param_count(h) || !name) return bfail(B2G_EINVAL, "bad tensor index");
  std::string nm = "bdq/eps";
  int64_t r = 1, c = 1;
  int nd = 0;
  if (idx > 0) {
    const int T = (int)h->tensors.size();
    const BTensor& t = h->tensors[(idx - 1) % T];
    nm = (idx - 1) < T ? t.name : std::string("bdq/target_q_func/model") + t.name.substr(strlen("bdq/model"));
    r = t.rows; c = t.cols; nd = t.is_weight ? 2 : 1;
  }
  snprintf(name, name_cap, "%s", nm.c_str());
  if (rows) *rows = r;
  if (cols) *cols = c;
  if (ndim) *ndim = nd;
  return 0;
}

static int bdq_copy(b2g_bdq* h, const char* name, float* arena_online, float* host, size_t numel, bool to_host, bool allow_target) {
  if (!h || !name || !host) return bfail(B2G_EINVAL, "NULL argument");
  std::string nm(name);
  if (nm.size() > 2 && nm.compare(nm.size() - 2, 2, ":0") == 0) nm.resize(nm.size() - 2);
  BCK(cudaSetDevice(h->cfg.device));
  BCK(cudaStreamSynchronize(h->stream));
  if (nm == "bdq/eps") {
    if (numel != 1) return bfail(B2G_EINVAL, "bdq/eps is a scalar");
    if (to_host) host[0] = h->eps_value; else h->eps_value = host[0];
    return 0;
  }
  bool target = false;
  const std::string tp = "bdq/target_q_func/model";
  if (nm.compare(0, tp.size(), tp) == 0) { target = true; nm = "bdq/model" + nm.substr(tp.size()); }
  if (target && !allow_target) return bfail(B2G_EINVAL, "not a trainable variable");
  auto it = h->tindex.find(nm);
  if (it == h->tindex.end()) return bfail(B2G_EINVAL, std::string("unknown variable: ") + name);
  const BTensor& t = h->tensors[it->second];
  const size_t rows = t.is_weight ? t.rows : 1, cols = t.is_weight ? t.cols : (size_t)(t.rows * t.cols);
  const size_t ccount = t.is_weight ? t.cols : (size_t)t.cols;
  if (numel != rows * ccount) return bfail(B2G_EINVAL, std::string("size mismatch for ") + name);
  float* dev = arena_online + t.off + (target ? h->n_train : 0);
  (void)cols;
  // repack between the zip layout [rows, cols] and the device row stride
  if (to_host) BCK(cudaMemcpy2D(host, ccount * sizeof(float), dev, t.stride * sizeof(float), ccount * sizeof(float), rows, cudaMemcpyDeviceToHost));
  else BCK(cudaMemcpy2D(dev, t.stride * sizeof(float), host, ccount * sizeof(float), ccount * sizeof(float), rows, cudaMemcpyHostToDevice));
  return 0;
}
int b2g_bdq_get_param(b2g_bdq* h, const char* name, float* dst, size_t numel) { return bdq_copy(h, name, h ? h->P : nullptr, dst, numel, true, true); }
int b2g_bdq_set_param(b2g_bdq* h, const char* name, const float* src, size_t numel) {
  return bdq_copy(h, name, h ? h->P : nullptr, const_cast<float*>(src), numel, false, true);
}
int b2g_bdq_get_grad(b2g_bdq* h, const char* name, float* dst, size_t numel) { return bdq_copy(h, name, h ? h->G : nullptr, dst, numel, true, false); }

int b2g_bdq_replay_add(b2g_bdq* h, const float* obs, const float* act_idx, const float* rew, const float* next_obs, const float* done, int64_t n) {
  if (!h || !obs || !act_idx || !rew || !next_obs || !done || n < 0) return bfail(B2G_EINVAL, "NULL argument");
  BCK(cudaSetDevice(h->cfg.device));
  const int64_t cap = h->cfg.buffer_capacity;
  int64_t done_n = 0;
  while (done_n < n) {
    const int64_t chunk = std::min(n - done_n, cap - h->r_pos);
    const size_t E = h->E, D = h->D;
    BCK(cudaMemcpyAsync(h->r_obs + h->r_pos * E, obs + done_n * E, chunk * E * sizeof(float), cudaMemcpyDefault, h->stream));
    BCK(cudaMemcpyAsync(h->r_next + h->r_pos * E, next_obs + done_n * E, chunk * E * sizeof(float), cudaMemcpyDefault, h->stream));
    BCK(cudaMemcpyAsync(h->r_act + h->r_pos * D, act_idx + done_n * D, chunk * D * sizeof(float), cudaMemcpyDefault, h->stream));
    BCK(cudaMemcpyAsync(h->r_rew + h->r_pos, rew + done_n, chunk * sizeof(float), cudaMemcpyDefault, h->stream));
    BCK(cudaMemcpyAsync(h->r_done + h->r_pos, done + done_n, chunk * sizeof(float), cudaMemcpyDefault, h->stream));
    if (h->per) {        // new transitions enter with the running maximum priority ([SB2] PrioritizedReplayBuffer.add)
      PerArgs pr{};
      pr.tsum = h->t_sum; pr.tmin = h->t_min; pr.C = h->per_C; pr.max_prio = h->max_prio; pr.alpha = h->cfg.per_alpha; pr.eps = h->cfg.per_eps;
      for (int64_t o = 0; o < chunk; o += 1024) {
        const int nn = (int)std::min<int64_t>(1024, chunk - o);
        per_write_kernel<<<1, ((nn + 31) / 32) * 32, 0, h->stream>>>(pr, nullptr, h->r_pos + o, cap, nn, 0);
      }
    }
    h->r_pos = (h->r_pos + chunk) % cap;
    h->r_size = std::min(cap, h->r_size + chunk);
    done_n += chunk;
  }
  const long long sz = h->r_size;
  BCK(cudaMemcpyAsync(h->counters + 5, &sz, sizeof(long long), cudaMemcpyHostToDevice, h->stream));
  BCK(cudaStreamSynchronize(h->stream));
  return 0;
}
int64_t b2g_bdq_replay_size(const b2g_bdq* h) { return h ? h->r_size : 0; }

int b2g_bdq_set_norm_stats(b2g_bdq* h, const double* obs_mean, const double* obs_var, double ret_var, double clip_obs, double clip_rew, double eps,
                           int norm_obs, int norm_reward) {
  if (!h) return bfail(B2G_EINVAL, "NULL handle");
  if (norm_obs && (!obs_mean || !obs_var)) return bfail(B2G_EINVAL, "norm_obs needs obs_mean/obs_var");
  BCK(cudaSetDevice(h->cfg.device));
  BCK(cudaStreamSynchronize(h->stream));
  if (norm_obs) {
    std::vector<double> istd(h->E);
    for (int i = 0; i < h->E; ++i) istd[i] = 1.0 / sqrt(obs_var[i] + eps);
    BCK(cudaMemcpy(h->d_mean, obs_mean, h->E * sizeof(double), cudaMemcpyHostToDevice));
    BCK(cudaMemcpy(h->d_istd, istd.data(), h->E * sizeof(double), cudaMemcpyHostToDevice));
  }
  const double nc[8] = {1.0 / sqrt(ret_var + eps), clip_obs, clip_rew, (double)norm_obs, (double)norm_reward, 0, 0, 0};
  BCK(cudaMemcpy(h->d_normc, nc, sizeof(nc), cudaMemcpyHostToDevice));
  return 0;
}

int b2g_bdq_step(b2g_bdq* h, int n_steps, float lr, b2g_bdq_metrics* out) {
  if (!h || n_steps < 0) return bfail(B2G_EINVAL, "bad argument");
  if (h->r_size < 1) return bfail(B2G_ESTATE, "replay buffer is empty");
  BCK(cudaSetDevice(h->cfg.device));
  if (int rc = bset_lr(h, lr)) return rc;
  static int no_graph = -1;
  if (no_graph < 0) { const char* e = getenv("B2G_NO_GRAPH"); no_graph = (e && e[0] == '1') ? 1 : 0; }
  if (!no_graph && !h->graph_exec) {       // the whole step (~14 launches of tiny layers) replays as one graph
    cudaGraph_t graph = nullptr;
    BCK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    const int rc = bdq_issue(h, true, true, nullptr);
    const cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) return bfail(B2G_ECUDA, std::string("BDQ graph capture failed: ") + cudaGetErrorString(e));
    const cudaError_t e2 = cudaGraphInstantiate(&h->graph_exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e2 != cudaSuccess) return bfail(B2G_ECUDA, std::string("BDQ graph instantiate failed: ") + cudaGetErrorString(e2));
  }
  for (int i = 0; i < n_steps; ++i) {
    if (h->graph_exec) BCK(cudaGraphLaunch(h->graph_exec, h->stream));
    else if (int rc = bdq_issue(h, true, true, nullptr)) return rc;
    ++h->n_updates;
  }
  return bfetch(h, out);
}

int b2g_bdq_set_per_beta(b2g_bdq* h, float beta) {
  if (!h) return bfail(B2G_EINVAL, "NULL handle");
  BCK(cudaSetDevice(h->cfg.device));
  BCK(cudaStreamSynchronize(h->stream));
  BCK(cudaMemcpy(h->d_beta, &beta, sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

int b2g_bdq_get_last_per(b2g_bdq* h, int32_t* slots, float* weights, float* priorities) {
  if (!h) return bfail(B2G_EINVAL, "NULL handle");
  BCK(cudaSetDevice(h->cfg.device));
  BCK(cudaStreamSynchronize(h->stream));
  if (slots) BCK(cudaMemcpy(slots, h->indices, h->B * sizeof(int32_t), cudaMemcpyDeviceToHost));
  if (weights) BCK(cudaMemcpy(weights, h->weights, h->B * sizeof(float), cudaMemcpyDeviceToHost));
  if (priorities) BCK(cudaMemcpy(priorities, h->prio_out, h->B * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

int b2g_bdq_step_explicit(b2g_bdq* h, const float* obs, const float* act_idx, const float* rew, const float* next_obs, const float* done,
                          const float* weights, float lr, int apply_update, b2g_bdq_metrics* out, float* td_out) {
  if (!h || !obs || !act_idx || !rew || !next_obs || !done) return bfail(B2G_EINVAL, "NULL argument");
  BCK(cudaSetDevice(h->cfg.device));
  if (int rc = bset_lr(h, lr)) return rc;
  const size_t B = h->B, E = h->E, D = h->D;
  BCK(cudaMemcpyAsync(h->s_obs, obs, B * E * sizeof(float), cudaMemcpyDefault, h->stream));
  BCK(cudaMemcpyAsync(h->s_next, next_obs, B * E * sizeof(float), cudaMemcpyDefault, h->stream));
  BCK(cudaMemcpyAsync(h->s_act, act_idx, B * D * sizeof(float), cudaMemcpyDefault, h->stream));
  BCK(cudaMemcpyAsync(h->s_rew, rew, B * sizeof(float), cudaMemcpyDefault, h->stream));
  BCK(cudaMemcpyAsync(h->s_done, done, B * sizeof(float), cudaMemcpyDefault, h->stream));
  if (weights) BCK(cudaMemcpyAsync(h->weights, weights, B * sizeof(float), cudaMemcpyDefault, h->stream));
  if (int rc = bdq_issue(h, false, apply_update != 0, weights ? h->weights : nullptr)) return rc;
  if (apply_update) ++h->n_updates;
  if (td_out) BCK(cudaMemcpyAsync(td_out, h->td, B * D * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
  return bfetch(h, out);
}

// greedy branch actions argmax_n Q_d(s, n) of the online network (the epsilon-greedy mixing is the caller's)
int b2g_bdq_act(b2g_bdq* h, const float* obs, int n, int32_t* act_idx_out) {
  if (!h || !obs || !act_idx_out || n < 0) return bfail(B2G_EINVAL, "bad argument");
  BCK(cudaSetDevice(h->cfg.device));
  const size_t E = h->E, D = h->D;
  for (int done_n = 0; done_n < n; done_n += h->B) {
    const int chunk = std::min(h->B, n - done_n);
    BCK(cudaMemcpyAsync(h->s_obs, obs + (size_t)done_n * E, chunk * E * sizeof(float), cudaMemcpyDefault, h->stream));
    GatherArgs g = bgather(h, false, false);
    gather_launch(g, h->stream);
    for (auto& gr : h->act) gg_simt_launch(gr.dev, (int)gr.host.size(), gr.total_tiles, h->stream);
    bdq_argmax_kernel<<<(chunk * (int)D + 127) / 128, 128, 0, h->stream>>>(h->d_Aptr, chunk, (int)D, h->n, h->NBS, h->act_idx_out);
    BCK(cudaMemcpyAsync(act_idx_out + (size_t)done_n * D, h->act_idx_out, chunk * D * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    BCK(cudaStreamSynchronize(h->stream));
  }
  BCK(cudaGetLastError());
  return 0;
}

}  // extern "C"
