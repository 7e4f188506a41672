// Internal declarations shared by sac.cu (handle, step orchestration, C ABI) and engine_v2.cu (TMA-fed tcgen05 engine).
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/b200grasp.h"
#include "cg.cuh"
#include "common.cuh"

extern thread_local std::string g_b2g_err;
int b2g_fail(int code, const std::string& msg);
#define B2G_CK(call)                                                                              \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      return b2g_fail(B2G_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + __FILE__ + ":" + \
                                     std::to_string(__LINE__));                                   \
  } while (0)

namespace b2g {
struct Tensor {
  std::string name;
  int ndim;
  int64_t shape[4];
  int64_t numel;
  int64_t off;    // float offset inside P
  int group;      // 0 pi, 1 values, 2 ent, 3 target
};


// ---- engine v2 (cg.cu): BF16 plane tensors, tensor maps and problem groups of the TMA-fed path
struct V2State {
  bool on = false;               // forward chain on the new engine
  bool bwd = false;              // backward chain on the new engine
  int KF = 576;                  // feature-row width of the F planes (513 features + actions, zero padded to 9 x 64)
  // activations: [which / net][plane]
  uint16_t* A1[2][3]{};          // im2col of the normalised image: [B*225][64*Ci]  (0 = obs, 1 = next_obs)
  uint16_t* H1[3][3]{};          // [B*225][32]
  uint16_t* H2[3][3]{};          // [B*36][64]
  uint16_t* H3[3][3]{};          // [B][1024]
  uint16_t* F[3][3]{};           // [B][KF]
  // gradient maps, 2 planes
  uint16_t* dz0pi[2]{};          // [B][64]
  uint16_t* dz0v[2]{};           // [B][192]
  uint16_t* dZ4[2][2]{};         // [net][plane] [B][512]
  uint16_t* dZ3[2][2]{};         // [B*16][64]
  uint16_t* dZ2[2][2]{};         // [B*36][64]
  uint16_t* dZ1[2]{};            // [plane] [B*225][2 nets][32]
  // weights as planes (refreshed every step by planes2_kernel)
  uint16_t* W1T[2][3]{};         // [0]: obs [64 = pi|vf][64*Ci], [1]: target [32][64*Ci]
  uint16_t* W2T[3][3]{};         // [net][plane] [64][512]
  uint16_t* W3T[3][3]{};         // [64][576]
  uint16_t* WfT[3][3]{};         // [512][1024]
  uint16_t* K0T[3][3]{};         // [0] pi [64][KF], [1] values vf|q1|q2 [192][KF], [2] target vf [64][KF]
  uint16_t* W2n[2][2]{};         // natural [512][64] (dgrad B operand), nets pi / values, 2 planes
  uint16_t* W3n[2][2]{};         // [576][64]
  uint16_t* Wfn[2][2]{};         // [1024][512]
  uint16_t* K0n[2][2]{};         // [0] pi [513 -> 576 rows][64], [1] values packed [576 rows][192]
  float* z0v = nullptr;          // fc0 pre-activations of vf|q1|q2: [B][192]
  void* plane_jobs = nullptr; int n_plane_jobs = 0, plane_ctas = 0;
  int* plane_cta_job = nullptr;  // job index of every CTA of the planes launch
  void* colsum_part[2]{}; int n_colsum_part[2]{}, colsum_ctas_part[2]{};   // [0] cnn_fc1 biases (early), [1] conv biases
  int sm_reserve = 0;            // SMs left to a collective that runs concurrently with the persistent GEMM grids
  std::vector<CUtensorMap> maps; // host copy
  std::vector<char> map_whole;   // per map: the box spans every plane
  std::vector<int> map_box_bytes;
  CUtensorMap* d_maps = nullptr;
  std::vector<CgGroup> fwd, bwd_groups;
  // fused launches: the layer groups above concatenated into one persistent launch each, chained by arrival counters
  std::vector<CgGroup> fwd_fused, bwd_fused;
  bool fuse = false;
  int split_fc1 = 3;             // K-splits of the cnn_fc1 forward tiles (1 = none)
  int split_fc1_dgrad = 1;       // K-splits of the cnn_fc1 dgrad tiles (B2G_SPLIT_FC1_DGRAD; opt-in)
  bool epi_colsum = true;        // conv / cnn_fc1 bias gradients come from the DGRAD epilogues (else: colsum2 launches over the planes)
  int* dep_ctr = nullptr; int n_dep_ctr = 0;
  std::vector<int*> tabs;
  int dbg = 0;
};

}  // namespace b2g

struct b2g_sac;
namespace b2g {
int v2_alloc(b2g_sac* h);    // plane tensors (before the v1 groups are built: the v1 backward reads planes 0 / 1 of the same buffers)
int v2_create(b2g_sac* h);   // tensor maps + problem groups
int v2_planes(b2g_sac* h, cudaStream_t s);
int v2_gather(b2g_sac* h, const GatherArgs& ga, cudaStream_t s);
int v2_compact_rows(b2g_sac* h, const float* src_full, float* dst, long long first_row, long long wrap, int n, cudaStream_t s);   // full [n][E] -> compact rows (first_row + i) % wrap
int v2_launch(b2g_sac* h, const CgGroup& g, cudaStream_t s);
int v2_colsum(b2g_sac* h, cudaStream_t s, int part);   // bias gradients from the gradient-map planes
}  // namespace b2g

using namespace b2g;   // (internal header: only library translation units include it)

struct b2g_sac {
  b2g_sac_cfg cfg{};
  bool cnn = false;
  int num_sms = 148;
  int B = 0, A = 0, H = 0, E = 0, Cimg = 0, feat_dim = 0, FS = 0;
  int Hi = 0, Wi = 0, H1 = 0, W1 = 0, H2 = 0, W2 = 0, H3 = 0, W3 = 0;
  std::vector<Tensor> tensors;
  std::map<std::string, int> tindex;
  int64_t n_pi = 0, n_values = 0, n_ent = 0, n_target = 0, n_train = 0, n_all = 0;
  float *P = nullptr, *Mo = nullptr, *Vo = nullptr, *G = nullptr;   // G has MET_COUNT extra floats (metrics ride the all-reduce)
  float* metrics = nullptr;
  cudaStream_t stream = nullptr;
  std::vector<void*> allocs;
  // replay
  float *r_obs = nullptr, *r_next = nullptr, *r_act = nullptr, *r_rew = nullptr, *r_done = nullptr;
  int64_t r_size = 0, r_pos = 0;
  // normalisation
  double *d_mean = nullptr, *d_istd = nullptr, *d_normc = nullptr;   // normc: ret_istd, clip_obs, clip_rew, norm_obs, norm_rew
  double ret_istd = 1.0, clip_obs = 10.0, clip_rew = 10.0;
  int norm_obs = 0, norm_rew = 0;
  // batch buffers
  float *x_obs = nullptr, *x_next = nullptr;
  float *h1[3]{}, *h2[3]{}, *h3[3]{}, *F[3]{};
  float *dZ4[2]{}, *dZ3p[2]{}, *dZ2p[2]{}, *dZ1[2]{};
  float *z0[5]{}, *a0[4]{}, *dz1[4]{}, *dz0_pi = nullptr, *dz0_v3 = nullptr;
  bool heads_wgrad_simt = true;  // head weight gradients on the CUDA cores (tail.cu) instead of a v1 tensor-engine launch
  // BF16 hi/lo planes ([..][0] = hi, [..][1] = lo) of the tensors that feed forward / dgrad contractions
  bool use_planes = false;
  uint16_t *xp[2][2]{}, *h1p[3][2]{}, *h2p[3][2]{}, *h3p[3][2]{};
  uint16_t *dZ4p[2][2]{}, *dZ3pp[2][2]{}, *dZ2pp[2][2]{}, *dZ1p[2][2]{};
  bool wgrad_planes = false;
  ColsumJob* d_colsum = nullptr;
  int n_colsum = 0, colsum_ctas = 0;
  uint16_t* wp[3][4][4]{};          // [net][cnn1,cnn2,cnn3,fc1][hi, lo, hiT, loT]
  PlaneJob* d_jobs = nullptr;
  int n_jobs = 0, job_tiles = 0;
  bool planes_dirty = true;
  long long* dbg_trace = nullptr;
  std::map<const int*, std::vector<int>> host_tabs;   // host copies of the offset tables (contract checks at build time)
  float *per_sample = nullptr, *pi_out = nullptr, *eps = nullptr, *rew_n = nullptr, *done_n = nullptr;
  int* indices = nullptr;
  float *s_obs = nullptr, *s_next = nullptr, *s_act = nullptr, *s_rew = nullptr, *s_done = nullptr;  // staged explicit batch
  // pipelined host-batch path: the big obs / next_obs copies ping-pong on a copy stream
  float *ps_obs[2]{}, *ps_next[2]{};
  // host-pipelined steps, compact transfer: the constant actuator plane never crosses PCIe (b2g_sac_step_host_pipelined)
  float *hc_obs[2]{}, *hc_next[2]{};   // pinned host staging, compact rows [B][Ec]
  bool staged_compact = false;         // the staged batch already is in compact rows
  cudaGraphExec_t pipe_graph[2]{};     // the step on staging slot j, captured once (b2g_sac_step_host_pipelined)
  bool pipe_graph_compact[2]{};
  int host_threads = 16;
  cudaStream_t cstream = nullptr;
  cudaEvent_t ev_h2d[2]{}, ev_consumed[2]{}, ev_met[2]{};
  float* pm_met[2]{};            // pinned: MET_COUNT floats + [log_alpha, grad log_alpha]
  long long* pm_cnt[2]{};        // pinned counters
  long long pipe_k = 0;
  bool pipe_pending = false;
  cudaEvent_t record_after_gather = nullptr;
  long long* counters = nullptr;
  double* step_consts = nullptr;
  float* d_lr = nullptr;
  float cur_lr = -1.f;
  // launches
  std::vector<GemmGroup> fwd_groups, bwd_groups, act_groups;
  cudaGraphExec_t graph_exec = nullptr;
  bool use_graph = true;
  void* nccl_comm = nullptr;
  void* nccl_comm2 = nullptr;          // second communicator: early all-reduce on the side stream
  // peer-memory data parallelism (b2g_sac_dp_export / _connect): every rank maps the others' parameter arena, gradient buffer and
  // exchange block through CUDA IPC; the optimiser kernel then IS the collective (optim.cu: dp_optim_kernel)
  // host-pipelined steps: copy / kernel schedule picked by timing both (b2g_sac_step_host_pipelined)
  bool pipe_serial = false; int pipe_tune = 0; double pipe_t0 = 0, pipe_period[2] = {0, 0};
  bool dp_p2p = false;
  int* dp_x = nullptr;                 // exchange block: int flags[2][8], float part[8][2]
  int* dp_sync = nullptr;              // local CTA counter + norm accumulators
  float* dp_recv = nullptr;            // receive arena: [src rank][my slice] gradient copies pushed by the other ranks
  float* dp_P[8]{}; float* dp_G[8]{}; int* dp_X[8]{};     // every rank's parameter arena, receive arena, exchange block (own = local)
  std::vector<void*> dp_opened;        // cudaIpcOpenMemHandle results
  int dp_skip[2][2]{};                 // float4 ranges of the gradient arena pushed by the cnn_fc1 wgrad epilogues
  cudaStream_t side = nullptr;
  cudaStream_t aux = nullptr;              // leaf work off the critical chain (zeroing, weight planes, leaf wgrads, bias sums)
  cudaEvent_t ev_aux[7]{};
  bool fork_leaves = false;
  std::map<std::tuple<const void*, const void*, const void*, int>, int> col_ids;
  bool tc_ranges = false;                  // contiguous cost-balanced tile ranges per CTA: measured SLOWER than round-robin
                                           // (split-R tiles of one output pile their atomics onto one CTA); B2G_TC_RANGES=1 enables
  bool early_opt = false;                  // early fc1/heads optimiser pass on the leaf branch: measured no gain (B2G_EARLY_OPT=1 enables)
  bool fuse_fwd = false;                   // B2G_FUSE_FWD=1: the CNN forward chain as one layer-synchronised launch
  unsigned* sync_ctr = nullptr;            // its completion counter (zeroed every step)
  bool a_rowlanes = true;                  // conv1 fwd gather with row-major lane order (B2G_ROWLANES=0 disables)
  bool fc0_split = true;                   // split-R heads_fc0 (needs z0 zeroed every step)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool overlap_ar = false;
  int ar_sms = 8;
  ColsumJob* d_colsum_early = nullptr;
  int n_colsum_early = 0, colsum_early_ctas = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float last_ms = 0.f;
  int launches = 0;
  std::vector<std::string> prof_names;
  b2g_sac_metrics* h_metrics_pinned = nullptr;
  float* h_met = nullptr;        // pinned MET_COUNT floats
  long long* h_cnt = nullptr;    // pinned counters

  b2g::V2State v2;
  // compact replay rows (engine v2, CNN policy): image planes + the ONE actuator value the policy reads (pixel [0,0] of the
  // last plane) + 3 pad floats; the rest of that plane is never read by augmented_nature_cnn (custom_obs_policy.py:28-30)
  bool compact = false;
  int Ec = 0;                        // floats per compact row
  float *cs_obs = nullptr, *cs_next = nullptr;     // compacted explicit batch [B][Ec]
  float* add_stage = nullptr;        // full-layout staging of replay_add chunks [2][ADD_CHUNK][E]
  double *d_mean_c = nullptr, *d_istd_c = nullptr; // statistics in the compact layout
  double* hp_stats[2]{};             // pinned staging of set_norm_stats (asynchronous upload, no stream sync)
  cudaEvent_t ev_stats[2]{};
  int stats_k = 0;
  bool v2_skip = false;          // (policy inference: the v1 forward groups write the separate z0 blocks)

  float* p(const std::string& n) { return P + tensors[tindex.at(n)].off; }
  float* g(const std::string& n) { return G + tensors[tindex.at(n)].off; }
};

