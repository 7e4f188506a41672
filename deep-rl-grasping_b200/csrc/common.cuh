// Shared declarations of libb200grasp (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace b2g {

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch: a kernel launched through launch_pdl() may start (smem carve-up,
// barrier init, TMEM allocation) while its predecessor on the stream is still running; it must call
// pdl_wait() before touching global memory, and pdl_trigger() lets ITS successor start early.
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif
bool pdl_enabled();   // sac.cu: B2G_PDL != 0
// NCCL through dlopen (sac.cu); 0 or a negative B2G_E* code with b2g_last_error() set
int nccl_comm_init(void** comm, int nranks, const void* id128, int rank, const char* lib);
int nccl_allreduce_sum_f32(void* comm, float* buf, size_t count, cudaStream_t s);
void nccl_comm_destroy(void* comm);

// ---------------------------------------------------------------------------------------------
// Gather-GEMM problem descriptor.  One engine serves every dense contraction on the path:
//   C[cM[m] + cN[n]]  (=|+=)  epi( sum_r  A[aM[m] + aR[r]] * B[bR[r] + bN[n]] )
// The offset tables (built once on the host, resident in HBM/L2) encode im2col for the forward
// convolutions, the transposed/patch-major views for wgrad, and the parity-class gathers over
// zero-bordered gradient maps for dgrad, so no im2col matrix is ever materialised.
// ---------------------------------------------------------------------------------------------
enum GemmFlags : int {
  GG_A_RVEC = 1 << 0,     // A offsets contiguous along r in aligned groups of 4 (else along m)
  GG_B_RVEC = 1 << 1,     // B offsets contiguous along r in aligned groups of 4 (else along n)
  GG_EPI_BIAS_RELU = 1 << 2,
  GG_EPI_MASK = 1 << 3,   // C = acc * (mask[kM[m] + kN[n]] > 0)
  GG_EPI_ATOMIC = 1 << 4, // split-R accumulate into pre-zeroed C
  GG_COLSUM = 1 << 5,     // colsum[n] += sum_r B(r, n)   (bias gradients; tile_m == 0 only)
  GG_PLANES = 1 << 6,     // operands come pre-split as BF16 hi/lo planes (A_hi.., B_hi..), copied by cp.async
  GG_A_ALIGN4 = 1 << 7,   // plane A rows are only 8-byte aligned (conv1 with one image channel)
  GG_EPI_BIAS = 1 << 10,  // C = acc + bias[n] (no activation; output layers)
  GG_EPI_SCALE = 1 << 11, // C = acc * alpha (after mask)
  GG_A_SCALAR = 1 << 12,  // fp32 engine, with GG_A_RVEC: no 4-element contiguity along r -> element-wise gather (1-channel convs)
  GG_A_ROWLANES = 1 << 14, // planes K-major producer: lanes walk 8 consecutive rows of one 16-byte column group (conv1: adjacent
                           // output pixels overlap in the image, so a warp copy touches 4-8 lines instead of 32)
  GG_EPI_BIAS_LRELU = 1 << 13, // C = leaky_relu(acc + bias[n], slope alpha)   (Keras LeakyReLU; encoder.cu)
  GG_MN_MAJOR = 1 << 9,   // planes mode, wgrad: both operands contiguous along their M / N index -> MN-major UMMA tiles
  GG_CN_AFFINE4 = 1 << 8, // host-verified: cN / kN contiguous inside aligned 4-column groups, outputs 16-byte aligned
};

struct GemmDesc {
  const float* A;
  const float* B;
  float* C;
  const int* aM; const int* aR;
  const int* bR; const int* bN;
  const int* cM; const int* cN;
  const int* kM; const int* kN;
  const float* bias;
  const float* mask;
  float* colsum;
  // BF16 hi/lo planes (same element offsets as the fp32 tensors; B planes may use their own tables)
  const uint16_t* A_hi; const uint16_t* A_lo;
  const uint16_t* B_hi; const uint16_t* B_lo;
  const int* bR_p; const int* bN_p;
  uint16_t* C_hi; uint16_t* C_lo;      // optional plane copy of the output (feeds the next contraction)
  float alpha;                         // GG_EPI_SCALE
  int M, N, R;
  int flags;
  int splitR;
  int tiles_m, tiles_n;
  int tile_start;   // first flattened CTA index of this problem inside a grouped launch
  int tile_count;
  int col_id;       // tcgen05 engine: identity of (cN, kN, bias, N); equal ids share the staged column tables
  int layer;        // host side: layer index inside a fused (layer-synchronised) launch
  int need_done;    // fused launch: tiles [0, need_done) of the launch must be complete before this problem's operands are read
};

struct GemmGroup {          // one grouped launch
  std::string name;
  std::vector<GemmDesc> host;
  GemmDesc* dev = nullptr;
  int total_tiles = 0;
  double flops = 0;
  bool tc = false;          // run on the tcgen05 engine
  bool tc_eligible = false; // large dense contraction (convs, cnn_fc1)
  int* dev_ranges = nullptr; // tcgen05 engine: contiguous tile range per CTA (gg_tc_ranges)
  int ranges_grid = 0;
  bool layer_sync = false;   // several dependent layers in ONE persistent launch (in-kernel completion counter)
};

// engines (gg_simt.cu / gg_tc.cu)
void gg_simt_launch(const GemmDesc* dev_descs, int ndesc, int total_tiles, cudaStream_t s);
constexpr int GG_SIMT_BM = 64, GG_SIMT_BN = 64, GG_SIMT_BK = 16;
// tcgen05 engine: 128 x 64 output tile, 64-wide r-chunks; x3 != 0 -> BF16 hi/lo split (3 MMAs)
cudaError_t gg_tc_launch(const GemmDesc* host_descs, int ndesc, int total_tiles, int mode_flags, int x3, int num_sms, cudaStream_t s,
                         const int* dev_ranges = nullptr, int ranges_grid = 0, unsigned* sync_ctr = nullptr);
// host side of the contiguous tile schedule: cost-balanced range boundaries [grid + 1] for a finalized group
std::vector<int> gg_tc_ranges(const GemmDesc* host_descs, int ndesc, int total_tiles, int grid);
constexpr int GG_TC_MAX_DESCS = 16;
int gg_tc_smem_bytes();
extern long long* g_tc_trace;   // bring-up hook (gg_tc.cu)
constexpr int GG_TC_BM = 128, GG_TC_BN = 64, GG_TC_BK = 64;

// ---------------------------------------------------------------------------------------------
// head "tail" kernel (tail.cu): everything after the fc0 contractions, per sample
// ---------------------------------------------------------------------------------------------
struct HeadW {           // pointers into the parameter arena for one MLP head
  const float* b0;       // fc0 bias [H]
  const float* k1;       // fc1 kernel [H,H]
  const float* b1;       // fc1 bias [H]
  const float* ko;       // output kernel [H, n_out]
  const float* bo;       // output bias [n_out]
  const float* k0;       // fc0 kernel [in, H] (qf heads: rows feat_dim.. are the action rows)
};
struct HeadG {           // matching gradient-arena pointers (small tensors accumulated by the tail)
  float* b1; float* ko; float* bo;
};

struct TailArgs {
  int B, H, A, feat_dim;
  float gamma, target_entropy;
  int grad_scale_B;            // divide means by this batch size (local batch)
  // fc0 pre-activations (no bias) [B,H] each
  const float* z0_pi; const float* z0_vf; const float* z0_q1; const float* z0_q2; const float* z0_vt;
  int z0v_ld;                  // row stride of z0_vf / z0_q1 / z0_q2 (H, or 3H when the three heads share one [B,3H] block)
  HeadW pi, vf, q1, q2, vt;
  const float* ksig; const float* bsig;      // pi: dense_1 (log_std) kernel/bias; pi.ko/bo = dense (mu)
  HeadG g_pi, g_vf, g_q1, g_q2;
  float* g_ksig; float* g_bsig;
  const float* log_alpha; float* g_log_alpha;
  const float* act;  int act_stride;         // replay actions (inside F_V rows)
  const float* eps;                          // [B,A]
  const float* rew; const float* done;       // normalised reward, done [B]
  // saved for the engine: post-ReLU fc0 activations and gradients
  float* a0_pi; float* a0_vf; float* a0_q1; float* a0_q2;   // [B,H]
  float* dz1_pi; float* dz1_vf; float* dz1_q1; float* dz1_q2; // [B,H]
  float* dz0_pi;                               // [B,H]
  float* dz0_v3;                               // [B,3H] = vf | q1 | q2
  uint16_t* dz0_pi_p[2]; uint16_t* dz0_v3_p[2];  // optional BF16 hi / lo planes of the same (engine v2 operands)
  float* per_sample;                           // 7 x [B]: q1,q2,v,logp,v_targ,q1_pi,q2_pi
  float* pi_out;                               // [B,A]
  float* metrics;                              // accumulators (see MET_* in sac.cu)
};
void tail_launch(const TailArgs& a, cudaStream_t s);

// Weight gradients of the head MLPs (fc0 / fc1 kernels and biases of pi, vf, qf1, qf2) in fp32 on the CUDA cores: 76 MFLOP
// of [B]-deep reductions, too small for a tensor-engine launch (which would also hold every SM while it runs).
struct HeadsWgradArgs {
  const float* X0[4];     // fc0 inputs: feature rows [B][x0_ld] (pi: F_pi; vf, qf1, qf2: F_values incl. the action columns)
  const float* dz0[4];    // fc0 pre-activation gradients [B][dz0_ld[q]] (column offset applied)
  const float* a0[4];     // fc0 activations [B][64]
  const float* dz1[4];    // fc1 pre-activation gradients [B][64]
  float* g_k0[4]; float* g_b0[4]; float* g_k1[4]; float* g_b1[4];
  int M0[4];              // fc0 input width of head q (kernel rows)
  int dz0_ld[4];
  int x0_ld, B;
};
void heads_wgrad_launch(const HeadsWgradArgs& a, cudaStream_t s);
// policy inference tail: tanh(mu) or tanh(mu + eps*std) for the first n rows of z0_pi
void act_launch(const TailArgs& t, int n, int deterministic, float* act_out, cudaStream_t s);

enum Metric : int {
  MET_POLICY_LOSS = 0, MET_QF1_LOSS, MET_QF2_LOSS, MET_VALUE_LOSS, MET_ENT_COEF_LOSS, MET_ENTROPY,
  MET_MEAN_Q1, MET_MEAN_Q2, MET_MEAN_V, MET_MEAN_LOGP, MET_GN_PI, MET_GN_VALUES, MET_COUNT = 16
};

// ---------------------------------------------------------------------------------------------
// optimiser (optim.cu): 3x TF-Adam + Polyak over the flat arenas, one launch
// ---------------------------------------------------------------------------------------------
struct OptimArgs {
  float* P; float* Mo; float* Vo; const float* G;   // trainable arenas
  float* T;                                          // target block (same relative layout as values block)
  int n_pi, n_values, n_ent;                         // padded segment lengths: [pi | values | ent]
  int n_target;                                      // padded length of the target block
  const double* step_consts;                         // [3] lr_t per optimiser (device, written by prep kernel)
  float tau;
  float grad_scale;                                  // 1/nranks after a sum all-reduce
  float* metrics;                                    // MET_GN_* accumulators
  int apply;                                         // 0 = only grad norms
  long long* bump_counter;                           // rng step counter advanced once per step (nullptr: prep did it)
  int r_lo[2], r_hi[2];                              // optional: update only arena ranges [r_lo, r_hi) (floats, multiples of 4);
                                                     // all zero = the whole arena
};
void optim_launch(const OptimArgs& a, cudaStream_t s);

// Data-parallel optimiser step over NVLink peer memory (optim.cu): reduce-scatter of the gradient arena, Adam / Polyak on the
// owned shard and all-gather of the updated parameters in ONE kernel.  Rank r owns the r-th 1/N of the arena: every rank pushes
// that slice of its gradients into r's receive arena (peer stores), r sums the N copies in fixed rank order (replicas stay
// bit-identical), updates its slice of P / m / v (the moments exist only on the owner) and stores the new parameters -- and the
// Polyak-averaged target slice -- into every rank's arena, again through peer stores.  Ranks meet twice through epoch flags in each other's memory (release /
// acquire at system scope): "my gradients are final" before the loads, "my slice is written everywhere" before the kernel ends.
constexpr int DP_MAX_RANKS = 8;
struct DpArgs {
  OptimArgs o;                         // local arenas, step sizes, metrics
  int rank, nranks;
  float* R_peer[DP_MAX_RANKS];         // receive arena of every rank: [src rank][slice] floats (slices pushed by their producers)
  float* P_peer[DP_MAX_RANKS];         // parameter arena of every rank
  int* x_peer[DP_MAX_RANKS];           // exchange block of every rank: int flags[2][8] (G ready, slice written), float part[8][2], loss[8][16]
  int skip_lo4[2], skip_hi4[2];        // float4 ranges of the arena whose slices were already pushed by the backward epilogues
  const long long* counters;           // counters[3] = optimiser step = flag epoch
  int* sync;                           // local: [0], [1] CTA arrival counters, [2..3] squared-norm accumulators (as float), [8..] bring-up stamps
};
void dp_optim_launch(const DpArgs& a, int ctas, cudaStream_t s);

// weights -> BF16 hi/lo planes, original [R,N] layout and transposed [N,R] (optim.cu)
struct PlaneJob {
  const float* src;            // [R, N] row-major fp32 (TF layout: HWIO filters flattened, dense [in,out])
  uint16_t* hi; uint16_t* lo;  // [R, N] planes (may be null)
  uint16_t* hiT; uint16_t* loT;// [N, R] planes (may be null)
  int R, N;
  int tile_start;              // first 32x32 tile of this job in the flattened launch
};
void planes_launch(const PlaneJob* dev_jobs, int njobs, int total_tiles, cudaStream_t s);
// bias gradients: dst[n] += sum over rows of src[row_off[m] + n]  (optim.cu)
struct ColsumJob { const float* src; const int* row_off; float* dst; int rows, N; int cta_start; };
void colsum_launch(const ColsumJob* dev_jobs, int njobs, int total_ctas, cudaStream_t s);

struct PrepArgs {          // 1-CTA kernel at the head of every step
  long long* counters;     // [0..2] Adam t per optimiser, [3] n_updates, [4] rng step counter
  double* step_consts;     // lr_t x3
  const float* lr;         // device scalar
  float* metrics;          // zeroed
  int* indices; float* eps;// generated when gen != 0
  int B, A; const long long* replay_size;   // nullptr -> counters[5]
  unsigned long long seed; int gen; int apply;
  int defer_bump;          // 1: the rng step counter [4] is advanced by the optimiser kernel at the end of the step
  int skip_indices;        // 1: the gather kernel draws the replay slots itself (same Philox stream)
};
void prep_launch(const PrepArgs& a, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// replay gather + VecNormalize + /255 (replay.cu)
// ---------------------------------------------------------------------------------------------
struct GatherArgs {
  const float* obs; const float* next_obs; const float* act; const float* rew; const float* done; // replay or staged batch
  const int* indices;        // [B] slot per sample (nullptr: identity)
  const double* mean; const double* var; // [obs_elems]; var[] holds 1/sqrt(var+eps)
  const double* normc;       // device: {1/sqrt(ret_var+eps), clip_obs, clip_rew, norm_obs, norm_rew}
  int B, H, W, Cfull;        // CNN: obs [H,W,Cfull]; MLP: H = 0, W = obs_dim
  float scale;               // 255 for CNN, 1 for MLP
  float* x_obs; float* x_next;   // CNN: [B,H,W,Cfull-1] image planes (scaled)
  uint16_t* x_obs_hi; uint16_t* x_obs_lo; uint16_t* x_next_hi; uint16_t* x_next_lo;   // optional BF16 planes of x
  float* F_pi; float* F_v; float* F_t; int FS; int feat_col; // feature rows: direct feature -> col feat_col; MLP: whole obs -> cols 0..
  float* rew_out; float* done_out; int n_act;
  // in-kernel slot draw (indices == nullptr && rng_counters != nullptr): Philox stream 0 of prep_kernel, same values
  const long long* rng_counters;   // [4] = rng step, [5] = replay size
  unsigned long long seed;
  int* indices_out;                // optional record of the drawn slots
};
void gather_launch(const GatherArgs& a, cudaStream_t s);

// Philox4x32-10 (counter-based RNG shared by prep_kernel and the in-kernel replay slot draw)
#ifdef __CUDACC__
__device__ __forceinline__ void philox_round(uint4& c, uint2& k) {
  const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
  const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
  c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
  k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
}
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) philox_round(c, k);
  return c;
}
// replay slot of sample b at rng step `step`: element (b & 3) of block b >> 2 of stream 0, scaled to [0, rsz)
__device__ __forceinline__ int philox_slot(unsigned long long seed, unsigned long long step, int b, unsigned long long rsz) {
  const uint4 r = philox4x32_10(make_uint4((unsigned)step, (unsigned)(step >> 32), (unsigned)(b >> 2), 0u),
                                make_uint2((unsigned)seed, (unsigned)(seed >> 32)));
  const unsigned v = (b & 3) == 0 ? r.x : (b & 3) == 1 ? r.y : (b & 3) == 2 ? r.z : r.w;
  return (int)(((unsigned long long)v * rsz) >> 32);
}
#endif

}  // namespace b2g
