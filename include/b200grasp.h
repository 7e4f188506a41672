/*
 * b200grasp.h -- C ABI of the B200-native SAC learner (libb200grasp.so).
 *
 * Drop-in boundary for the ONE hot path of BarisYazici/deep-rl-grasping: the replay-buffer
 * minibatch gradient step that the reference delegates to stable_baselines.SAC (TF1) --
 *   constructed at  manipulation_main/training/sb_helper.py:104-128
 *   driven from     manipulation_main/training/sb_helper.py:175   (model.learn -> SAC._train_step)
 *   queried from    manipulation_main/utils.py:71                 (agent.predict)
 *   (de)serialised  manipulation_main/training/sb_helper.py:228-247, train_stable_baselines.py:95-104
 *
 * Conventions: every function returns 0 on success or a negative B2G_E* code and never throws
 * across the ABI; b2g_last_error() gives the message for the calling thread's last failure.  A
 * handle is single-owner and not thread-safe; it owns one CUDA stream and all of its device
 * memory.  Host arrays passed in are caller-owned and copied before the call returns (or, for the
 * *_async calls, before the next b2g_sync).  All floating point data is IEEE fp32 unless noted.
 * Parameter tensors use the TF variable names and layouts of the shipped SB zips (conv filters
 * HWIO, conv bias (1,n,1,1), dense kernels [in,out]) so get/set round-trips with them.
 */
#ifndef B200GRASP_H_
#define B200GRASP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2G_OK 0
#define B2G_EINVAL (-1)   /* bad argument / unknown variable name / size mismatch            */
#define B2G_ECUDA (-2)    /* CUDA runtime error (no device, launch failure, out of memory)   */
#define B2G_ESTATE (-3)   /* call not valid in this state (e.g. sampling from an empty buffer) */
#define B2G_ENCCL (-4)    /* NCCL unavailable or failed                                      */

/* precision modes of the dense contractions (convs, cnn_fc1, fc0 layers) */
#define B2G_PREC_FP32_SIMT 0   /* fp32 FFMA on CUDA cores: bit-faithful fp32 arithmetic            */
#define B2G_PREC_BF16X3 1      /* tcgen05 BF16 hi/lo split, 3 MMAs, fp32 TMEM accumulate (~2^-16)  */
#define B2G_PREC_BF16 2        /* tcgen05 single-pass BF16 (fast mode; tolerance reported)         */

typedef struct b2g_sac b2g_sac;

/* Replaces the keyword arguments of sb.SAC(policy, env, policy_kwargs, gamma, buffer_size,
 * batch_size, learning_rate, ...) -- sb_helper.py:120-128 -- plus the shapes the policy class
 * would read from env.observation_space / action_space (robot.py:207-228). */
typedef struct b2g_sac_cfg {
  int32_t obs_h, obs_w, obs_c; /* CNN policy: NHWC obs, obs_c = image channels + 1 feature plane
                                  (custom_obs_policy.py:28-32).  obs_h == 0 selects the MLP policy */
  int32_t obs_dim;             /* MLP policy: flat observation size (101 for the encoder config)   */
  int32_t n_act;               /* 5 (actuator.py:72-73); <= 8                                      */
  int32_t hidden;              /* SAC.layers = [hidden, hidden] (config/gripper_grasp.yaml:81); 64 */
  int32_t batch;               /* per-rank minibatch size                                          */
  int64_t buffer_capacity;     /* replay slots on this rank (config/gripper_grasp.yaml:82)         */
  float gamma, tau, target_entropy;
  uint64_t seed;               /* replay-index and policy-noise streams                            */
  int32_t precision;           /* B2G_PREC_*                                                       */
  int32_t device;              /* CUDA ordinal                                                     */
  int32_t rank, nranks;        /* data-parallel group; nranks == 1 -> no collective                */
  const void* nccl_id;         /* 128-byte ncclUniqueId shared by all ranks (nranks > 1)           */
  const char* nccl_lib;        /* optional path of libnccl.so.2 to dlopen; NULL = default search   */
} b2g_sac_cfg;

/* What SAC._train_step returns / logs (logs.csv columns policy_loss, qf1_loss, qf2_loss,
 * value_loss, entropy, ent_coef, ent_coef_loss) plus the parity scalars north_star names. */
typedef struct b2g_sac_metrics {
  float policy_loss, qf1_loss, qf2_loss, value_loss, ent_coef_loss, entropy, ent_coef;
  float grad_norm_pi, grad_norm_values, grad_ent;
  float mean_q1, mean_q2, mean_v, mean_logp;
  int64_t n_updates;
} b2g_sac_metrics;

const char* b2g_last_error(void);
int b2g_version(void);
/* writes a fresh 128-byte ncclUniqueId (rank 0 calls this, then shares it out of band) */
int b2g_nccl_unique_id(void* out128, const char* nccl_lib);

int b2g_sac_create(const b2g_sac_cfg* cfg, b2g_sac** out);
int b2g_sac_destroy(b2g_sac* h);
/* Peer-memory data parallelism (nranks > 1, one process per GPU of one NVLink node).  Every rank exports B2G_DP_EXPORT_BYTES
 * (CUDA IPC handles of its parameter arena, gradient receive arena and exchange block), the caller gathers the nranks blobs in rank
 * order (any out-of-band channel: the Python layer uses torch.distributed) and hands the concatenation to every rank.
 * From then on the optimiser launch of each gradient step is the collective: reduce-scatter of the gradients through peer
 * stores, Adam / Polyak on the owned slice, all-gather of the new parameters through peer stores -- no NCCL call on the path.
 * Replaces the all-reduce the reference's data-parallel wrapper would issue (reference: none -- SB 2.10 SAC is single-process;
 * SURVEY.md section 8e defines the N > 1 semantics this implements). */
#define B2G_DP_EXPORT_BYTES 192
int b2g_sac_dp_export(b2g_sac* h, void* out192);
int b2g_sac_dp_connect(b2g_sac* h, const void* all_exports /* nranks x B2G_DP_EXPORT_BYTES, rank order */, int nranks);
int b2g_debug_dp_stamps(b2g_sac* h, long long* out5);
/* host-only: [n][hw][cfull] observations -> compact replay rows [n][hw*(cfull-1)+4] (image planes | actuator value | 3 x 0), the layout
 * b2g_replay_add / b2g_sac_step_host_pipelined store and copy; needs no device */
int b2g_debug_compact_host(const float* src, float* dst, int n, int hw, int cfull, int threads);   /* bring-up: %globaltimer at the phase boundaries of the last launch */
int b2g_sync(b2g_sac* h);

/* ---- parameters: SB-zip variable names without the ":0" suffix (get_parameters / load_parameters,
 *      sb_helper.py:114-115) */
int b2g_param_count(const b2g_sac* h);
int b2g_param_info(const b2g_sac* h, int idx, const char** name, int64_t* numel, int32_t* ndim, int64_t shape[4]);
int b2g_get_param(b2g_sac* h, const char* name, float* dst, size_t numel);
int b2g_set_param(b2g_sac* h, const char* name, const float* src, size_t numel);
int b2g_get_grad(b2g_sac* h, const char* name, float* dst, size_t numel); /* gradient of the last step */
int b2g_get_adam(b2g_sac* h, const char* name, float* m, float* v, size_t numel);
int b2g_reset_optimizer(b2g_sac* h);   /* zero Adam moments and step counters (fresh tf.Session) */

/* ---- replay buffer (ReplayBuffer.add; stores UN-normalised obs/reward as SB does when a
 *      VecNormalize wraps the env) and VecNormalize statistics used at sample time
 *      (sb_helper.py:118-119; float64 like numpy) */
int b2g_replay_add(b2g_sac* h, const float* obs, const float* act, const float* rew, const float* next_obs,
                   const float* done, int64_t n);
int64_t b2g_replay_size(const b2g_sac* h);
/* ReplayBuffer.storage[slot] ([SB2] common/buffers.py): one stored (raw) transition back to the host; any output may be
 * NULL.  slot in [0, b2g_replay_size). */
int b2g_replay_get(b2g_sac* h, int64_t slot, float* obs, float* act, float* rew, float* next_obs, float* done);
/* What the LAST gradient step (any entry point, the CUDA-graph path included) drew and produced: the replay slots
 * indices[batch] (sampled steps only), the policy noise eps[batch, n_act], the per-sample rows q1,q2,v,logp,v_targ,
 * q1_pi,q2_pi (7 x [batch]) and the squashed actions pi[batch, n_act].  Any pointer may be NULL.  This is what lets a
 * test replay the very batch of a sampled step in the oracle. */
int b2g_get_last_batch(b2g_sac* h, int32_t* indices, float* eps, float* per_sample, float* pi_out);
int b2g_set_norm_stats(b2g_sac* h, const double* obs_mean, const double* obs_var, double ret_var, double clip_obs,
                       double clip_rew, double eps, int norm_obs, int norm_reward);

/* ---- the hot path.  One call = n_steps x { sample -> normalise -> fwd -> bwd -> [allreduce] ->
 *      3x Adam -> Polyak }  (SAC._train_step + target_update_op).  Indices and policy noise come
 *      from the handle's counter-based generators.  metrics (may be NULL) = last step. */
int b2g_sac_step(b2g_sac* h, int n_steps, float lr, b2g_sac_metrics* out);
/* Same, but metrics stay on the device until b2g_sync/next blocking call (no host round trip). */
int b2g_sac_step_async(b2g_sac* h, int n_steps, float lr);

/* Parity entry point: the caller supplies the RAW batch (host pointers; normalised on the device
 * with the current statistics) and the N(0,1) noise eps[batch, n_act], so results are comparable
 * with the oracle.  per_sample (may be NULL) receives q1,q2,v,logp,v_targ,q1_pi,q2_pi as 7 rows of
 * [batch]; pi_out (may be NULL) receives tanh-squashed actions [batch, n_act].
 * apply_update == 0 computes losses and gradients only. */
int b2g_sac_step_explicit(b2g_sac* h, const float* obs, const float* act, const float* rew, const float* next_obs,
                          const float* done, const float* eps, float lr, int apply_update, b2g_sac_metrics* out,
                          float* per_sample, float* pi_out);

/* Same work as b2g_sac_step_explicit(apply_update = 1) for a caller that keeps its replay buffer on the HOST
 * (as stable-baselines does): the step is enqueued and the call returns the losses of the PREVIOUSLY enqueued step
 * (*have_prev = 0 on the first call), so the host->device copy of step k overlaps the compute of step k-1.  The
 * host arrays must stay valid until the next call or b2g_sac_pipeline_flush (use pinned memory for true overlap). */
int b2g_sac_step_host_pipelined(b2g_sac* h, const float* obs, const float* act, const float* rew, const float* next_obs,
                                const float* done, const float* eps, float lr, b2g_sac_metrics* prev_out, int* have_prev);
/* waits for the last pipelined step and returns its losses */
int b2g_sac_pipeline_flush(b2g_sac* h, b2g_sac_metrics* last_out);

/* policy_tf.step (SAC.predict, utils.py:71): obs are RAW, normalised with the current stats.
 * deterministic -> tanh(mu); else tanh(mu + eps*std) with eps from the handle's generator. */
int b2g_sac_act(b2g_sac* h, const float* obs, int n, int deterministic, float* act_out);

/* number of kernel launches one gradient step issues (bench.py's gpu_launches) */
int b2g_launches_per_step(const b2g_sac* h);
/* device-time of the last b2g_sac_step call measured with CUDA events on the handle's stream (ms) */
float b2g_last_step_ms(const b2g_sac* h);
/* per-kernel-group device time of ONE extra profiled step (events around each launch);
 * names/ms arrays of capacity cap; returns the number of groups */
int b2g_profile_step(b2g_sac* h, float lr, const char** names, float* ms, int cap);

/* ------------------------------------------------------------------------------------------------
 * BDQ (branching dueling Q-network) learner -- the `sb.BDQ` object of train_stable_baselines.py:103-104 and
 * sb_helper.py:202-226 (author's fork `bdq_sb`, absent from the reference tree: parity unpinned).
 * Variable names / shapes follow trained_models/BDQ_8pads/BDQ_simple_8pads.zip.  Actions are stored as branch
 * bin indices (floats holding integers); bin k of a branch maps to linspace(-1, 1, n_bins)[k].
 * ------------------------------------------------------------------------------------------------ */
typedef struct b2g_bdq b2g_bdq;
typedef struct b2g_bdq_cfg {
  int32_t obs_dim;             /* 100 in the shipped zips                                                */
  int32_t n_branches;          /* action dimensions (3 simplified / 5 full); <= 8                        */
  int32_t n_bins;              /* num_actions_pad (config/gripper_grasp.yaml:112)                        */
  int32_t trunk0, trunk1;      /* layers[0] = common_net                                                 */
  int32_t branch_hidden;       /* layers[1] = layers[2] (branch and state-value hidden width)            */
  int32_t batch;
  int64_t buffer_capacity;
  float gamma;
  int32_t target_update_freq;  /* hard copy every N updates (target_network_update_freq)                 */
  int32_t trunk_grad_rescale;  /* 1: scale the gradient entering the trunk by 1/(n_branches+1) (paper)   */
  uint64_t seed;
  int32_t device;
  int32_t rank, nranks;        /* data-parallel group (BASELINE config 4): each rank owns a replay shard, one NCCL
                                  all-reduce averages the gradients; nranks == 1 -> no collective                    */
  const void* nccl_id;         /* 128-byte ncclUniqueId shared by all ranks (nranks > 1)                           */
  const char* nccl_lib;        /* optional libnccl path; NULL = default search                                     */
  int32_t prioritized_replay;  /* 1: proportional prioritised replay (zip data: prioritized_replay True,
                                  config/simplified_object_picking.yaml:108-110): device sum / min segment trees   */
  float per_alpha, per_eps;    /* priority exponent (0.6) and the epsilon added to |TD| (1e-6)                     */
} b2g_bdq_cfg;
typedef struct b2g_bdq_metrics {
  float loss, mean_q, grad_norm;
  int64_t n_updates;
} b2g_bdq_metrics;

int b2g_bdq_create(const b2g_bdq_cfg* cfg, b2g_bdq** out);
int b2g_bdq_destroy(b2g_bdq* h);
int b2g_bdq_param_count(const b2g_bdq* h);
int b2g_bdq_param_info(const b2g_bdq* h, int idx, char* name, size_t name_cap, int64_t* rows, int64_t* cols, int32_t* ndim);
int b2g_bdq_get_param(b2g_bdq* h, const char* name, float* dst, size_t numel);
int b2g_bdq_set_param(b2g_bdq* h, const char* name, const float* src, size_t numel);
int b2g_bdq_get_grad(b2g_bdq* h, const char* name, float* dst, size_t numel);
int b2g_bdq_replay_add(b2g_bdq* h, const float* obs, const float* act_idx, const float* rew, const float* next_obs,
                       const float* done, int64_t n);
int64_t b2g_bdq_replay_size(const b2g_bdq* h);
int b2g_bdq_set_norm_stats(b2g_bdq* h, const double* obs_mean, const double* obs_var, double ret_var, double clip_obs,
                           double clip_rew, double eps, int norm_obs, int norm_reward);
/* n_steps x { uniform sample -> forward (online s, online s', target s') -> double-Q TD loss -> backward -> Adam ->
 * hard target copy every target_update_freq updates } */
int b2g_bdq_step(b2g_bdq* h, int n_steps, float lr, b2g_bdq_metrics* out);
/* prioritised replay: importance-sampling exponent beta of the NEXT sampled steps (SB anneals beta0 -> 1 over
 * prioritized_replay_beta_iters); last_out (may be NULL) = slots[batch], weights[batch], new priorities[batch] of the last
 * sampled step, for inspection / tests */
int b2g_bdq_set_per_beta(b2g_bdq* h, float beta);
int b2g_bdq_get_last_per(b2g_bdq* h, int32_t* slots, float* weights, float* priorities);
/* parity entry point: caller-supplied batch (+ optional importance weights); td_out (may be NULL): [batch, n_branches] */
int b2g_bdq_step_explicit(b2g_bdq* h, const float* obs, const float* act_idx, const float* rew, const float* next_obs,
                          const float* done, const float* weights, float lr, int apply_update, b2g_bdq_metrics* out,
                          float* td_out);
/* greedy branch indices argmax_n Q_d(s, n) of the online network for n observations */
int b2g_bdq_act(b2g_bdq* h, const float* obs, int n, int32_t* act_idx_out);

/* ------------------------------------------------------------------------------------------------------------
 * Row a12: auto-encoder ENCODER forward (perception for the `encoded depth` observation).
 * Replaces SimpleAutoEncoder.encode  (/root/reference/manipulation_main/gripperEnv/encoders.py:59-61; graph :87-108)
 * called per env step from EncodedDepthImgSensor.get_state (manipulation_main/gripperEnv/sensor.py:218-222).
 * Layer spec = config.yaml `network` (filters / kernel_size / strides, padding 'same'), LeakyReLU(alpha) after every
 * conv and after Dense(encoding_dim).  Weights are the Keras arrays from model.h5: conv kernels [k,k,in,out], dense
 * kernel [flat,out] (flatten order H,W,C), biases [out].
 * ------------------------------------------------------------------------------------------------------------ */
#define B2G_ENC_MAX_LAYERS 8
typedef struct b2g_encoder b2g_encoder;
typedef struct b2g_encoder_cfg {
  int32_t height, width, channels;      /* Input(shape=(64, 64, 1)) in the reference */
  int32_t n_layers;                     /* conv layers */
  int32_t filters[B2G_ENC_MAX_LAYERS];
  int32_t kernel[B2G_ENC_MAX_LAYERS];
  int32_t strides[B2G_ENC_MAX_LAYERS];
  int32_t encoding_dim;
  float alpha;                          /* LeakyReLU slope, config.get('alpha', 0.1) */
  int32_t max_batch;
  int32_t device;
} b2g_encoder_cfg;
int b2g_encoder_create(const b2g_encoder_cfg* cfg, b2g_encoder** out);
int b2g_encoder_destroy(b2g_encoder* h);
int b2g_encoder_n_layers(const b2g_encoder* h);                      /* conv layers + 1 (dense) */
int b2g_encoder_layer_shape(const b2g_encoder* h, int layer, int64_t* kernel_numel, int64_t* bias_numel);
int b2g_encoder_set_weights(b2g_encoder* h, int layer, const float* kernel, size_t kernel_numel, const float* bias,
                            size_t bias_numel);
/* imgs: host [n, height, width, channels] fp32 -> out: host [n, encoding_dim]; B2G_ESTATE until every layer is loaded */
int b2g_encoder_encode(b2g_encoder* h, const float* imgs, int n, float* out);

/* ------------------------------------------------------------------------------------------------------------
 * Bring-up hook (not on the product path): C[M,N] = A[M,K] * B[N,K]^T through the tcgen05 engine; host pointers,
 * K a multiple of 8; x3 != 0 -> BF16 hi/lo split (3 MMAs); split_k > 1 -> that many partial accumulators summed
 * with fp32 atomics.  tools/tc_accum_probe.py uses it to measure the accumulation behaviour of the tensor core.
 * ------------------------------------------------------------------------------------------------------------ */
int b2g_debug_gemm(int M, int N, int K, const float* A, const float* B, float* C, int x3, int split_k);

#ifdef __cplusplus
}
#endif
#endif /* B200GRASP_H_ */
