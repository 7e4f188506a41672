// TMA-fed tcgen05 contraction engine ("cg"): declarations shared by cg.cu (kernel) and sac.cu (problem builders).
//
// Every dense contraction of the step is a list of 128-row output tiles; the operands of a tile are fetched, K-chunk by
// K-chunk, by cp.async.bulk.tensor (TMA) boxes over BF16 plane tensors straight into 128B-swizzled shared-memory UMMA
// tiles.  Convolutions need no im2col buffer: their patches / shifted windows / zero borders are expressed as tensor-map
// VIEWS (overlapping strides, element strides, out-of-bound zero fill) of the NHWC activation planes, so two elected
// lanes feed the whole ring (tools/tma_probe.cu checks each view behaviour on the device).  A launch is a list of problems;
// problems of one launch may depend on each other tile by tile (fused layers) and may split K with in-kernel finalisation.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2g {

constexpr int CG_MAX_LOADS = 4;     // TMA boxes per K-chunk and plane (A atoms + B atoms)
constexpr int CG_MAX_PROBLEMS = 16; // problems per grouped launch (the whole list travels as a __grid_constant__ kernel parameter)
constexpr int CG_MAX_STAGES = 8;
constexpr int CG_MAX_PLANES = 3;
constexpr int CG_EPI_WARPS = 8;    // epilogue warps of a CTA; each signals a finished tile once (dependency counters)

enum CgEpiKind : int {
  CG_EPI_ACT = 0,     // relu(acc + bias[n]) -> BF16 planes (activations; optional fp32 copy)
  CG_EPI_RAW = 1,     // acc -> fp32 (fc0 pre-activations)
  CG_EPI_DGRAD = 2,   // acc * (mask[row, n] > 0) -> BF16 planes (gradient maps)
  CG_EPI_WGRAD = 3,   // acc -> fp32 weight gradient, plain store or atomic accumulation (split-K)
};

struct CgLoad {
  int map;            // index of the tensor map (all planes of the tensor: plane = outermost dimension)
  int rank;           // tensor-map rank including the plane dimension (3..5)
  int box_bytes;      // bytes one plane of the box occupies (host-side check)
  int smem_off;       // byte offset of the plane-0 box inside a stage (A boxes below b_off, B boxes from b_off on)
  int plane_stride;   // bytes between the plane copies of the box inside the stage (filled by cg_finalize: a_pstride | b_pstride)
  int plane_box;      // 1: the box spans every plane -> one instruction per chunk, planes land box_bytes apart (== plane_stride);
                      // 0: one instruction per plane (plane = last coordinate)
  int c0[5];          // box start coordinates: c0 + tm*d_tm + tn*d_tn + c1*d_c1 + c2*d_c2 (+ tm_tab)
  int d_tm[5], d_tn[5], d_c1[5], d_c2[5];
};

struct CgProblem {
  // ---- tile grid: tile -> (split, tm, tn); K-chunk c = c1 * n2 + c2
  int tile_start, tiles_m, tiles_n, splits;
  int chunks, n2;
  // ---- operand fetch
  int nloads, planes;
  int a_pstride, b_pstride;   // stage layout: [A plane 0 | A plane 1 | ..][B plane 0 | B plane 1 | ..]; bytes of one A / B plane.
                              // The B planes are CONTIGUOUS so that one UMMA descriptor spans [B0|B1|B2] (see nprod)
  int tx_bytes;               // bytes all boxes of one stage deliver (planes x sum of box bytes)
  int slot_bytes, nstages;    // stage ring geometry of THIS problem (cg_finalize): the problems of a launch share the ring's bytes, not
                              // its partition -- the ring is drained when the partition changes
  CgLoad ld[CG_MAX_LOADS];
  const int* tm_tab;          // optional [tiles_m][CG_MAX_LOADS][2]: extra offsets of coordinates 1 and 2 per (tm, load)
  // ---- MMA
  int mn_major;               // 0: K-major A and B (rows = M|N, 128 B of K); 1: MN-major (rows = K, 128 B of M|N)
  int ksteps;                 // UMMA K = 16 steps per chunk
  int a_off, b_off;           // region offsets inside a stage (b_off = planes * a_pstride)
  int a_kstep, b_kstep;       // descriptor start-address advance per k-step (bytes)
  int a_lbo, b_lbo;           // MN-major: byte stride between 64-element atoms along M|N
  int umma_n;                 // tile width (multiple of 16, <= 256)
  int nprod;                  // products per k-step: 1 (hi*hi), 3 (+hi*lo, lo*hi), 6 (+mid terms of the 3-plane split), issued as
                              // `planes` wide MMAs: A_p x [B_0 | .. | B_(planes-1-p)] into accumulator columns [p*n, planes*n), so
                              // column group g collects the products of order 2^(-8g) (planes * umma_n <= 256)
  // ---- epilogue
  int epi;
  int rows_tile;              // real rows of a full tile (<= 128)
  int lim_rows;               // tm * rows_tile + r < lim_rows
  int d0, d1;                 // r -> i0 = r % d0, i1 = (r / d0) % d1, i2 = r / (d0 * d1)
  long long o_tm; int o0, o1, o2; long long o_base;    // output element offset of (tm, r)
  long long m_tm; int m0, m1, m2; long long m_base;    // mask element offset of (tm, r)
  int n_valid;                // columns < n_valid are stored (N of the problem)
  int grp_stride;             // element distance between consecutive 32-column groups of the output (32 = contiguous)
  int grp_tab;                // 1: per-group output / mask offsets and row limits come from the tables below (conv2 dgrad: one
                              //    accumulator column group per output-parity class)
  int grp_off[8], grp_moff[8], grp_lim0[8], grp_lim1[8];
  int out_planes;             // planes written by ACT / DGRAD
  uint16_t* out_p[CG_MAX_PLANES];
  float* out_f;               // fp32 output (RAW / WGRAD; optional extra copy for ACT), same offsets, ld = o0-based
  long long f_tm; int f0;     // fp32 copy: offset = tm * f_tm + r * f0 + col   (ACT extra copy only; 0 = none)
  const float* bias;          // [N]
  int bias_grp;               // element distance between the bias blocks of consecutive 32-column groups (32 = contiguous)
  long long f_grp;            // same for the fp32 copy
  const uint16_t* mask;       // hi plane of the forward activation (DGRAD)
  float* colsum;              // DGRAD: bias gradient of the layer = column sums of the masked gradient map, accumulated (red.add) from the
  int colsum_mask;            //        fp32 accumulators: colsum[(column) & colsum_mask]   (nullptr: none)
  int atomic;                 // WGRAD: 1 = red.add (split-K or shared output), 0 = store
  float scale;                // WGRAD: multiply before accumulation (1 = none)
  // ---- dependencies between the problems of ONE launch (fused layers).  Tiles are dealt to the CTAs in increasing order and
  //      every CTA of the grid is resident, so a tile may wait for lower-numbered tiles of an earlier problem: the producers of
  //      tile tm spin until the row-tiles [tm * dep_rows / dep_rows_tile, ((tm + 1) * dep_rows - 1) / dep_rows_tile] of the
  //      producing problem have each collected dep_expect arrivals (one per epilogue warp and (tn, split) tile), then order the
  //      generic-proxy stores they observed before their own async-proxy (TMA) reads.
  int* done_ctr;              // arrival counters of THIS problem, one per tm (nullptr: nobody waits on it); zeroed before the launch
  const int* dep_ctr;         // counters of the producing problem (nullptr: no dependency)
  int dep_rows, dep_rows_tile, dep_tiles, dep_expect;
  // ---- split-K with in-kernel finalisation (ACT problems with splits > 1; the late layers have too few tiles for 148 SMs): every
  //      split tile adds its fp32 partial sums into ws[tile][128][umma_n] (red.add), then each epilogue warp bumps its own arrival
  //      counter of the tile; the warp that arrives LAST reads the sums back, clears them for the next step, and runs the normal
  //      bias / ReLU / plane-split / store path (and alone signals done_ctr)
  float* ws;                  // nullptr: splits are plain (RAW / WGRAD atomics)
  int* ws_cnt;                // [tiles_m * tiles_n][CG_EPI_WARPS]
  // ---- data parallel (N > 1): a weight gradient that is FINAL when its tile is stored (no split-K) is also pushed, float4 by float4,
  //      into the receive arena of the rank that owns that part of the gradient arena (optim.cu: dp_optim_kernel) -- 80 % of the
  //      gradient bytes leave while the backward pass is still running
  float* dp_recv[8];          // receive arenas (nullptr entries: none)
  const float* dp_gbase;      // base of the local gradient arena (owner of float4 i4 = i4 / dp_per4)
  int dp_rank, dp_n, dp_per4;
  int dep_by_chunk;           // 1: the rows are indexed by the tile's K-chunk range [c_begin, c_end) instead of its tm (weight gradients)
};

struct CgGroup {               // one launch
  CgProblem host[CG_MAX_PROBLEMS];
  int n = 0;
  int total_tiles = 0;
  int slot_bytes = 0, nstages = 0, ring_bytes = 0;     // largest slot / most stages of any problem; bytes of the ring
  const char* name = "";
  double flops = 0;
};

// encodes a BF16 tiled tensor map (SWIZZLE_128B, zero OOB fill); dims/box innermost first, strides in BYTES for dims 1..rank-1
int cg_encode_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                  const uint32_t* elem_strides);
// finalises tile_start / total_tiles / ring geometry of a group (host side)
int cg_finalize(CgGroup& g, int smem_budget);
cudaError_t cg_launch(const CgGroup& g, const CUtensorMap* dev_maps, int num_sms, cudaStream_t s, bool pdl, int debug_flags);
int cg_smem_limit();

}  // namespace b2g
