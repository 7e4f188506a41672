// TMA-fed tcgen05 contraction engine (sm_100a).  See cg.cuh for the problem description.
//
// Persistent, warp-specialised kernel, one CTA per SM, grid = min(#tiles, #SMs):
//   warps 0..1  : TMA producers (whole warps, converged; one elected lane issues).  The cp.async.bulk.tensor boxes of a K-chunk
//                 (operands x planes) are dealt round-robin to the two warps; warp 0 posts the chunk's expect_tx.  More issuing
//                 warps do not help: with 2 or 4 the 72 KB of a 3-plane chunk take ~1400 cycles to land either way (51 B/clk
//                 per SM, ~14 TB/s chip-wide out of L2 -- profiles/cg_trace_r2.txt), which is the fabric, not the issue rate;
//   warp 2      : owns the TMEM allocation; converged warp, one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (BF16
//                 planes, fp32 accumulation in TMEM, two accumulator buffers so the epilogue of tile i overlaps the mainloop
//                 of tile i+1).  Precision modes per problem: 1 product (hi*hi), 3 products (2-plane split) or 6 products
//                 (3-plane split hi/mid/lo: everything down to 2^-24), issued as 1..3 WIDE MMAs per k-step (see the issue
//                 loop) -- profiles/precision_r2.md explains why the forward pass needs the 6-product mode;
//   warps 3..10 : epilogue.  tcgen05.ld one accumulator row per thread, apply bias/ReLU or the ReLU mask (+ the bias-gradient column
//                 sums), split into BF16 planes and store the row's 32 columns of every plane directly (16-byte vectors).
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>

#include "cg.cuh"
#include "common.cuh"

namespace b2g {
namespace {

// warps 0..1: TMA producers, warp 2: MMA issuer (owns TMEM), warps 3..10: epilogue
constexpr int NPROD_WARPS = 2, MMA_WARP = NPROD_WARPS;
constexpr int EPI_WARP0 = NPROD_WARPS + 1, NEPI_WARPS = CG_EPI_WARPS;
constexpr int NTHREADS = 32 * (EPI_WARP0 + NEPI_WARPS);
constexpr int TMEM_COLS = 512;          // two accumulators of up to 256 fp32 columns

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "CG_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra CG_DONE;\n\t"
      "bra CG_WAIT;\n\t"
      "CG_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
// one lane of a CONVERGED warp.  The producer and MMA roles run their loops with all 32 lanes so that every address,
// descriptor and coordinate is warp-uniform and lives in the uniform register file UTMALDG / UTCHMMA read; a role entered
// by one lane only (if (lane == 0) {...}) made the compiler wrap every such instruction in an R2UR + ELECT/BRA.U.ANY
// value-serialisation loop -- ~75 cycles per MMA issued (profiles/cg_trace_r2.txt).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor): start >> 4 in [0,14), LBO >> 4 in [16,30), SBO >> 4 in
// [32,46), version 1 in [46,48), layout type 2 = SWIZZLE_128B in [61,64).  K-major: rows of 128 B, 8-row groups 1024 B
// apart (SBO), LBO unused.  MN-major: rows = K, 128 B = 64 elements along M|N; SBO = stride between 8-row K groups
// (1024 B inside a TMA box), LBO = stride between 64-element atoms along M|N (one box each).
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D = f32 (bit 4), A = B = BF16 (bits 7, 10), MN-major A / B (bits 15, 16), N >> 3 at 17, M >> 4 at 24
__device__ __forceinline__ uint32_t make_idesc(int m, int n, bool mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (mn_major ? (3u << 15) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tma_load(uint32_t dst, const CUtensorMap* map, uint32_t bar, int rank, const int (&c)[5]) {
  switch (rank) {
    case 2:
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst), "l"(map),
                   "r"(bar), "r"(c[0]), "r"(c[1])
                   : "memory");
      break;
    case 3:
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                   "l"(map), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2])
                   : "memory");
      break;
    case 4:
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                   "l"(map), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3])
                   : "memory");
      break;
    default:
      asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(
                       dst),
                   "l"(map), "r"(bar), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4])
                   : "memory");
      break;
  }
}

// The problem list lives in the kernel-parameter constant bank: every role re-reads descriptor fields per tile, and a
// constant-cache hit costs tens of cycles where a global-memory descriptor cost a dependent L2 round trip per field
// group (the "skeleton" of the round-1 engine: profiles/tc_ablation_r1.txt).
struct CgPack { CgProblem p[CG_MAX_PROBLEMS]; };
struct Tile { int p, tm, tn, c_begin, c_end; };
// Tile walker: a CTA's tiles increase monotonically, so the problem index only moves forward and the tile-grid fields of
// the current problem stay in registers (re-read from the constant bank only when the problem changes).
struct Walker {
  int p = -1, next_start = 0, tile_start = 0, tiles_n = 1, per = 1, splits = 1, chunks = 0, cps = 0;
  __device__ __forceinline__ bool advance(const CgProblem* __restrict__ probs, int nprob, int tile) {   // true: problem changed
    bool changed = false;
    while (p < 0 || (p + 1 < nprob && tile >= next_start)) {
      ++p;
      const CgProblem& P = probs[p];
      tile_start = P.tile_start; tiles_n = P.tiles_n; per = P.tiles_m * P.tiles_n; splits = P.splits; chunks = P.chunks;
      cps = (chunks + splits - 1) / splits;
      next_start = p + 1 < nprob ? probs[p + 1].tile_start : 0x7fffffff;
      changed = true;
    }
    return changed;
  }
  __device__ __forceinline__ Tile tile(int t_abs) const {
    int t = t_abs - tile_start, split = 0;
    if (splits > 1) { split = t / per; t -= split * per; }
    Tile ti;
    ti.p = p;
    ti.tm = tiles_n == 1 ? t : t / tiles_n;
    ti.tn = t - ti.tm * tiles_n;
    ti.c_begin = split * cps;
    ti.c_end = min(chunks, ti.c_begin + cps);
    return ti;
  }
};

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {    // lo -> bits [0,16)
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}

__global__ void __launch_bounds__(NTHREADS, 1)
cg_kernel(const __grid_constant__ CgPack pk, int nprob, int total_tiles, const CUtensorMap* __restrict__ maps, int max_stages,
          int dbg, long long* __restrict__ trace, int epi_tiles) {
  const CgProblem* __restrict__ probs = pk.p;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[CG_MAX_STAGES], bar_empty[CG_MAX_STAGES], bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform for the compiler, not just in fact
  const int trace_cta = dbg >> 8;                    // bring-up: the CTA whose roles write clock stamps
  const bool dbg_noload = dbg & 1, dbg_nomma = dbg & 2, dbg_nostore = dbg & 4, dbg_nost = dbg & 16;    // 16: epilogue math without the global stores
  if (tid == 0) {
    for (int s = 0; s < max_stages; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&bar_acc_full[b]), 1); mbar_init(smem_u32(&bar_acc_empty[b]), epi_tiles ? NEPI_WARPS / 2 : NEPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  pdl_trigger();
  pdl_wait();

  if (warp < NPROD_WARPS) {
    // ============================================================================================ TMA producers
    {
      // Ring slot s and the phase parity of every slot (bit s of ph): the partition of the ring (slot size, slot count) belongs
      // to the problem, so a slot's barrier may have completed a different number of phases than its neighbours'.
      uint32_t gc = 0, s = 0, ph = 0;
      int slot_bytes = 0, nstages = 1;
      Walker w;
      int n2 = 1, nloads = 0, planes = 0, tx = 0;
      const int* tab = nullptr;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        if (w.advance(probs, nprob, tile)) {
          const CgProblem& P = probs[w.p];
          n2 = P.n2; nloads = P.nloads; planes = P.planes; tx = P.tx_bytes; tab = P.tm_tab;
          if (P.slot_bytes != slot_bytes || P.nstages != nstages) {
            // new partition: every slot of the old one must have been consumed before its bytes are overwritten (a fresh
            // barrier passes the parity-1 wait at once, so never-used slots cost nothing)
            for (int q = 0; q < nstages; ++q) mbar_wait(smem_u32(&bar_empty[q]), ((ph >> q) & 1u) ^ 1u);
            slot_bytes = P.slot_bytes; nstages = P.nstages; s = 0;
          }
        }
        const Tile ti = w.tile(tile);
        const CgProblem& P = probs[ti.p];
        if (P.dep_ctr) {                                                 // fused layers: wait for the tiles this one reads
          const int* __restrict__ ctr = P.dep_ctr;
          const int x0 = P.dep_by_chunk ? ti.c_begin : ti.tm, x1 = P.dep_by_chunk ? ti.c_end : ti.tm + 1;
          const int lo = x0 * P.dep_rows / P.dep_rows_tile, hi = min(P.dep_tiles - 1, (x1 * P.dep_rows - 1) / P.dep_rows_tile);
          const int expect = P.dep_expect * (epi_tiles ? NEPI_WARPS / 2 : NEPI_WARPS);
          for (int j = lo + lane; j <= hi; j += 32) {
            int seen;
            do {
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(ctr + j) : "memory");
              if (seen < expect) __nanosleep(100);
            } while (seen < expect);
          }
          __syncwarp();
          asm volatile("fence.proxy.async.global;" ::: "memory");
        }
        // per-tile box origins: everything but the chunk terms
        int base[CG_MAX_LOADS][5];
#pragma unroll
        for (int l = 0; l < CG_MAX_LOADS; ++l) {
          if (l < nloads) {
            const CgLoad& L = P.ld[l];
#pragma unroll
            for (int d = 0; d < 5; ++d) base[l][d] = L.c0[d] + ti.tm * L.d_tm[d] + ti.tn * L.d_tn[d];
            if (tab) { base[l][1] += tab[(ti.tm * CG_MAX_LOADS + l) * 2]; base[l][2] += tab[(ti.tm * CG_MAX_LOADS + l) * 2 + 1]; }
          }
        }
        int c1 = n2 > 1 ? ti.c_begin / n2 : 0, c2 = ti.c_begin - c1 * n2;
        for (int c = ti.c_begin; c < ti.c_end; ++c, ++gc) {
          const bool tr = trace && blockIdx.x == trace_cta && warp == 0 && gc < 64 && lane == 0;
          if (tr) trace[gc * 8 + 0] = clock64();
          mbar_wait(smem_u32(&bar_empty[s]), ((ph >> s) & 1u) ^ 1u);
          if (tr) trace[gc * 8 + 1] = clock64();
          const uint32_t full = smem_u32(&bar_full[s]);
          const bool leader = elect_one();
          if (dbg_noload) { if (warp == 0 && leader) mbar_arrive(full); }
          else {
            if (warp == 0 && leader) mbar_expect_tx(full, (uint32_t)tx);        // the one arrival of the phase; boxes may land before it
            const uint32_t sbase = ring + s * (uint32_t)slot_bytes;
            int j = 0;                                                 // box index inside the chunk, dealt round-robin to the producer warps
#pragma unroll
            for (int l = 0; l < CG_MAX_LOADS; ++l) {
              if (l < nloads) {
                const CgLoad& L = P.ld[l];
                int crd[5];
#pragma unroll
                for (int d = 0; d < 5; ++d) crd[d] = base[l][d] + c1 * L.d_c1[d] + c2 * L.d_c2[d];
                if (L.plane_box) {                               // all planes in one box (plane = outermost coordinate, 0)
                  if ((j & (NPROD_WARPS - 1)) == warp && leader) tma_load(sbase + (uint32_t)L.smem_off, maps + L.map, full, L.rank, crd);
                  ++j;
                } else {
                  for (int pl = 0; pl < planes; ++pl, ++j) {
#pragma unroll
                    for (int d = 2; d < 5; ++d) if (d == L.rank - 1) crd[d] = pl;
                    if ((j & (NPROD_WARPS - 1)) == warp && leader)
                      tma_load(sbase + (uint32_t)(pl * L.plane_stride + L.smem_off), maps + L.map, full, L.rank, crd);
                  }
                }
              }
            }
          }
          if (tr) { trace[gc * 8 + 2] = clock64(); trace[gc * 8 + 6] = ti.p * 100000 + ti.tm * 10 + ti.tn; }
          if (++c2 == n2) { c2 = 0; ++c1; }
          ph ^= 1u << s;
          if (++s == (uint32_t)nstages) s = 0;
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ============================================================================================ MMA issuer
    {
      uint32_t it = 0, s = 0, ph = 0, gcm = 0;
      int slot_bytes = 0, nstages = 1;
      Walker w;
      bool mnm = false;
      uint32_t idesc[3] = {0, 0, 0}, a_off = 0, b_off = 0, a_ks = 0, b_ks = 0, a_lbo = 0, b_lbo = 0, a_ps = 0, ncol = 0;
      int ksteps = 0, nprod = 1;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        if (w.advance(probs, nprob, tile)) {
          const CgProblem& P = probs[w.p];
          mnm = P.mn_major != 0;
          ncol = (uint32_t)P.umma_n;
#pragma unroll
          for (int j = 0; j < 3; ++j) idesc[j] = make_idesc(128, P.umma_n * (j + 1), mnm);      // N = n, 2n, 3n
          ksteps = P.ksteps; nprod = P.nprod; a_ps = (uint32_t)P.a_pstride;
          a_off = P.a_off; b_off = P.b_off; a_ks = P.a_kstep; b_ks = P.b_kstep; a_lbo = P.a_lbo; b_lbo = P.b_lbo;
          if (P.slot_bytes != slot_bytes || P.nstages != nstages) { slot_bytes = P.slot_bytes; nstages = P.nstages; s = 0; }
        }
        const Tile ti = w.tile(tile);
        if (ti.c_end <= ti.c_begin) continue;
        const uint32_t buf = it & 1;
        if (it >= 2) mbar_wait(smem_u32(&bar_acc_empty[buf]), ((it >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t acc = tmem + buf * 256u;
        for (int c = ti.c_begin; c < ti.c_end; ++c) {
          const bool tr = trace && blockIdx.x == trace_cta && gcm < 64 && lane == 0;
          if (tr) trace[gcm * 8 + 3] = clock64();
          mbar_wait(smem_u32(&bar_full[s]), (ph >> s) & 1u);
          if (tr) trace[gcm * 8 + 4] = clock64();
          tc_fence_after();
          const uint32_t sbase = ring + s * (uint32_t)slot_bytes;
          const bool leader = elect_one();
          if (!dbg_nomma) {
            for (int k = 0; k < ksteps; ++k) {
// `planes` wide MMAs per k-step: A_p x [B_0 | .. | B_(planes-1-p)] -> accumulator columns [p*n, planes*n).
              // Column group g therefore collects the products of order 2^(-8g): g0 = hi*hi, g1 = hi*mid + mid*hi,
              // g2 = hi*lo + mid*mid + lo*hi.  The tensor core truncates the fp32 accumulator at every MMA (measured:
              // tools/tc_accum_probe.py, ~2^-25 relative per accumulation), so keeping each order in its own columns keeps that
              // truncation 2^-8g smaller on the correction terms; the epilogue adds the groups small-to-large in fp32.
              // One wide MMA reads the A plane from shared memory once for up to three products -- at N = 64 the SS-mode
              // MMA is shared-memory bound (55 cycles where the tensor floor is 32: profiles/cg_trace_r2.txt).
              const uint32_t pb = sbase + b_off + (uint32_t)k * b_ks;
              const uint64_t db = mnm ? desc_mn(pb, b_lbo) : desc_k(pb);
              uint64_t da[3];
#pragma unroll
              for (int pl = 0; pl < 3; ++pl) {
                const uint32_t pa = sbase + (uint32_t)pl * a_ps + a_off + (uint32_t)k * a_ks;
                da[pl] = mnm ? desc_mn(pa, a_lbo) : desc_k(pa);
              }
              const uint32_t first = (c == ti.c_begin && k == 0) ? 0u : 1u;
              if (leader) {
                if (nprod >= 6) {
                  umma(acc, da[0], db, idesc[2], first); umma(acc + ncol, da[1], db, idesc[1], 1u); umma(acc + 2u * ncol, da[2], db, idesc[0], 1u);
                } else if (nprod >= 3) {
                  umma(acc, da[0], db, idesc[1], first); umma(acc + ncol, da[1], db, idesc[0], 1u);
                } else {
                  umma(acc, da[0], db, idesc[0], first);
                }
              }
            }
          }
          if (leader) umma_commit(smem_u32(&bar_empty[s]));
          __syncwarp();
          if (tr) trace[gcm * 8 + 5] = clock64();
          ++gcm;
          ph ^= 1u << s;
          if (++s == (uint32_t)nstages) s = 0;
        }
        if (elect_one()) umma_commit(smem_u32(&bar_acc_full[buf]));
        __syncwarp();
        ++it;
      }
    }
    __syncwarp();
  } else {
    // ============================================================================================ epilogue
    const int ew = warp - EPI_WARP0, q = warp & 3, half = ew >> 2;
    uint32_t it = 0;
    Walker w;
    const int r = q * 32 + lane;                         // accumulator row of this thread
    // per-problem constants of this thread, recomputed only when the CTA moves to another problem: the row's offset
    // inside a tile (the r -> (i0, i1, i2) decomposition needs integer divisions) and every descriptor field the tile loop reads
    int epi = 0, rows_tile = 0, lim_rows = 0, umma_n = 0, out_planes = 0, grp_stride = 32, n_valid = 0;
    long long roff = 0, rmoff = 0, o_tm = 0, m_tm = 0;
    int ri0 = 0, ri1 = 0, grp_tab = 0;
    int ngrp_acc = 1;
    const float* __restrict__ bias = nullptr;
    const uint16_t* __restrict__ mask = nullptr;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      if (w.advance(probs, nprob, tile)) {
        const CgProblem& Q = probs[w.p];
        epi = Q.epi; rows_tile = Q.rows_tile; lim_rows = Q.lim_rows; umma_n = Q.umma_n; out_planes = Q.out_planes;
        grp_stride = Q.grp_stride; n_valid = Q.n_valid; o_tm = Q.o_tm; m_tm = Q.m_tm; bias = Q.bias; mask = Q.mask;
        const int d0 = Q.d0, d1 = Q.d1;
        const int i0 = r % d0, i12 = r / d0, i1 = i12 % d1, i2 = i12 / d1;
        roff = Q.o_base + (long long)i0 * Q.o0 + (long long)i1 * Q.o1 + (long long)i2 * Q.o2;
        rmoff = Q.m_base + (long long)i0 * Q.m0 + (long long)i1 * Q.m1 + (long long)i2 * Q.m2;
        ri0 = i0; ri1 = i1; grp_tab = Q.grp_tab; ngrp_acc = Q.nprod >= 6 ? 3 : (Q.nprod >= 3 ? 2 : 1);
      }
      const Tile ti = w.tile(tile);
      if (ti.c_end <= ti.c_begin) continue;
      const CgProblem& P = probs[ti.p];
      const uint32_t buf = it & 1;
      // epi_tiles: the two warp quads alternate TILES (quad h owns accumulator buffer h) instead of splitting the column groups of
      // one tile, so that one quad's TMEM / ALU phase overlaps the other's store phase
      if (epi_tiles && (int)buf != half) { ++it; continue; }
      const bool valid0 = r < rows_tile && ti.tm * rows_tile + r < lim_rows;
      const long long off0 = roff + (long long)ti.tm * o_tm, moff0 = rmoff + (long long)ti.tm * m_tm;
      const int n0 = ti.tn * umma_n, ngroups = umma_n >> 5;
      const bool tre = trace && blockIdx.x == trace_cta && (ew & 3) == 0 && lane == 0 && it < 16 && (epi_tiles || ew == 0);
      if (tre) trace[512 + it * 4 + 0] = clock64();
      mbar_wait(smem_u32(&bar_acc_full[buf]), (it >> 1) & 1);
      if (tre) trace[512 + it * 4 + 1] = clock64();
      tc_fence_after();
      const bool split_fin = P.ws != nullptr;          // split-K tile of an ACT problem: partial sums first, the last arriver finishes
      float* __restrict__ wsrow = split_fin ? P.ws + ((size_t)(ti.tm * w.tiles_n + ti.tn) * 128 + r) * umma_n : nullptr;
      bool last_arriver = !split_fin;
      for (int pass = 0; pass < (split_fin ? 2 : 1); ++pass) {
      if (pass == 1) {
        __syncwarp();
        int old = 0;
        if (lane == 0) { __threadfence(); old = atomicAdd(P.ws_cnt + (ti.tm * w.tiles_n + ti.tn) * NEPI_WARPS + ew, 1); }
        old = __shfl_sync(0xffffffffu, old, 0);
        last_arriver = old == w.splits - 1;
        if (!last_arriver) break;
        __threadfence();
        if (lane == 0) P.ws_cnt[(ti.tm * w.tiles_n + ti.tn) * NEPI_WARPS + ew] = 0;     // next step starts from zero
      }
      for (int g = epi_tiles ? 0 : half; g < ngroups; g += epi_tiles ? 1 : 2) {
        const int ng = n0 + 32 * g;                      // first problem column of the group
        // ---- everything that does not depend on the accumulator is fetched BEFORE the TMEM loads are waited for: the epilogue of
        //      a tile is a dependent chain of long-latency operations on 8 warps, so latency, not bandwidth, sets its length
        bool valid = valid0;
        long long off = off0, moff = moff0;
        if (grp_tab) {
          const int gg = ng >> 5;
          valid = valid0 && ri0 < P.grp_lim0[gg] && ri1 < P.grp_lim1[gg];
          off = off0 + P.grp_off[gg]; moff = moff0 + P.grp_moff[gg] - ng;     // (the mask load below adds ng)
        }
        const bool live = !dbg_nostore && ng < n_valid;
        float4 bv[8];
        uint4 mk[4];
        if (epi == CG_EPI_ACT && live) {
#pragma unroll
          for (int j = 0; j < 8; ++j) bv[j] = __ldg(reinterpret_cast<const float4*>(bias + (long long)(ng >> 5) * P.bias_grp) + j);
        }
        if (epi == CG_EPI_DGRAD && live) {
#pragma unroll
          for (int j = 0; j < 4; ++j) mk[j] = valid ? __ldg(reinterpret_cast<const uint4*>(mask + moff + ng) + j) : make_uint4(0, 0, 0, 0);
        }
        float x[32];
        if (pass == 0) {
        uint32_t v[32], u[32], t[32];
        const uint32_t taddr = tmem + buf * 256u + ((uint32_t)(q * 32) << 16) + (uint32_t)(32 * g);
#define CG_TMEM_LD32(dst, addr)                                                                                                              \
  asm volatile(                                                                                                                             \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                                             \
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"     \
      : "=r"(dst[0]), "=r"(dst[1]), "=r"(dst[2]), "=r"(dst[3]), "=r"(dst[4]), "=r"(dst[5]), "=r"(dst[6]), "=r"(dst[7]), "=r"(dst[8]),       \
        "=r"(dst[9]), "=r"(dst[10]), "=r"(dst[11]), "=r"(dst[12]), "=r"(dst[13]), "=r"(dst[14]), "=r"(dst[15]), "=r"(dst[16]),              \
        "=r"(dst[17]), "=r"(dst[18]), "=r"(dst[19]), "=r"(dst[20]), "=r"(dst[21]), "=r"(dst[22]), "=r"(dst[23]), "=r"(dst[24]),             \
        "=r"(dst[25]), "=r"(dst[26]), "=r"(dst[27]), "=r"(dst[28]), "=r"(dst[29]), "=r"(dst[30]), "=r"(dst[31])                             \
      : "r"(addr))
        CG_TMEM_LD32(v, taddr);
        if (ngrp_acc > 1) CG_TMEM_LD32(u, taddr + (uint32_t)umma_n);
        if (ngrp_acc > 2) CG_TMEM_LD32(t, taddr + 2u * (uint32_t)umma_n);
#undef CG_TMEM_LD32
        if (tre && g == 0) trace[576 + it * 4 + 0] = clock64();
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (tre && g == 0) trace[576 + it * 4 + 1] = clock64();
        if (ngrp_acc > 2) {                                // correction column groups, smallest order first
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]) + (__uint_as_float(u[j]) + __uint_as_float(t[j]));
        } else if (ngrp_acc > 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]) + __uint_as_float(u[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
        }
        if (split_fin) {                                   // partial sums of this split -> workspace
#pragma unroll
          for (int j = 0; j < 8; ++j)
            asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(wsrow + 32 * g + 4 * j), "f"(x[4 * j]), "f"(x[4 * j + 1]), "f"(x[4 * j + 2]),
                         "f"(x[4 * j + 3])
                         : "memory");
          continue;
        }
        } else {                                           // last arriver: the complete sums, and a clean workspace for the next step
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 sum = __ldcg(reinterpret_cast<const float4*>(wsrow + 32 * g) + j);
            x[4 * j] = sum.x; x[4 * j + 1] = sum.y; x[4 * j + 2] = sum.z; x[4 * j + 3] = sum.w;
            __stcg(reinterpret_cast<float4*>(wsrow + 32 * g) + j, make_float4(0.f, 0.f, 0.f, 0.f));
          }
        }
        if (!live) continue;
        // Every lane owns one output row and writes its 32 columns itself (64 B per BF16 plane, 128 B of fp32): 16-byte vector
        // stores, no shared-memory transpose -- partial-sector writes are cheap next to the staging round trips they replace.
        if (epi == CG_EPI_RAW) {
          if (valid) {
            float* dst = P.out_f + off + (long long)(ng >> 5) * P.f_grp;
            if (P.atomic) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + 4 * j), "f"(x[4 * j]), "f"(x[4 * j + 1]), "f"(x[4 * j + 2]),
                             "f"(x[4 * j + 3])
                             : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
            }
          }
          continue;
        }
        if (epi == CG_EPI_WGRAD) {
          if (valid) {
            const float sc = P.scale;
            float* dst = P.out_f + off + (long long)(ng >> 5) * P.f_grp;
            if (P.atomic) {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst + 4 * j), "f"(x[4 * j] * sc), "f"(x[4 * j + 1] * sc),
                             "f"(x[4 * j + 2] * sc), "f"(x[4 * j + 3] * sc)
                             : "memory");
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(x[4 * j] * sc, x[4 * j + 1] * sc, x[4 * j + 2] * sc, x[4 * j + 3] * sc);
              if (P.dp_n > 1) {                          // final values: push them to their owner now (posted NVLink stores)
                const int i4 = (int)((dst - P.dp_gbase) >> 2), per4 = P.dp_per4;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const int owner = (i4 + j) / per4;
                  if (owner != P.dp_rank)
                    reinterpret_cast<float4*>(P.dp_recv[owner])[(size_t)P.dp_rank * per4 + (i4 + j - owner * per4)] =
                        make_float4(x[4 * j] * sc, x[4 * j + 1] * sc, x[4 * j + 2] * sc, x[4 * j + 3] * sc);
                }
              }
            }
          }
          continue;
        }
        if (epi == CG_EPI_ACT) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            x[4 * j] = fmaxf(x[4 * j] + bv[j].x, 0.f); x[4 * j + 1] = fmaxf(x[4 * j + 1] + bv[j].y, 0.f);
            x[4 * j + 2] = fmaxf(x[4 * j + 2] + bv[j].z, 0.f); x[4 * j + 3] = fmaxf(x[4 * j + 3] + bv[j].w, 0.f);
          }
          if (P.f0 > 0 && valid && !dbg_nost) {        // optional fp32 copy (cnn_fc1 features for the head kernels)
            float* dst = P.out_f + (long long)ti.tm * P.f_tm + (long long)r * P.f0 + (long long)(ng >> 5) * P.f_grp;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + 8 * j), "f"(x[8 * j]), "f"(x[8 * j + 1]), "f"(x[8 * j + 2]),
                           "f"(x[8 * j + 3]), "f"(x[8 * j + 4]), "f"(x[8 * j + 5]), "f"(x[8 * j + 6]), "f"(x[8 * j + 7])
                           : "memory");
          }
        } else {   // CG_EPI_DGRAD: ReLU mask = hi plane of the forward activation at the same position
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t w[4] = {mk[j].x, mk[j].y, mk[j].z, mk[j].w};
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
              if (!(w[uu] & 0x00007FFFu)) x[8 * j + 2 * uu] = 0.f;
              if (!(w[uu] & 0x7FFF0000u)) x[8 * j + 2 * uu + 1] = 0.f;
            }
          }
          if (P.colsum) {
            // bias gradient: column sums over this warp's 32 rows (invalid rows are zero: their mask words were not loaded) by a
            // transposing butterfly -- 31 shuffles leave lane c with the sum of column c -- then one 128-byte red.add per warp
            float cs[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) cs[j] = x[j];
#pragma unroll
            for (int wd = 16; wd >= 1; wd >>= 1) {
              const bool up = (lane & wd) != 0;
#pragma unroll
              for (int j = 0; j < wd; ++j) {
                const float send = up ? cs[j] : cs[j + wd], keep = up ? cs[j + wd] : cs[j];
                cs[j] = keep + __shfl_xor_sync(0xffffffffu, send, wd);
              }
            }
            atomicAdd(P.colsum + ((ng + lane) & P.colsum_mask), cs[0]);
          }
        }
        if (tre && g == 0) trace[576 + it * 4 + 2] = clock64();
        // ---- BF16 planes: hi = bf16(x), then the residual feeds the next plane
        const long long gcol = grp_tab ? 0 : (long long)(ng >> 5) * grp_stride;
#pragma unroll 1
        for (int pl = 0; pl < out_planes; ++pl) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            pk[j] = pack_bf16x2(x[2 * j], x[2 * j + 1]);
            x[2 * j] -= __uint_as_float(pk[j] << 16);
            x[2 * j + 1] -= __uint_as_float(pk[j] & 0xFFFF0000u);
          }
          if (valid && !dbg_nost) {              // 2 x 256-bit stores: every lane writes whole 32-byte sectors
            uint16_t* dst = P.out_p[pl] + off + gcol;
#pragma unroll
            for (int j = 0; j < 2; ++j)
              asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + 16 * j), "r"(pk[8 * j]), "r"(pk[8 * j + 1]), "r"(pk[8 * j + 2]),
                           "r"(pk[8 * j + 3]), "r"(pk[8 * j + 4]), "r"(pk[8 * j + 5]), "r"(pk[8 * j + 6]), "r"(pk[8 * j + 7])
                           : "memory");
          }
        }
      }
      }   // pass
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&bar_acc_empty[buf]));
        if (P.done_ctr && last_arriver) {                                // this warp's rows of the tile are in global memory
          __threadfence();
          atomicAdd(P.done_ctr + ti.tm, 1);
        }
      }
      if (tre) { trace[512 + it * 4 + 2] = clock64(); trace[512 + it * 4 + 3] = ti.p * 100000 + ti.tm * 10 + ti.tn; }
      ++it;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

}  // namespace

int cg_smem_limit() { return 226 * 1024; }   // dynamic part; the kernel's static shared memory (barriers) takes < 1 KiB of the 227 KiB

int cg_encode_map(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                  const uint32_t* elem_strides) {
  if (!g_encode) {
    cudaDriverEntryPointQueryResult q;
    void* fn = nullptr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn || q != cudaDriverEntryPointSuccess) return -1;
    g_encode = (EncodeTiledFn)fn;
  }
  cuuint64_t gd[5], gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = elem_strides ? elem_strides[i] : 1; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  const CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}

int cg_finalize(CgGroup& g, int smem_budget) {
  int start = 0, slot = 0;
  g.flops = 0;
  for (int i = 0; i < g.n; ++i) {
    CgProblem& P = g.host[i];
    P.tile_start = start;
    start += P.tiles_m * P.tiles_n * P.splits;
    for (int l = 0; l < P.nloads; ++l) {
      CgLoad& L = P.ld[l];
      L.plane_stride = L.smem_off >= P.b_off ? P.b_pstride : P.a_pstride;
      if (L.plane_box && L.box_bytes != L.plane_stride) return -2;      // stacked planes must land where the MMA descriptors look
    }
    const int need = P.b_off + P.planes * P.b_pstride;
    P.slot_bytes = (need + 1023) / 1024 * 1024;
    slot = slot > P.slot_bytes ? slot : P.slot_bytes;
  }
  g.total_tiles = start;
  const int avail = smem_budget - 1024 /* alignment slack */;
  g.slot_bytes = slot; g.nstages = 0;              // (largest slot; the most stages any problem uses)
  int ring = 0;
  for (int i = 0; i < g.n; ++i) {
    CgProblem& P = g.host[i];
    P.nstages = P.slot_bytes > 0 ? avail / P.slot_bytes : 0;
    if (P.nstages > CG_MAX_STAGES) P.nstages = CG_MAX_STAGES;
    if (P.nstages < 2) return -1;
    P.slot_bytes = avail / P.nstages / 1024 * 1024;        // one slot size per stage count: problems with equal depth share the partition
    g.nstages = g.nstages > P.nstages ? g.nstages : P.nstages;
    ring = ring > P.nstages * P.slot_bytes ? ring : P.nstages * P.slot_bytes;
  }
  g.ring_bytes = ring;
  return 0;
}

long long* g_cg_trace = nullptr;      // bring-up: clock64 stamps of CTA 0 (set by tools through b2g_debug_cg_trace)

cudaError_t cg_launch(const CgGroup& g, const CUtensorMap* dev_maps, int num_sms, cudaStream_t s, bool pdl, int debug_flags) {
  if (g.total_tiles <= 0 || (debug_flags & 8)) return cudaSuccess;     // bit 3: skip the launch (timing ablation)
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(cg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cg_smem_limit());
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  static int epi_tiles = -1;
  if (epi_tiles < 0) { const char* e = getenv("B2G_EPI_TILES"); epi_tiles = (e && e[0] == '1') ? 1 : 0; }
  const int smem = 1024 + g.ring_bytes;
  const int grid = g.total_tiles < num_sms ? g.total_tiles : num_sms;
  CgPack pk;              // (host staging; the launch copies it into the parameter buffer)
  static_assert(sizeof(CgPack) < 28 * 1024, "problem list must fit the kernel parameter space (32,764 B on sm_100)");
  for (int i = 0; i < g.n; ++i) pk.p[i] = g.host[i];
  cudaError_t e = launch_pdl(cg_kernel, dim3(grid), dim3(NTHREADS), (size_t)smem, s, pdl, pk, g.n, g.total_tiles, dev_maps, g.nstages, debug_flags, g_cg_trace, epi_tiles);
  if (e != cudaSuccess) fprintf(stderr, "cg_launch %s: %s (grid %d, %d threads, smem %d)\n", g.name, cudaGetErrorString(e), grid, NTHREADS, smem);
  return e;
}

}  // namespace b2g
