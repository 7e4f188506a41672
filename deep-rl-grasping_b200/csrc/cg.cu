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
//   warps 3..10 : epilogue.  tcgen05.ld one accumulator row per thread, apply bias/ReLU or the ReLU mask, split into
//                 BF16 planes, transpose through a 1 KiB warp-private staging tile and write coalesced rows.
#include <cuda_bf16.h>

#include <cstdio>

#include "cg.cuh"
#include "common.cuh"

namespace b2g {
namespace {

// warps 0..1: TMA producers, warp 2: MMA issuer (owns TMEM), warps 3..10: epilogue
constexpr int NPROD_WARPS = 2, MMA_WARP = NPROD_WARPS;
constexpr int EPI_WARP0 = NPROD_WARPS + 1, NEPI_WARPS = 8;
constexpr int NTHREADS = 32 * (EPI_WARP0 + NEPI_WARPS);
constexpr int STG_PER_WARP = 1024;
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
cg_kernel(const __grid_constant__ CgPack pk, int nprob, int total_tiles, const CUtensorMap* __restrict__ maps, int slot_bytes, int nstages,
          int dbg, long long* __restrict__ trace) {
  const CgProblem* __restrict__ probs = pk.p;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[CG_MAX_STAGES], bar_empty[CG_MAX_STAGES], bar_acc_full[2], bar_acc_empty[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform for the compiler, not just in fact
  const bool dbg_noload = dbg & 1, dbg_nomma = dbg & 2, dbg_nostore = dbg & 4;
  if (tid == 0) {
    for (int s = 0; s < nstages; ++s) { mbar_init(smem_u32(&bar_full[s]), 1); mbar_init(smem_u32(&bar_empty[s]), 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&bar_acc_full[b]), 1); mbar_init(smem_u32(&bar_acc_empty[b]), NEPI_WARPS); }
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
  const uint32_t stg_base = ring + (uint32_t)nstages * (uint32_t)slot_bytes;
  pdl_trigger();
  pdl_wait();

  if (warp < NPROD_WARPS) {
    // ============================================================================================ TMA producers
    {
      uint32_t gc = 0, s = 0, ph = 0;                    // ring slot and its phase, advanced without divisions
      Walker w;
      int n2 = 1, nloads = 0, planes = 0, tx = 0;
      const int* tab = nullptr;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        if (w.advance(probs, nprob, tile)) {
          const CgProblem& P = probs[w.p];
          n2 = P.n2; nloads = P.nloads; planes = P.planes; tx = P.tx_bytes; tab = P.tm_tab;
        }
        const Tile ti = w.tile(tile);
        const CgProblem& P = probs[ti.p];
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
          const bool tr = trace && blockIdx.x == 0 && warp == 0 && gc < 64 && lane == 0;
          if (tr) trace[gc * 8 + 0] = clock64();
          if (gc >= (uint32_t)nstages) mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1);
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
                  for (int pl = 0; pl < planes; ++pl, ++j)
                    if ((j & (NPROD_WARPS - 1)) == warp && leader)
                      tma_load(sbase + (uint32_t)(pl * L.plane_stride + L.smem_off), maps + L.map + pl, full, L.rank, crd);
                }
              }
            }
          }
          if (tr) trace[gc * 8 + 2] = clock64();
          if (++c2 == n2) { c2 = 0; ++c1; }
          if (++s == (uint32_t)nstages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ============================================================================================ MMA issuer
    {
      uint32_t it = 0, s = 0, ph = 0, gcm = 0;
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
        }
        const Tile ti = w.tile(tile);
        if (ti.c_end <= ti.c_begin) continue;
        const uint32_t buf = it & 1;
        if (it >= 2) mbar_wait(smem_u32(&bar_acc_empty[buf]), ((it >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t acc = tmem + buf * 256u;
        for (int c = ti.c_begin; c < ti.c_end; ++c) {
          const bool tr = trace && blockIdx.x == 0 && gcm < 64 && lane == 0;
          if (tr) trace[gcm * 8 + 3] = clock64();
          mbar_wait(smem_u32(&bar_full[s]), ph);
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
          if (++s == (uint32_t)nstages) { s = 0; ph ^= 1; }
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
    const uint32_t stg = stg_base + (uint32_t)ew * STG_PER_WARP;
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
      const bool valid0 = r < rows_tile && ti.tm * rows_tile + r < lim_rows;
      const long long off0 = roff + (long long)ti.tm * o_tm, moff0 = rmoff + (long long)ti.tm * m_tm;
      const int n0 = ti.tn * umma_n, ngroups = umma_n >> 5;
      const bool tre = trace && blockIdx.x == 0 && ew == 0 && lane == 0 && it < 16;
      if (tre) trace[512 + it * 4 + 0] = clock64();
      mbar_wait(smem_u32(&bar_acc_full[buf]), (it >> 1) & 1);
      if (tre) trace[512 + it * 4 + 1] = clock64();
      tc_fence_after();
      for (int g = half; g < ngroups; g += 2) {
        const int ng = n0 + 32 * g;                      // first problem column of the group
        uint32_t v[32];
        const uint32_t taddr = tmem + buf * 256u + ((uint32_t)(q * 32) << 16) + (uint32_t)(32 * g);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
              "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
              "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
              "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (ngrp_acc > 1) {                                // correction column groups, smallest order first
          uint32_t u[32];
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
              : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]), "=r"(u[10]),
                "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]),
                "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]),
                "=r"(u[31])
              : "r"(taddr + (uint32_t)umma_n));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (ngrp_acc > 2) {
            uint32_t t[32];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]), "=r"(t[8]), "=r"(t[9]), "=r"(t[10]),
                  "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15]), "=r"(t[16]), "=r"(t[17]), "=r"(t[18]), "=r"(t[19]), "=r"(t[20]),
                  "=r"(t[21]), "=r"(t[22]), "=r"(t[23]), "=r"(t[24]), "=r"(t[25]), "=r"(t[26]), "=r"(t[27]), "=r"(t[28]), "=r"(t[29]), "=r"(t[30]),
                  "=r"(t[31])
                : "r"(taddr + 2u * (uint32_t)umma_n));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) u[j] = __float_as_uint(__uint_as_float(u[j]) + __uint_as_float(t[j]));
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(u[j]));
        }
        if (dbg_nostore || ng >= n_valid) continue;
        bool valid = valid0;
        long long off = off0, moff = moff0;
        if (grp_tab) {
          const int gg = ng >> 5;
          valid = valid0 && ri0 < P.grp_lim0[gg] && ri1 < P.grp_lim1[gg];
          off = off0 + P.grp_off[gg]; moff = moff0 + P.grp_moff[gg] - ng;     // (the mask load below adds ng)
        }
        float x[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
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
          const float sc = P.scale;
          float* __restrict__ G = P.out_f;
          const bool atomic = P.atomic != 0;
          // 8 rows x 128 B per pass through the staging tile, then 8 lanes cover one row's 32 floats
#pragma unroll 1
          for (int pass = 0; pass < 4; ++pass) {
            if ((lane >> 3) == pass) {
              const int rr = lane & 7;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const uint32_t a = stg + (uint32_t)(rr * 128 + ((j ^ rr) << 4));
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a), "f"(x[4 * j] * sc), "f"(x[4 * j + 1] * sc), "f"(x[4 * j + 2] * sc),
                             "f"(x[4 * j + 3] * sc)
                             : "memory");
              }
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int idx = lane + 32 * i, rr = idx >> 3, ch = idx & 7, rq = pass * 8 + rr;
              const long long o = __shfl_sync(0xffffffffu, off, rq);
              const bool ok = __shfl_sync(0xffffffffu, (int)valid, rq) != 0;
              float4 w;
              const uint32_t a = stg + (uint32_t)(rr * 128 + ((ch ^ rr) << 4));
              asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(w.x), "=f"(w.y), "=f"(w.z), "=f"(w.w) : "r"(a));
              if (ok) {
                float* dst = G + o + (long long)(ng >> 5) * P.f_grp + 4 * ch;
                if (atomic) asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "f"(w.x), "f"(w.y), "f"(w.z), "f"(w.w) : "memory");
                else *reinterpret_cast<float4*>(dst) = w;
              }
            }
            __syncwarp();
          }
          continue;
        }
        if (epi == CG_EPI_ACT) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + (long long)(ng >> 5) * P.bias_grp) + j);
            x[4 * j] = fmaxf(x[4 * j] + b.x, 0.f); x[4 * j + 1] = fmaxf(x[4 * j + 1] + b.y, 0.f);
            x[4 * j + 2] = fmaxf(x[4 * j + 2] + b.z, 0.f); x[4 * j + 3] = fmaxf(x[4 * j + 3] + b.w, 0.f);
          }
          if (P.f0 > 0 && valid) {        // optional fp32 copy (cnn_fc1 features for the head kernels)
            float* dst = P.out_f + (long long)ti.tm * P.f_tm + (long long)r * P.f0 + (long long)(ng >> 5) * P.f_grp;
#pragma unroll
            for (int j = 0; j < 8; ++j) *reinterpret_cast<float4*>(dst + 4 * j) = make_float4(x[4 * j], x[4 * j + 1], x[4 * j + 2], x[4 * j + 3]);
          }
        } else {   // CG_EPI_DGRAD: ReLU mask = hi plane of the forward activation at the same position
          uint4 mk[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) mk[j] = valid ? __ldg(reinterpret_cast<const uint4*>(mask + moff + ng) + j) : make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t w[4] = {mk[j].x, mk[j].y, mk[j].z, mk[j].w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (!(w[u] & 0x00007FFFu)) x[8 * j + 2 * u] = 0.f;
              if (!(w[u] & 0x7FFF0000u)) x[8 * j + 2 * u + 1] = 0.f;
            }
          }
        }
        // ---- BF16 planes: hi = bf16(x), then the residual feeds the next plane; 16 rows x 64 B per pass through staging
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
          uint16_t* __restrict__ dstp = P.out_p[pl];
#pragma unroll 1
          for (int pass = 0; pass < 2; ++pass) {
            if ((lane >> 4) == pass) {
              const int rr = lane & 15;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t a = stg + (uint32_t)(rr * 64 + ((j ^ ((rr >> 1) & 3)) << 4));
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(pk[4 * j]), "r"(pk[4 * j + 1]), "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                             : "memory");
              }
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int idx = lane + 32 * i, rr = idx >> 2, ch = idx & 3, rq = pass * 16 + rr;
              const long long o = __shfl_sync(0xffffffffu, off, rq);
              const bool ok = __shfl_sync(0xffffffffu, (int)valid, rq) != 0;
              uint4 w;
              const uint32_t a = stg + (uint32_t)(rr * 64 + ((ch ^ ((rr >> 1) & 3)) << 4));
              asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "r"(a));
              if (ok) *reinterpret_cast<uint4*>(dstp + o + gcol + 8 * ch) = w;
            }
            __syncwarp();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bar_acc_empty[buf]));
      if (tre) trace[512 + it * 4 + 2] = clock64();
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
    for (int l = 0; l < P.nloads; ++l) P.ld[l].plane_stride = P.ld[l].smem_off >= P.b_off ? P.b_pstride : P.a_pstride;
    const int need = P.planes * (P.a_pstride + P.b_pstride);
    slot = slot > need ? slot : need;
  }
  g.total_tiles = start;
  g.slot_bytes = (slot + 1023) / 1024 * 1024;
  const int avail = smem_budget - 1024 /* alignment slack */ - NEPI_WARPS * STG_PER_WARP;
  g.nstages = g.slot_bytes > 0 ? avail / g.slot_bytes : 0;
  if (g.nstages > CG_MAX_STAGES) g.nstages = CG_MAX_STAGES;
  return g.nstages >= 2 ? 0 : -1;
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
  const int smem = 1024 + g.nstages * g.slot_bytes + NEPI_WARPS * STG_PER_WARP;
  const int grid = g.total_tiles < num_sms ? g.total_tiles : num_sms;
  CgPack pk;              // (host staging; the launch copies it into the parameter buffer)
  static_assert(sizeof(CgPack) < 16 * 1024, "problem list must fit the kernel parameter space");
  for (int i = 0; i < g.n; ++i) pk.p[i] = g.host[i];
  cudaError_t e = launch_pdl(cg_kernel, dim3(grid), dim3(NTHREADS), (size_t)smem, s, pdl, pk, g.n, g.total_tiles, dev_maps, g.slot_bytes, g.nstages, debug_flags, g_cg_trace);
  if (e != cudaSuccess) fprintf(stderr, "cg_launch %s: %s (grid %d, %d threads, smem %d = %d stages x %d)\n", g.name, cudaGetErrorString(e), grid, NTHREADS, smem, g.nstages, g.slot_bytes);
  return e;
}

}  // namespace b2g
