// tcgen05 gather-GEMM engine (sm_100a): the dense contractions of the SAC step on the 5th-gen
// tensor cores with fp32 accumulation in TMEM.
//
//   C[cM[m] + cN[n]] (=|+=) epi( sum_r A[aM[m] + aR[r]] * B[bR[r] + bN[n]] )       (common.cuh)
//
// One CTA computes one 128(m) x <=64(n) output tile over its r-range:
//   * 16 producer warps gather fp32 operands through the offset tables (implicit im2col / wgrad /
//     dgrad views), split every value into BF16 hi + BF16 lo (x = hi + lo to ~2^-17) and write
//     both as K-major, 128B-swizzled UMMA tiles into a 3-stage shared-memory ring
//     (generic-proxy stores -> fence.proxy.async -> mbarrier arrive);
//   * 1 MMA warp (one elected thread) issues tcgen05.mma.cta_group::1.kind::f16 over the ring:
//     mode BF16X3: hi*hi + hi*lo + lo*hi (fp32-faithful to ~1e-5 relative, parity mode),
//     mode BF16  : hi*hi only (fast mode); tcgen05.commit frees ring slots and signals the epilogue;
//   * the producer warps then become the epilogue: tcgen05.ld the accumulator rows from TMEM,
//     apply bias+ReLU / ReLU-mask / split-R atomics and store.
// Accumulator: 128 lanes x 64 fp32 columns of TMEM.  Ring: 3 x 48 KiB = 144 KiB of shared memory.
#include <cuda_bf16.h>

#include "common.cuh"

namespace b2g {
namespace {

constexpr int TM = GG_TC_BM, TN = GG_TC_BN, TK = 64;   // tile: 128 x 64 x 64
constexpr int STAGES = 3;
constexpr int NPROD = 512;                              // producer / epilogue threads (16 warps)
constexpr int NTHREADS = NPROD + 32;                    // + MMA warp
constexpr int A_BYTES = TM * TK * 2;                    // 16 KiB per (hi | lo)
constexpr int B_BYTES = TN * TK * 2;                    // 8 KiB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;  // 48 KiB
constexpr int RTAB = 1024;                              // r-offset table entries staged in smem per operand
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 in
// [0,14), LBO (ignored for swizzled K-major) in [16,30), SBO = 1024 B between 8-row groups in
// [32,46), version 1 in [46,48), layout type 2 (SWIZZLE_128B) in [61,64).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: D=f32 (bits 4-5 = 1), A=B=BF16 (bits 7-9, 10-12 = 1), K-major A/B,
// N>>3 at bit 17, M>>4 at bit 24.
__device__ __forceinline__ uint32_t umma_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
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

// 8 fp32 -> 8 bf16 hi (4 x b32) + 8 bf16 lo.  cvt.rn.bf16x2.f32 d, a, b packs a into the upper half.
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(x[2 * i + 1]), "f"(x[2 * i]));
    const float r0 = x[2 * i] - __uint_as_float(d << 16);
    const float r1 = x[2 * i + 1] - __uint_as_float(d & 0xFFFF0000u);
    uint32_t e;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(e) : "f"(r1), "f"(r0));
    h[i] = d; l[i] = e;
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// byte offset of the 16-byte chunk (row, c) inside a K-major SWIZZLE_128B tile (128 B per row)
__device__ __forceinline__ uint32_t sw128(int row, int c) { return (uint32_t)(row * 128 + ((c ^ (row & 7)) << 4)); }

__device__ __forceinline__ void st_shared16(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ float4 ldg4(const float* p) { return *reinterpret_cast<const float4*>(p); }

template <bool a_rvec, bool b_rvec>
__global__ void __launch_bounds__(NTHREADS, 1) gg_tc_kernel(const GemmDesc* __restrict__ descs, int ndesc, int x3) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ GemmDesc sd;
  __shared__ __align__(8) uint64_t bar_full[STAGES], bar_empty[STAGES], bar_accum;
  __shared__ uint32_t tmem_slot;
  __shared__ float cs[TN];
  __shared__ int s_aR[RTAB], s_bR[RTAB];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    int p = 0;
    const int t = blockIdx.x;
    while (p + 1 < ndesc && t >= descs[p + 1].tile_start) ++p;
    sd = descs[p];
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(smem_u32(&bar_full[s]), NPROD);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_accum), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < TN) cs[tid] = 0.f;
  if (warp == NPROD / 32) {   // MMA warp owns the TMEM allocation
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const GemmDesc& d = sd;

  int t = blockIdx.x - d.tile_start;
  const int per = d.tiles_m * d.tiles_n;
  const int split = t / per;
  t -= split * per;
  const int tm = t / d.tiles_n, tn = t - tm * d.tiles_n;
  const int m0 = tm * TM, n0 = tn * TN;
  const int chunk_r = (((d.R + d.splitR - 1) / d.splitR) + TK - 1) / TK * TK;
  const int r_begin = split * chunk_r;
  const int r_end = min(d.R, r_begin + chunk_r);
  const int nchunks = r_end > r_begin ? (r_end - r_begin + TK - 1) / TK : 0;
  const int un = min(TN, ((d.N - n0) + 15) / 16 * 16);   // UMMA N for this tile (multiple of 16)

  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;

  // r-offset tables of this CTA's r-range -> shared memory (falls back to global reads if too long)
  const bool tab_smem = (r_end - r_begin) <= RTAB;
  if (tab_smem) {
    for (int i = tid; i < r_end - r_begin; i += NTHREADS) { s_aR[i] = d.aR[r_begin + i]; s_bR[i] = d.bR[r_begin + i]; }
  }
  __syncthreads();

  if (warp < NPROD / 32) {
    // =========================================================================== producers
    const bool do_colsum = (d.flags & GG_COLSUM) && tm == 0 && !b_rvec;
    const float* __restrict__ A = d.A;
    const float* __restrict__ Bp = d.B;
    const int c8 = tid & 7;             // 16-byte chunk (8 r values) inside the 64-wide r-chunk
    const int q = tid >> 3;             // 0..63
    // A, r-contiguous: all threads, rows q + 64 i (i < 2).  A, m-contiguous: threads < 256, rows 4 q .. 4 q + 3.
    const bool a_thread = a_rvec ? true : (q < 32);
    int a_off[4];
    bool a_ok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + (a_rvec ? q + 64 * i : 4 * q + i);
      a_ok[i] = a_thread && m < d.M && (a_rvec ? i < 2 : true);
      a_off[i] = a_ok[i] ? d.aM[m] : 0;
    }
    // B, n-contiguous: threads 256..383 (q - 32 in 0..15), n = 4 (q-32) .. +3.  B, r-contiguous: all threads, n = q.
    const int qb = b_rvec ? q : q - 32;
    const bool b_thread = b_rvec ? true : (qb >= 0 && qb < TN / 4);
    int b_off[4];
    bool b_ok[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + (b_rvec ? qb : 4 * qb + i);
      b_ok[i] = b_thread && n < d.N && (b_rvec ? i < 1 : true);
      b_off[i] = b_ok[i] ? d.bN[n] : 0;
    }
    float4 csum = make_float4(0, 0, 0, 0);

    for (int ch = 0; ch < nchunks; ++ch) {
      const int s = ch % STAGES;
      const uint32_t sA_hi = ring + s * STAGE_BYTES, sA_lo = sA_hi + A_BYTES;
      const uint32_t sB_hi = sA_lo + A_BYTES, sB_lo = sB_hi + B_BYTES;
      const int rl = ch * TK + c8 * 8;                 // r relative to r_begin
      const int r0 = r_begin + rl;
      const int nr = min(8, r_end - r0);               // valid r values (<= 0: none)
      int ar[8], br[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ar[j] = br[j] = 0;
        if (j < nr) {
          ar[j] = tab_smem ? s_aR[rl + j] : d.aR[r0 + j];
          br[j] = tab_smem ? s_bR[rl + j] : d.bR[r0 + j];
        }
      }
      // ---------------------------------------------------------------- issue every global load first
      float xa[4][8];     // rvec: [unit i][r j] (i < 2);  mvec: [m i][r j]
      float xb[4][8];     // rvec: [0][r j];                nvec: [n i][r j]
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) xa[i][j] = xb[i][j] = 0.f;
      if (a_rvec) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (a_ok[i] && nr > 0) {
            if (nr == 8) {
              const float4 v = ldg4(A + a_off[i] + ar[0]), w = ldg4(A + a_off[i] + ar[4]);
              xa[i][0] = v.x; xa[i][1] = v.y; xa[i][2] = v.z; xa[i][3] = v.w;
              xa[i][4] = w.x; xa[i][5] = w.y; xa[i][6] = w.z; xa[i][7] = w.w;
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (j < nr) xa[i][j] = A[a_off[i] + ar[j]];
            }
          }
        }
      } else if (a_thread) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < nr) {
            if (a_ok[3]) {
              const float4 v = ldg4(A + a_off[0] + ar[j]);
              xa[0][j] = v.x; xa[1][j] = v.y; xa[2][j] = v.z; xa[3][j] = v.w;
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (a_ok[i]) xa[i][j] = A[a_off[i] + ar[j]];
            }
          }
        }
      }
      if (b_rvec) {
        if (b_ok[0] && nr > 0) {
          if (nr == 8) {
            const float4 v = ldg4(Bp + b_off[0] + br[0]), w = ldg4(Bp + b_off[0] + br[4]);
            xb[0][0] = v.x; xb[0][1] = v.y; xb[0][2] = v.z; xb[0][3] = v.w;
            xb[0][4] = w.x; xb[0][5] = w.y; xb[0][6] = w.z; xb[0][7] = w.w;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j < nr) xb[0][j] = Bp[b_off[0] + br[j]];
          }
        }
      } else if (b_thread) {
        // wgrad mode (A m-contiguous): A-threads and B-threads are disjoint, so B reuses the xa registers
        float (&xq)[4][8] = a_rvec ? xb : xa;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < nr) {
            if (b_ok[3]) {
              const float4 v = ldg4(Bp + b_off[0] + br[j]);
              xq[0][j] = v.x; xq[1][j] = v.y; xq[2][j] = v.z; xq[3][j] = v.w;
            } else {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (b_ok[i]) xq[i][j] = Bp[b_off[i] + br[j]];
            }
          }
        }
        if (do_colsum) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { csum.x += xq[0][j]; csum.y += xq[1][j]; csum.z += xq[2][j]; csum.w += xq[3][j]; }
        }
      }
      // ---------------------------------------------------------------- ring slot free?  then split + store
      if (ch >= STAGES) mbar_wait(smem_u32(&bar_empty[s]), ((ch / STAGES) - 1) & 1);
      if (a_thread) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (a_rvec && i >= 2) break;
          uint4 hi, lo;
          split8(xa[i], hi, lo);
          const uint32_t o = sw128(a_rvec ? q + 64 * i : 4 * q + i, c8);
          st_shared16(sA_hi + o, hi);
          if (x3) st_shared16(sA_lo + o, lo);
        }
      }
      if (b_thread) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (b_rvec && i >= 1) break;
          uint4 hi, lo;
          split8((a_rvec || b_rvec) ? xb[i] : xa[i], hi, lo);
          const uint32_t o = sw128(b_rvec ? qb : 4 * qb + i, c8);
          st_shared16(sB_hi + o, hi);
          if (x3) st_shared16(sB_lo + o, lo);
        }
      }
      fence_proxy_async();                 // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      mbar_arrive(smem_u32(&bar_full[s]));
    }

    // =========================================================================== epilogue
    if (nchunks > 0) {
      mbar_wait(smem_u32(&bar_accum), 0);
      tc_fence_after();
    }
    const int lq = warp & 3;                     // TMEM lane quarter this warp may access
    const int ch0 = (warp >> 2) * (TN / 4);      // column quarter
    const int m = m0 + lq * 32 + lane;
    const bool m_ok = m < d.M;
    const int cm = m_ok ? d.cM[m] : 0;
    const int km = (m_ok && (d.flags & GG_EPI_MASK)) ? (d.kM ? d.kM[m] : cm) : 0;
#pragma unroll
    for (int cb = 0; cb < TN / 4; cb += 8) {
      uint32_t v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (nchunks > 0 && ch0 + cb < un) {       // warp-uniform
        const uint32_t taddr = tmem + ((uint32_t)(lq * 32) << 16) + (uint32_t)(ch0 + cb);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      }
      if (!m_ok) continue;
      float o[8];
      int co[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = n0 + ch0 + cb + j;
        o[j] = __uint_as_float(v[j]);
        co[j] = -1;
        if (n < d.N) {
          const int cn = d.cN[n];
          co[j] = cm + cn;
          if (d.flags & GG_EPI_BIAS_RELU) o[j] = fmaxf(o[j] + d.bias[n], 0.f);
          if (d.flags & GG_EPI_MASK) {
            const int kn = d.kN ? d.kN[n] : cn;
            o[j] = d.mask[km + kn] > 0.f ? o[j] : 0.f;
          }
        }
      }
      if (d.flags & GG_EPI_ATOMIC) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (co[j] >= 0) atomicAdd(d.C + co[j], o[j]);
      } else {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int b = 4 * g;
          if (co[b + 3] >= 0 && co[b + 1] == co[b] + 1 && co[b + 2] == co[b] + 2 && co[b + 3] == co[b] + 3 && (co[b] & 3) == 0) {
            *reinterpret_cast<float4*>(d.C + co[b]) = make_float4(o[b], o[b + 1], o[b + 2], o[b + 3]);
          } else {
#pragma unroll
            for (int j = b; j < b + 4; ++j)
              if (co[j] >= 0) d.C[co[j]] = o[j];
          }
        }
      }
    }
    if (do_colsum) {
      if (b_thread) {
        atomicAdd(&cs[4 * qb + 0], csum.x); atomicAdd(&cs[4 * qb + 1], csum.y);
        atomicAdd(&cs[4 * qb + 2], csum.z); atomicAdd(&cs[4 * qb + 3], csum.w);
      }
    }
    tc_fence_before();
    asm volatile("bar.sync 1, %0;" ::"n"(NPROD));           // producers only
    if ((d.flags & GG_COLSUM) && tm == 0 && !b_rvec && tid < TN && n0 + tid < d.N) atomicAdd(d.colsum + n0 + tid, cs[tid]);
  } else {
    // =========================================================================== MMA issuer
    if (lane == 0 && nchunks > 0) {
      const uint32_t idesc = umma_idesc(TM, un);
      for (int ch = 0; ch < nchunks; ++ch) {
        const int s = ch % STAGES;
        mbar_wait(smem_u32(&bar_full[s]), (ch / STAGES) & 1);
        tc_fence_after();
        const uint32_t sA_hi = ring + s * STAGE_BYTES, sA_lo = sA_hi + A_BYTES;
        const uint32_t sB_hi = sA_lo + A_BYTES, sB_lo = sB_hi + B_BYTES;
#pragma unroll
        for (int k = 0; k < TK / 16; ++k) {
          const uint64_t ah = umma_desc(sA_hi + k * 32), bh = umma_desc(sB_hi + k * 32);
          umma_bf16(tmem, ah, bh, idesc, (ch | k) ? 1u : 0u);
          if (x3) {
            const uint64_t al = umma_desc(sA_lo + k * 32), bl = umma_desc(sB_lo + k * 32);
            umma_bf16(tmem, ah, bl, idesc, 1u);
            umma_bf16(tmem, al, bh, idesc, 1u);
          }
        }
        umma_commit(smem_u32(&bar_empty[s]));      // frees the ring slot when these MMAs retire
      }
      umma_commit(smem_u32(&bar_accum));           // accumulator complete -> epilogue
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NPROD / 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TN));
  }
}
}  // namespace

int gg_tc_smem_bytes() { return SMEM_BYTES; }

template <bool AR, bool BR>
static cudaError_t launch_mode(const GemmDesc* dev_descs, int ndesc, int total_tiles, int x3, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gg_tc_kernel<AR, BR>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  gg_tc_kernel<AR, BR><<<total_tiles, NTHREADS, SMEM_BYTES, s>>>(dev_descs, ndesc, x3);
  return cudaSuccess;
}

// All problems of one launch share the operand-contiguity mode (flags & (GG_A_RVEC | GG_B_RVEC)).
cudaError_t gg_tc_launch(const GemmDesc* dev_descs, int ndesc, int total_tiles, int mode_flags, int x3, cudaStream_t s) {
  if (total_tiles <= 0) return cudaSuccess;
  const bool ar = mode_flags & GG_A_RVEC, br = mode_flags & GG_B_RVEC;
  if (ar && br) return launch_mode<true, true>(dev_descs, ndesc, total_tiles, x3, s);
  if (ar) return launch_mode<true, false>(dev_descs, ndesc, total_tiles, x3, s);
  if (br) return launch_mode<false, true>(dev_descs, ndesc, total_tiles, x3, s);
  return launch_mode<false, false>(dev_descs, ndesc, total_tiles, x3, s);
}

}  // namespace b2g
