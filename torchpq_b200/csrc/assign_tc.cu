// MultiKMeans assignment on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
// Replaces the reference's fused SIMT distance-GEMM + arg-max (torchpq/kernels/cuda/max_sim.cu:182-309,
// 128x128 tiles, 8x8 register micro-tiles, float atomicMax epilogue) for the shape the PQ codec uses
// (MultiKMeans.get_labels, clustering/MultiKMeans.py:314-333): data [l, d, n], centroids [l, d, k<=256].
//
//   label_i = argmax_j ( 2 <x_i, c_j> - |c_j|^2 )         (|x_i|^2 does not change the arg-max)
//
// The inner products run as TF32 UMMA (M=128 points x N<=256 centroids x K=8 per instruction), accumulators in
// TMEM.  Operands are read by TMA exactly as they lie in HBM: both matrices are "MN-major" (points / centroids
// contiguous), which tcgen05 accepts for TF32, so a TMA box of 32 columns x d rows with the 128B/32B-atom swizzle
// IS the canonical UMMA shared-memory layout -- no transposition, no conversion pass.
// The value written to `maxsims` is then recomputed EXACTLY (fp32, fmaf(-dif, dif, acc), e ascending, as
// max_sim.cu:78-98) for the chosen centroid from the same shared-memory tiles, so values are bit-identical to the
// exact kernel wherever the label agrees; labels can differ only where two centroids are within TF32 rounding
// (~1e-3 relative) of each other.
//
// Persistent, warp-specialised: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2..9 = arg-max
// over the TMEM accumulator (4 lane quadrants x 2 column halves), warps 10..17 = finishers (exact similarity + stores),
// two groups alternating tiles so the dependent fp32 chain of one tile overlaps the next tile's arg-max.
//
// Measured (profiles/r01_assign_tc_raw.csv, C5): 5.6 ms = 3.0 TB/s of HBM traffic (the stream alone would take
// 2.65 ms), tensor pipe 38 % busy, issue slots 63 % busy, L1/shared pipe 57 % busy of which two thirds are bank
// conflicts of the finishers' centroid-column gathers (random columns of the swizzled B tile) -- the next thing to
// attack (TMEM read-back is only 4 % of its peak, so the arg-max read of the accumulators is not the limit).
#include <cuda.h>
#include "common.cuh"

namespace tpq {

constexpr int TC_M = 128;          // points per tile
constexpr int TC_STAGES = 4;
constexpr int TC_THREADS = 64 + 512;

// ----------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// UMMA shared-memory descriptor, MN-major 32-bit operand.  For MN-major TF32 the only legal shared-memory layout is
// SWIZZLE_128B_BASE32B (cutlass sm100_common.inl:92): 128-byte rows, the four 32-byte chunks of a row XOR-ed with
// (row % 4) -- Swizzle<2,5,2> -- which is what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
// Fields (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start>>4 [0,14) | LBO>>4 [16,30) (stride between 32-element
// MN groups) | SBO>>4 [32,46) (stride between 4-row K groups) | version=1 [46,48) | layout_type=1 [61,64)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(tmem_c), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}

// element (row e, column c) of a [rows x 32-column] TMA box written with the 128B / 32B-atom swizzle
__device__ __forceinline__ uint32_t box_off(int e, int c) {
  return (uint32_t)(e * 128 + ((((c >> 3) ^ (e & 3)) << 5) | ((c & 7) << 2)));
}

constexpr float kNegHuge = -3.0e38f;   // below every score; not -inf (index bits OR-ed into -inf would make a NaN)

#ifdef TPQ_DEBUG_KNOBS
#define TC_WAIT(slot, bar, par) do { const long long t0__ = clock64(); mbar_wait(bar, par); \
    if (P.wait && (threadIdx.x & 31) == 0) atomicAdd(P.wait + (slot), (unsigned long long)(clock64() - t0__)); } while (0)
#else
#define TC_WAIT(slot, bar, par) mbar_wait(bar, par)
#endif

struct TcParams {
  int l, d, n, k, kpad;          // kpad = k rounded up to 32 (TMA box granularity), N of the MMA = kpad
  int tiles_per_l;               // ceil(n / 128)
  float* maxsims; int64_t* labels;
  const float* cent;             // for |c|^2 of padded columns nothing is read: TMA zero-fills
  float* dbg;                    // debug: raw accumulators of tile 0 of CTA 0, [128][256]
  int exact_values;              // 1: maxsims recomputed exactly in fp32 for the chosen centroid; 0: 2*score - |x|^2 from the TF32 score
  unsigned long long* wait;      // debug build: summed wait cycles per barrier kind [8], [7] = total CTA cycles
};
static float* g_tc_dbg = nullptr;
static unsigned long long* g_tc_wait = nullptr;

__global__ void __launch_bounds__(TC_THREADS, 1)
assign_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_c, TcParams P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int d = P.d;
  const uint32_t box_bytes = (uint32_t)d * 128;                       // one 32-column box
  const uint32_t b_bytes = (uint32_t)(P.kpad / 32) * box_bytes;
  const uint32_t a_bytes = 4 * box_bytes;
  uint8_t* sB = smem;                                                  // 1024-aligned (box_bytes is a multiple of 1024: d % 8 == 0)
  uint8_t* sA = smem + b_bytes;
  uint8_t* sAx = sA + TC_STAGES * a_bytes;                             // 4 boxes x 1 KB: the constant "ones" K-block of A
  uint8_t* sBx = sAx + 4 * 1024;                                       // 8 boxes x 1 KB: -|c_j|^2/2 as (hi, lo) rows of B
  float* c2 = reinterpret_cast<float*>(sBx + 8 * 1024);                // [256] scratch
  uint64_t* bars = reinterpret_cast<uint64_t*>(c2 + 256);
  uint64_t* full_a = bars;                    // [STAGES]
  uint64_t* empty_a = bars + TC_STAGES;       // [STAGES]
  uint64_t* tmem_full = bars + 2 * TC_STAGES; // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint64_t* b_full = tmem_empty + 2;          // [1]
  uint64_t* b_free = b_full + 1;              // [1]
  uint64_t* c2_ready = b_free + 1;            // [1]
  uint64_t* cand_full = c2_ready + 1;         // [2]
  uint64_t* cand_empty = cand_full + 2;       // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(cand_empty + 2);
  float* cand_v = reinterpret_cast<float*>(tmem_ptr + 2);              // [2 tile parities][2 column halves][128] partial winners
  int* cand_i = reinterpret_cast<int*>(cand_v + 512);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#ifdef TPQ_DEBUG_KNOBS
  const long long t_cta0 = clock64();
#endif
  const long long total_tiles = (long long)P.l * P.tiles_per_l;
  // Contiguous tile range per CTA (one or two centroid-tile loads per CTA).  A round-robin deal (adjacent strips across
  // the grid, for DRAM page locality) was measured: 4.99 vs 4.70 ms -- the per-k-means reload of the centroid tile costs
  // more than the locality buys.
  const long long t_begin = total_tiles * blockIdx.x / gridDim.x;
  const long long t_end = total_tiles * (blockIdx.x + 1) / gridDim.x;
  const long long t_step = 1;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&map_c) : "memory");
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full_a[s], 1); mbar_init(&empty_a[s], 1 + 4); }
    for (int t = 0; t < 2; ++t) { mbar_init(&tmem_full[t], 1); mbar_init(&tmem_empty[t], 8); mbar_init(&cand_full[t], 8); mbar_init(&cand_empty[t], 4); }
    mbar_init(b_full, 1); mbar_init(b_free, 1 + 8); mbar_init(c2_ready, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // extra K-block operands: A rows 0,1 = 1.0 (rows 2..7 = 0); B rows filled per k-means by the epilogue
  for (int i = threadIdx.x; i < 3 * 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(sAx)[i] = 0u;   // sAx + sBx = 12 KB
  __syncthreads();
  if (threadIdx.x < 128) {
    const int c = threadIdx.x;
    *reinterpret_cast<float*>(sAx + (c >> 5) * 1024 + box_off(0, c & 31)) = 1.f;
    *reinterpret_cast<float*>(sAx + (c >> 5) * 1024 + box_off(1, c & 31)) = 1.f;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int cur_l = -1; uint32_t bgen = 0;
      long long it = 0;
      for (long long t = t_begin; t < t_end; t += t_step, ++it) {
        const int li = (int)(t / P.tiles_per_l), ti = (int)(t % P.tiles_per_l);
        if (li != cur_l) {
          if (cur_l >= 0) { mbar_wait(b_free, bgen & 1); ++bgen; }     // every MMA + epilogue read of the old B is done
          mbar_expect_tx(b_full, b_bytes);
          for (int g = 0; g < P.kpad / 32; ++g) tma_load_3d(sB + g * box_bytes, &map_c, b_full, g * 32, 0, li);
          cur_l = li;
        }
        const int s = (int)(it % TC_STAGES); const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1;
        TC_WAIT(0, &empty_a[s], ph ^ 1);
        mbar_expect_tx(&full_a[s], a_bytes);
        for (int g = 0; g < 4; ++g) tma_load_3d(sA + s * a_bytes + g * box_bytes, &map_x, &full_a[s], ti * TC_M + g * 32, 0, li);
      }
    }
  } else if (warp == 1) {
    // ===================================================== MMA issuer
    // instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): D=f32 [4,6)=1, A=tf32 [7,10)=2,
    // B=tf32 [10,13)=2, A MN-major bit15, B MN-major bit16, N>>3 [17,23), M>>4 [24,29)
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                           ((uint32_t)(P.kpad >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    int cur_l = -1; uint32_t bgen = 0;
    long long it = 0;
    for (long long t = t_begin; t < t_end; t += t_step, ++it) {
      const int li = (int)(t / P.tiles_per_l);
      if (li != cur_l) { mbar_wait(b_full, bgen & 1); mbar_wait(c2_ready, bgen & 1); ++bgen; cur_l = li; }
      const int s = (int)(it % TC_STAGES); const uint32_t ph = (uint32_t)(it / TC_STAGES) & 1;
      const int tb = (int)(it & 1); const uint32_t tph = (uint32_t)(it >> 1) & 1;
      TC_WAIT(1, &tmem_empty[tb], tph ^ 1);
      TC_WAIT(2, &full_a[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
        const uint32_t a0 = smem_u32(sA + s * a_bytes), b0 = smem_u32(sB);
        for (int kb = 0; kb < d / 8; ++kb) {
          const uint64_t da = umma_desc(a0 + kb * 1024, box_bytes, 512);
          const uint64_t db = umma_desc(b0 + kb * 1024, box_bytes, 512);
          umma_tf32(tmem_base + tb * 256, da, db, idesc, kb > 0 ? 1u : 0u);
        }
        // + 1 * hi(-|c|^2/2) + 1 * lo(-|c|^2/2): the accumulator then holds <x,c> - |c|^2/2, whose arg-max is the label
        umma_tf32(tmem_base + tb * 256, umma_desc(smem_u32(sAx), 1024, 512), umma_desc(smem_u32(sBx), 1024, 512), idesc, 1u);
        umma_commit(&empty_a[s]);                                      // MMA done with this A stage
        umma_commit(&tmem_full[tb]);                                   // accumulator ready
        const bool last_of_l = (t + t_step >= t_end) || ((int)((t + t_step) / P.tiles_per_l) != li);
        if (last_of_l) umma_commit(b_free);
      }
      __syncwarp();
    }
  } else if (warp < 10) {
    // ===================================================== arg-max warps: 8 warps = 4 TMEM lane quadrants x 2 column halves.
    // A warp may only touch TMEM lanes [32 (warp % 4), +32).  Thread (quad, lane) owns point `row` and scans the 128
    // accumulator columns of its half; the two partial winners of a row go to shared memory for the finisher warps.
    const int quad = warp & 3;
    const int ch = (warp - 2) >> 2;                                    // column half: columns [128 ch, 128 ch + 128)
    const int row = quad * 32 + lane;
    const int at = threadIdx.x - 64;                                   // 0..255 among arg-max threads = centroid column
    int cur_l = -1; uint32_t bgen = 0;
    long long it = 0;
    for (long long t = t_begin; t < t_end; t += t_step, ++it) {
      const int li = (int)(t / P.tiles_per_l);
      const int tb = (int)(it & 1); const uint32_t tph = (uint32_t)(it >> 1) & 1;
      if (li != cur_l) {
        // -|c_j|^2 / 2 of the new centroid set as (hi, lo) TF32 rows of the extra K-block of B
        mbar_wait(b_full, bgen & 1); ++bgen; cur_l = li;
        const int j = at;
        float v = -1e30f;                                              // padded columns can never win
        if (j < P.k) {
          const uint8_t* bj = sB + (j >> 5) * box_bytes;
          float s2 = 0.f;
          for (int e = 0; e < d; ++e) { const float cv = *reinterpret_cast<const float*>(bj + box_off(e, j & 31)); s2 = fmaf(cv, cv, s2); }
          v = -0.5f * s2;
        }
        const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);   // exactly representable in TF32
        const float lo = v - hi;
        if (j < P.kpad) {
          *reinterpret_cast<float*>(sBx + (j >> 5) * 1024 + box_off(0, j & 31)) = hi;
          *reinterpret_cast<float*>(sBx + (j >> 5) * 1024 + box_off(1, j & 31)) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(c2_ready);
      }
      TC_WAIT(3, &tmem_full[tb], tph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float best = kNegHuge; int besti = 0;
      #pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int c0 = ch * 128 + half * 64;
        if (c0 < P.kpad) {
          uint32_t v0[32], v1[32];
          const uint32_t ta = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(tb * 256 + c0);
          tmem_ld32(ta, v0);
          tmem_ld32(ta + 32, v1);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          if (P.dbg && t == 0) {
            for (int j = 0; j < 32; ++j) { P.dbg[row * 256 + c0 + j] = __uint_as_float(v0[j]); P.dbg[row * 256 + c0 + 32 + j] = __uint_as_float(v1[j]); }
          }
          const bool v1ok = c0 + 32 < P.kpad;                          // beyond kpad: stale TMEM
          // Packed arg-max: the column index rides in the 8 low mantissa bits of the score (a perturbation of 2^-15
          // relative, far below the TF32 rounding of the score itself), so one LOP3 + (half of) one 3-input FMNMX per
          // column replace compare + two selects -- the arg-max warps were the kernel's issue-slot limiter (63 % busy).
          // 255 - column: among equal truncated positive scores the lowest column wins, as before.
          // The index embedded per column is the compile-time position inside this 64-column chunk (63 - jj: the lowest
          // column wins among equal truncated positive scores, as before); the chunk offset is resolved once per chunk.
          float b0 = kNegHuge, b1 = kNegHuge;                          // two independent chains for ILP
          #pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float p0 = __uint_as_float((v0[j] & 0xFFFFFF00u) | (uint32_t)(63 - j));
            const float p1 = __uint_as_float((v0[j + 1] & 0xFFFFFF00u) | (uint32_t)(62 - j));
            b0 = fmaxf(fmaxf(b0, p0), p1);
            const float q0 = __uint_as_float((v1[j] & 0xFFFFFF00u) | (uint32_t)(31 - j));
            const float q1 = __uint_as_float((v1[j + 1] & 0xFFFFFF00u) | (uint32_t)(30 - j));
            b1 = fmaxf(fmaxf(b1, q0), q1);
          }
          if (!v1ok) b1 = kNegHuge;
          const float bc = fmaxf(b0, b1);
          if (bc > best) { best = bc; besti = c0 + 63 - (int)(__float_as_uint(bc) & 0xFFu); }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[tb]);
      TC_WAIT(4, &cand_empty[tb], tph ^ 1);                            // finishers are done with this candidate buffer
      cand_v[(tb * 2 + ch) * 128 + row] = best;
      cand_i[(tb * 2 + ch) * 128 + row] = besti;
      __syncwarp();
      if (lane == 0) mbar_arrive(&cand_full[tb]);                      // (mbarrier.arrive has release semantics)
    }
  } else {
    // ===================================================== finisher warps: 2 groups x 4 warps; group f takes tiles with parity f.
    // Merge the two column halves, recompute the winner's similarity exactly in fp32, store, release the A stage.
    const int f = (warp - 10) >> 2;
    const int row = ((warp - 10) & 3) * 32 + lane;
    int cur_l = -1; uint32_t bgen = 0;
    long long it = 0;
    for (long long t = t_begin; t < t_end; t += t_step, ++it) {
      const int li = (int)(t / P.tiles_per_l), ti = (int)(t % P.tiles_per_l);
      const bool last_of_l = (t + t_step >= t_end) || ((int)((t + t_step) / P.tiles_per_l) != li);
      if (li != cur_l) { mbar_wait(b_full, bgen & 1); ++bgen; cur_l = li; }   // B tile of this k-means is in shared memory
      const int tb = (int)(it & 1);
      if (tb == f) {
        const int s = (int)(it % TC_STAGES); const uint32_t tph = (uint32_t)(it >> 1) & 1;
        const uint32_t aph = (uint32_t)(it / TC_STAGES) & 1;
        // The swizzle term of box_off has period 4 in the row: four precomputed bases per operand, the row offset
        // (e * 128) folds into the load's immediate after unrolling.
        const uint8_t* xa = sA + s * (size_t)a_bytes + (row >> 5) * box_bytes;
        const uint8_t* xq[4];
        #pragma unroll
        for (int r = 0; r < 4; ++r) xq[r] = xa + box_off(r, row & 31);
        float x2 = 0.f;
        if (!P.exact_values) {
          // |x|^2 as soon as the tile has LANDED (in parallel with its MMAs), then the A stage is released at once: the
          // stage used to stay busy through MMA + arg-max + finisher (ncu: HBM 42 %, tensor pipe 48 %, the 4 x 32 KB ring
          // covered barely one memory latency of loads in flight); now only through max(MMA, this loop).
          TC_WAIT(5, &full_a[s], aph);
          #pragma unroll 2
          for (int e4 = 0; e4 < d; e4 += 4) {
            #pragma unroll
            for (int r = 0; r < 4; ++r) { const float xv = *reinterpret_cast<const float*>(xq[r] + e4 * 128); x2 = fmaf(xv, xv, x2); }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_a[s]);
        }
        TC_WAIT(6, &cand_full[tb], tph);
        float best = cand_v[(tb * 2 + 0) * 128 + row]; int besti = cand_i[(tb * 2 + 0) * 128 + row];
        const float ov = cand_v[(tb * 2 + 1) * 128 + row]; const int oi = cand_i[(tb * 2 + 1) * 128 + row];
        if (ov > best) { best = ov; besti = oi; }                      // packed keys: the index bits break exact ties
        __syncwarp();
        if (lane == 0) mbar_arrive(&cand_empty[tb]);
        float acc = 0.f;
        if (P.exact_values) {
          // exact fp32 similarity of the chosen centroid (max_sim.cu:78-98 arithmetic)
          const uint8_t* cb = sB + (besti >> 5) * box_bytes;
          const uint8_t* cq[4];
          #pragma unroll
          for (int r = 0; r < 4; ++r) cq[r] = cb + box_off(r, besti & 31);
          #pragma unroll 2
          for (int e4 = 0; e4 < d; e4 += 4) {
            #pragma unroll
            for (int r = 0; r < 4; ++r) {
              const float xv = *reinterpret_cast<const float*>(xq[r] + e4 * 128);
              const float cvv = *reinterpret_cast<const float*>(cq[r] + e4 * 128);   // random column: bank conflicts
              const float dif = xv - cvv;
              acc = fmaf(-dif, dif, acc);
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_a[s]);                     // this warp is done reading the A stage
        } else {
          acc = fmaf(2.f, best, -x2);                                  // -|x - c|^2 = 2 (<x,c> - |c|^2/2) - |x|^2 with the TF32 score
        }
        const long long p = (long long)ti * TC_M + row;
        if (p < P.n) {
          P.maxsims[(size_t)li * P.n + p] = acc;
          P.labels[(size_t)li * P.n + p] = besti;
        }
      }
      if (last_of_l) {
        __syncwarp();
        if (lane == 0) mbar_arrive(b_free);                            // done with this k-means' B tile (both groups arrive)
      }
    }
  }
#ifdef TPQ_DEBUG_KNOBS
  if (P.wait && threadIdx.x == 0) atomicAdd(P.wait + 7, (unsigned long long)(clock64() - t_cta0));
#endif
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem_base) : "memory");
  }
}

// ----------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [l][rows][cols] fp32, cols contiguous; box = 32 cols x rows x 1, 128-byte swizzle, zero fill out of bounds
static int make_map(CUtensorMap* map, const float* base, int l, int rows, long long cols) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return TPQ_ERR_CUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)l};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 4, (cuuint64_t)cols * rows * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return TPQ_ERR_CUDA; }
  return TPQ_OK;
}

// shapes the tensor-core kernel takes; everything else goes to the exact SIMT kernel
bool assign_tc_supported(int l, int d, long long n, int k, const void* data, const void* cent) {
  return l >= 1 && d >= 8 && d <= 64 && d % 8 == 0 && k >= 8 && k <= 256 && k % 4 == 0 && n % 4 == 0 && n >= 128 &&
         n < (1ll << 31) && (reinterpret_cast<uintptr_t>(data) & 15) == 0 && (reinterpret_cast<uintptr_t>(cent) & 15) == 0;
}

int launch_assign_tc(const float* data, const float* cent, int l, int d, long long n, int k, int exact_values,
                     float* maxsims, int64_t* labels, cudaStream_t st) {
  CUtensorMap mx, mc;
  if (int rc = make_map(&mx, data, l, d, n)) return rc;
  if (int rc = make_map(&mc, cent, l, d, k)) return rc;
  TcParams P;
  P.l = l; P.d = d; P.n = (int)n; P.k = k; P.kpad = (k + 31) / 32 * 32;
  P.tiles_per_l = (int)((n + TC_M - 1) / TC_M);
  P.maxsims = maxsims; P.labels = labels; P.cent = cent; P.dbg = g_tc_dbg; P.exact_values = exact_values;
  P.wait = g_tc_wait;
  const size_t box = (size_t)d * 128;
  const size_t smem = (size_t)(P.kpad / 32) * box + (size_t)TC_STAGES * 4 * box + 12 * 1024 + 256 * 4 + 32 * 8 + 8192 + 1024;
  TPQ_CUDA(cudaFuncSetAttribute(assign_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long total = (long long)l * P.tiles_per_l;
  int grid = (int)(total < sms ? total : sms);
  assign_tc_kernel<<<grid, TC_THREADS, smem, st>>>(mx, mc, P);
  TPQ_LAUNCH_CHECK("assign_tc_kernel");
  return TPQ_OK;
}

}  // namespace tpq

extern "C" void tpq_debug_set_tc_dump(float* p) { tpq::g_tc_dbg = p; }
extern "C" void tpq_debug_set_tc_wait(unsigned long long* p) { tpq::g_tc_wait = p; }
