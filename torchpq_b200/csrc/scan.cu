// The hot path: fused IVF list scan + top-k over the scan layout (relayout.cu).
//
// Replaces the reference's ivfpq_topk kernel (torchpq/kernels/cuda/ivfpq_topk.cu:822-971,
// >90 % of search time there) together with the LUT hand-off, the [nq,n_probe] gathers
// (IVFPQIndex.py:420-421) and get_id_by_address (container/BaseContainer.py:58-65).
//
// Per (query, slice) CTA:
//   * the query's LUT sits in shared memory TRANSPOSED, lut[g][code][s] (g = m / 64,
//     s = m % 64, fp32, 256-byte rows), so an entry's byte address is
//         g*65536 + code*256 + s*4
//     and, because the codes are lane-rotated (relayout.cu), lane l at step t reads slot
//     s = 32*(h&1) + ((l + t) & 31): 32 lanes -> 32 distinct banks, zero conflicts.
//   * one PRMT builds the whole address: byte 1 = the code byte, byte 0 = the lane's
//     precomputed slot offset, bytes 2..3 = sign-replicated zero; the half-group and
//     group offsets are LDS immediates.  Inner loop = PRMT + LDS + FADD per code byte.
//   * the LUT itself is built by the CTA, in place, from the L2-resident transposed codebook (d/M in
//     {1,2,4}; otherwise it is staged through HBM by lut_scan_kernel).
//   * each warp streams whole 32-vector blocks with coalesced 128-bit loads in units of 64
//     sub-quantizers, software pipelined one unit ahead (ping-pong register sets).
//   * top-k (CtaTopK, common.cuh): one compare against the CTA-wide running k-th best + one ballot per
//     block; survivors are staged per warp and merged into the CTA's single sorted list under a lock;
//     the threshold is bootstrapped, and the buffers are drained, by CTA-wide sorts.
//   * RES variant: residual IVFPQ (per-cell LUT = query half + precomputed cell half).
// Sub-quantizer sums are fp32 in four interleaved partial sums (exact for integer LUTs;
// ~1e-7 relative otherwise -- the reference's m-ascending order lives in scan_ref.cu).
// Experiment knobs (TPQ_SCAN_CFG, TPQ_BOOT_R, TPQ_LUT_MODE) and per-phase cycle counters exist only in the
// -DTPQ_DEBUG_KNOBS build (make dbg -> libtpq_b200_dbg.so, used by scripts/); the product library reads no environment.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace tpq {

// ----------------------------------------------------------------------------- LUT in scan layout (global)
// lut_scan[q][g][code][s] = LUT[64 g + s][q][code], zero for padding sub-quantizers.  Same arithmetic as lut.cu
// (PQCodec.py:62-75, MultiKMeans.py:183-223).  Used when d/M is too large for the in-CTA build.
// CTA = 256 codes x 16 sub-quantizers x LS_QT queries: every codebook value is loaded once (coalesced over the
// code) and reused for LS_QT queries; query sub-vectors are broadcast from shared memory.
constexpr int LS_QT = 8, LS_MT = 16;
constexpr int kLutPart1 = 2;          // lut_scan_kernel `metric` value: the query half of the residual LUT
__global__ void __launch_bounds__(256)
lut_scan_kernel(const float* __restrict__ x, const float* __restrict__ cb, int d, int M, int MP, int nq,
                int q_base, int n_chunk, int metric, float* __restrict__ lut_scan) {
  extern __shared__ __align__(16) float xs[];            // [LS_QT][LS_MT * dsub] sub-vectors, then [LS_QT][LS_MT] |x_m|^2
  const int dsub = d / M;
  const int m0 = blockIdx.y * LS_MT;
  const int q0 = blockIdx.x * LS_QT;                     // within this chunk
  const int c = threadIdx.x;
  float* a2s = xs + LS_QT * LS_MT * dsub;
  for (int i = threadIdx.x; i < LS_QT * LS_MT * dsub; i += blockDim.x) {
    const int qq = i / (LS_MT * dsub), r = i % (LS_MT * dsub);
    const int row = m0 * dsub + r;
    xs[i] = (q0 + qq < n_chunk && row < d) ? x[(size_t)row * nq + q_base + q0 + qq] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < LS_QT * LS_MT; i += blockDim.x) {
    const float* v = xs + (size_t)i * dsub;
    float a2 = 0.f;
    for (int e = 0; e < dsub; ++e) a2 = __fadd_rn(a2, __fmul_rn(v[e], v[e]));
    a2s[i] = a2;
  }
  __syncthreads();
  const int MG = (MP + 63) / 64;
  for (int mm = 0; mm < LS_MT; ++mm) {
    const int m = m0 + mm;
    if (m >= MP) break;
    float dot[LS_QT];
    #pragma unroll
    for (int qq = 0; qq < LS_QT; ++qq) dot[qq] = 0.f;
    float b2 = 0.f;
    if (m < M) {
      const float* p = cb + (size_t)m * dsub * 256 + c;
      #pragma unroll 4
      for (int e = 0; e < dsub; ++e) {
        const float pv = __ldg(p + (size_t)e * 256);
        b2 = __fadd_rn(b2, __fmul_rn(pv, pv));
        #pragma unroll
        for (int qq = 0; qq < LS_QT; ++qq) dot[qq] = fmaf(xs[(qq * LS_MT + mm) * dsub + e], pv, dot[qq]);
      }
    }
    const int g = m >> 6, sl = m & 63;
    #pragma unroll
    for (int qq = 0; qq < LS_QT; ++qq) {
      if (q0 + qq >= n_chunk) break;
      float y = 0.f;
      if (m < M) {
        y = dot[qq];
        if (metric == TPQ_METRIC_EUCLIDEAN) y = __fsub_rn(__fsub_rn(__fmul_rn(dot[qq], 2.f), a2s[qq * LS_MT + mm]), b2);
        else if (metric == kLutPart1) y = __fmul_rn(dot[qq], 2.f);   // residual IVFPQ: part1 = 2 <x_m, p> (IVFPQIndex.py:377)
      }
      lut_scan[((size_t)(q0 + qq) * MG + g) * 16384 + c * 64 + sl] = y;
    }
  }
}

// ----------------------------------------------------------------------------- scan kernel
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}
// LUT entry for code byte J of `w`, slot offset byte J of `off`; imm = group/half-group byte offset.
template <int J>
__device__ __forceinline__ float lut_at(const uint8_t* lut, uint32_t w, uint32_t off, int imm) {
  constexpr uint32_t sel = (4u + J) | (uint32_t(J) << 4) | ((8u | (4u + J)) << 8) | ((8u | (4u + J)) << 12);
  return *reinterpret_cast<const float*>(lut + prmt(w, off, sel) + imm);
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// A block (32 vectors x MP code bytes) is consumed in UNITS of up to 64 sub-quantizers (4 chunks = 16 words per
// lane), so the register footprint and the software pipeline are the same for every M: while unit u is being
// looked up, unit u+1 (the next 64 sub-quantizers of the block, or the first unit of the warp's next block) is in flight.
template <int MP>
struct Units {
  static constexpr int NSB = (MP + 63) / 64;
  __host__ __device__ static constexpr int words(int sb) { return (sb == NSB - 1 && (MP % 64) != 0) ? 8 : 16; }
};

template <int MP, int SB>
__device__ __forceinline__ void load_unit(uint32_t (&w)[16], const uint8_t* __restrict__ codes, int64_t B, int lane) {
  const uint4* p = reinterpret_cast<const uint4*>(codes + (size_t)B * (MP * 32)) + SB * 128 + lane;
  #pragma unroll
  for (int j = 0; j < Units<MP>::words(SB) / 4; ++j) {
    const uint4 v = ldg_stream(p + j * 32);
    w[4 * j + 0] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
  }
}

template <int MP, int SB>
__device__ __forceinline__ void adc_unit(const uint32_t (&w)[16], const uint32_t (&off)[8], const uint8_t* lut, float (&acc)[4]) {
  #pragma unroll
  for (int i = 0; i < Units<MP>::words(SB); ++i) {
    const int imm = SB * 65536 + (i >> 3) * 128;        // group SB, half-group i/8
    acc[0] += lut_at<0>(lut, w[i], off[i & 7], imm);
    acc[1] += lut_at<1>(lut, w[i], off[i & 7], imm);
    acc[2] += lut_at<2>(lut, w[i], off[i & 7], imm);
    acc[3] += lut_at<3>(lut, w[i], off[i & 7], imm);
  }
}

// Per-warp scan state.  regs[2][16] is the ping-pong pair of unit buffers; every index into it is a
// compile-time constant after unrolling (PAR / SB are template parameters), so it lives in registers.
template <int MP, int NW>
struct Scanner {
  static constexpr int NSB = Units<MP>::NSB;
  const uint8_t* codes; const uint32_t* valid; const uint8_t* lut;
  const int32_t* seg_prefix; const int32_t* seg_blk0; const uint32_t* seg_addr0;
  int* next;                                            // CTA-wide cursor: next block (of the concatenated probe list) to hand out
  int lane;
  uint32_t off[8];
  uint32_t regs[2][16];
  CtaTopK tk;
  uint64_t thr;
  float base;                                           // residual IVFPQ: coarse similarity of the current cell (else 0)
  int seg, seg_lo, seg_hi, seg_b0; uint32_t seg_a0;     // current probe segment, cached in registers
  int64_t B_cur, B_nxt; uint32_t a0_cur, a0_nxt, valid_cur, valid_nxt;

  __device__ __forceinline__ void locate(int b, int64_t& B, uint32_t& addr0) {
    if (b >= seg_hi) {
      do { ++seg; seg_lo = seg_hi; seg_hi = seg_prefix[seg + 1]; } while (b >= seg_hi);
      seg_b0 = seg_blk0[seg]; seg_a0 = seg_addr0[seg];
    }
    const int rel = b - seg_lo;
    B = (int64_t)seg_b0 + rel;
    addr0 = seg_a0 + (uint32_t)rel * 32u;
  }
  template <int PAR, int SB>
  __device__ __forceinline__ void step(float (&acc)[4], bool has_next) {
    constexpr int BUF = (PAR + SB) & 1;
    if constexpr (SB + 1 < NSB) {
      load_unit<MP, SB + 1>(regs[BUF ^ 1], codes, B_cur, lane);
    } else {
      if (has_next) { load_unit<MP, 0>(regs[BUF ^ 1], codes, B_nxt, lane); valid_nxt = __ldg(valid + B_nxt); }
    }
    adc_unit<MP, SB>(regs[BUF], off, lut, acc);
    if constexpr (SB + 1 < NSB) step<PAR, SB + 1>(acc, has_next);
  }
  // consume the block at the cursor (its unit 0 is already in regs[PAR]); prefetch the next block's unit 0
  template <int PAR>
  __device__ __forceinline__ void process_block(bool has_next) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    step<PAR, 0>(acc, has_next);
    const float score = base + ((acc[0] + acc[1]) + (acc[2] + acc[3]));
    const uint64_t key = make_key(score, a0_cur + lane);
    const bool live = (valid_cur >> lane) & 1u;
    if (PAR == 0) thr = tk.threshold();                  // running k-th best of the CTA
    tk.push(live && key > thr, key, lane);
    B_cur = B_nxt; a0_cur = a0_nxt; valid_cur = valid_nxt;
  }
  template <int SB>
  __device__ __forceinline__ void simple_units(float (&acc)[4], int64_t B) {
    load_unit<MP, SB>(regs[0], codes, B, lane);
    adc_unit<MP, SB>(regs[0], off, lut, acc);
    if constexpr (SB + 1 < NSB) simple_units<SB + 1>(acc, B);
  }
  // score one block without pipelining and return this lane's key (0 if the slot is not live)
  __device__ __forceinline__ uint64_t key_of_block(int b) {
    int64_t B; uint32_t a0;
    locate(b, B, a0);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    simple_units<0>(acc, B);
    const uint32_t v = __ldg(valid + B);
    const float score = base + ((acc[0] + acc[1]) + (acc[2] + acc[3]));
    return ((v >> lane) & 1u) ? make_key(score, a0 + lane) : 0ull;
  }
  // Blocks are handed out dynamically: `next` is a CTA-wide counter in shared memory.  (A static stripe b = warp,
  // warp + NW, ... left the CTA waiting for its slowest warp -- the warp schedulers favour the oldest warps, so the
  // youngest ones finished 5 % (C3) to 30 % (C4, 16 warps, one CTA per SM) later than warp 0.)  The index of the
  // block after next is fetched before the current block is consumed, so the atomic's latency is off the critical path.
  __device__ __forceinline__ int grab() {
    int b = 0;
    if (lane == 0) b = atomicAdd(next, 1);
    return __shfl_sync(0xffffffffu, b, 0);
  }
  __device__ __forceinline__ void run(int b_end) {
    int b = grab();
    if (b >= b_end) return;
    locate(b, B_cur, a0_cur);
    load_unit<MP, 0>(regs[0], codes, B_cur, lane);
    valid_cur = __ldg(valid + B_cur);
    for (;;) {
      b = grab();
      bool more = b < b_end;
      if (more) locate(b, B_nxt, a0_nxt);
      process_block<0>(more);
      if (!more) break;
      if constexpr (NSB % 2 == 1) {                       // odd unit count: the buffers swap roles every block
        b = grab();
        more = b < b_end;
        if (more) locate(b, B_nxt, a0_nxt);
        process_block<1>(more);
        if (!more) break;
      }
    }
  }
};

struct ScanSmem { size_t lut, part1, seg_blk0, seg_addr0, seg_cell, seg_prefix, thr, lock, rep, list, bufs, total; };
static ScanSmem scan_smem(int MP, int n_probe, int nw, int kp, bool residual = false, bool part1_in_smem = false) {
  ScanSmem s; size_t off = 0;
  s.lut = off;        off += (size_t)((MP + 63) / 64) * 65536;
  s.part1 = off;      off += (residual && part1_in_smem) ? (size_t)((MP + 63) / 64) * 65536 : 0;   // query half of the residual LUT
  s.seg_cell = off;   off += residual ? (size_t)n_probe * 4 : 0;
  s.seg_blk0 = off;   off += (size_t)n_probe * 4;
  s.seg_addr0 = off;  off += (size_t)n_probe * 4;
  s.seg_prefix = off; off += (size_t)(n_probe + 1) * 4;
  off = align_up(off, 8);
  s.thr = off;        off += 8;
  s.lock = off;       off += 16;                                      // list lock, drain counter, block cursor
  s.rep = off;        off += (size_t)nw * 8;
  s.list = off;       off += (size_t)kp * 8;
  s.bufs = off;       off += (size_t)nw * kStage * 8;
  s.total = off;
  return s;
}

struct ScanArgs {
  const uint8_t* codes; const uint32_t* valid; const int32_t* cell_block_start; const int64_t* cell_start;
  const int64_t* cell_size;
  const float* lut_scan;          // [nq_chunk][MG][256][64] (staged-LUT variant, DSUB == 0)
  const float* x;                 // [d, nq] queries (in-CTA LUT variant)
  const float* cbt;               // pq_codebook_t [256, MP, dsub]
  const float* part2_scan;        // residual: [C][MG][256][64] per-cell half of the LUT
  const float* base_sims;         // residual: [nq, n_probe] coarse similarities
  int M, metric;
  const int64_t* cells;           // [nq, n_probe]
  const int64_t* n_probe_list;    // [nq]
  uint64_t* keys_out;             // [nq, S, k]
  uint64_t* peer_keys[8];         // fused exchange: every rank's gather buffer [world, nq, k] (NVLink peer memory)
  int n_peers, peer_slot;         // n_peers == 0: keys go to keys_out
  int nq, q_base, n_probe, k, kp, S;
  int boot_r;                     // blocks per warp in the bootstrap (1 or 2)
  int n_cells;                    // probe entries outside [0, n_cells) are treated as empty segments
#ifdef TPQ_DEBUG_KNOBS
  unsigned long long* phase;      // [8] summed clock64 deltas: prologue, bootstrap, scan, drain; [4] = CTAs
  int boot_mode;                  // 1 = merge-tree bootstrap (A/B against the quick start)
#endif
};

// The CTA's sorted top-k keys leave the SM here.  Fused exchange (n_peers > 0, dist.py): instead of a local buffer that an
// NCCL all-gather would then copy, the CTA stores its k keys straight into slot `peer_slot` of EVERY rank's gather buffer
// over NVLink (coalesced 8-byte stores, k * 8 B per peer), so the transfer overlaps the scans still running; a
// symmetric-memory barrier and the merge kernel follow.
template <int NW>
__device__ __forceinline__ void write_keys(const ScanArgs& A, const uint64_t* list, int q, int slice, int tid) {
  if (A.n_peers > 0) {
    const size_t o = ((size_t)A.peer_slot * A.nq + q) * A.k;
    for (int p = 0; p < A.n_peers; ++p) {
      uint64_t* out = A.peer_keys[p] + o;
      for (int i = tid; i < A.k; i += NW * 32) out[i] = list[i];
    }
    return;
  }
  uint64_t* out = A.keys_out + ((size_t)q * A.S + slice) * A.k;
  for (int i = tid; i < A.k; i += NW * 32) out[i] = list[i];
}

// DSUB > 0: the CTA builds its query's LUT itself from the (L2-resident) transposed codebook -- no LUT
// round trip through HBM, no LUT workspace.  DSUB == 0: the LUT was written by lut_scan_kernel and is copied in.
// RES: residual IVFPQ (ivfpq_topk_residual_precomputed, ivfpq_topk.cu:1039-1207).  The LUT of a (query, cell) pair
// is part1[query] + part2[cell]: the query half is built once into a second shared-memory table, and the CTA
// walks the probed cells in lock step, re-forming the LUT (one float4 add per 16 bytes) before each cell; the
// cell's coarse similarity seeds every score.
template <int MP, int NW, int MINB, int DSUB, bool RES>
__global__ void __launch_bounds__(NW * 32, MINB)
ivfpq_scan_kernel(ScanArgs A, ScanSmem L) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* lut = smem;                                       // L.lut == 0 by construction
  int32_t*  seg_blk0   = reinterpret_cast<int32_t*>(smem + L.seg_blk0);
  uint32_t* seg_addr0  = reinterpret_cast<uint32_t*>(smem + L.seg_addr0);
  int32_t*  seg_prefix = reinterpret_cast<int32_t*>(smem + L.seg_prefix);
  constexpr int MG = (MP + 63) / 64;

  const int qi = blockIdx.x / A.S, slice = blockIdx.x % A.S;  // query within this chunk
  const int q = A.q_base + qi;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#ifdef TPQ_DEBUG_KNOBS
  long long t_ph = clock64();
#endif

  // --- probe segments first (warp 0): their dependent global loads overlap the other warps' LUT work (same visiting rules as scan_ref.cu / ivfpq_topk.cu:837-870)
  int P = (int)A.n_probe_list[q];
  P = RES ? max(0, min(P, A.n_probe))                         // residual kernel: `for cCell < nProbe` (:1080), may be empty
          : max(1, min(P, A.n_probe));                        // plain kernel: first cell entered unconditionally
  int32_t* seg_cell = reinterpret_cast<int32_t*>(smem + L.seg_cell);
  if (warp == 0) {
    int carry = 0;
    if (lane == 0) seg_prefix[0] = 0;
    const int64_t* cq = A.cells + (size_t)q * A.n_probe;
    for (int j0 = 0; j0 < P; j0 += 32) {
      const int j = j0 + lane;
      int nb = 0;
      if (j < P) {
        const int64_t c = cq[j];
        const bool ok = c >= 0 && c < A.n_cells;            // caller-supplied probe lists (search_cells) are not trusted
        const int64_t s = ok ? A.cell_start[c] : -1;
        bool skip = !ok;
        if (ok && j > 0) {
          const int64_t cp = cq[j - 1];
          skip = cp >= 0 && cp < A.n_cells && s == A.cell_start[cp];
        }
        const int b0 = ok ? A.cell_block_start[c] : 0;
        if (!skip) {                                        // the layout holds a block for every slot of the cell's capacity
          const int have = A.cell_block_start[c + 1] - b0;  // (0 for another shard's cell); the live extent is cell_size
          const int64_t sz = A.cell_size[c];
          nb = (int)min((int64_t)have, sz > 0 ? (sz + 31) >> 5 : (int64_t)0);
        }
        seg_blk0[j] = b0;
        seg_addr0[j] = (uint32_t)s;
        if constexpr (RES) seg_cell[j] = ok ? (int32_t)c : 0;
      }
      int incl = nb;
      #pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
      if (j < P) seg_prefix[j + 1] = carry + incl;
      carry += __shfl_sync(0xffffffffu, incl, 31);
    }
  }
  if constexpr (DSUB == 0 && RES) {
    // residual with a staged query half: part1 stays in global memory (L2) and is added to part2 cell by cell
  } else if constexpr (DSUB == 0) {
    // --- stage LUT (MG * 64 KB, coalesced 16-byte copies)
    const float4* src = reinterpret_cast<const float4*>(A.lut_scan + (size_t)qi * MG * 16384);
    float4* dst = reinterpret_cast<float4*>(lut);
    #pragma unroll 4
    for (int i = tid; i < MG * 4096; i += NW * 32) dst[i] = __ldg(src + i);
  } else {
    // --- build the LUT in place (PQCodec.py:62-75; MultiKMeans.py:183-223; same arithmetic as lut.cu).
    // Lane l owns sub-quantizers m = 32 h + l: reads of the transposed codebook are contiguous across the
    // warp and the store bank is l, so both sides are conflict-free / coalesced.
    constexpr int NH = MP / 32;
    float xr[NH][DSUB], a2[NH];
    #pragma unroll
    for (int h = 0; h < NH; ++h) {
      const int m = 32 * h + lane;
      float s2 = 0.f;
      #pragma unroll
      for (int i = 0; i < DSUB; ++i) {
        const float v = (m < A.M) ? __ldg(A.x + (size_t)(m * DSUB + i) * A.nq + q) : 0.f;
        xr[h][i] = v;
        s2 = __fadd_rn(s2, __fmul_rn(v, v));
      }
      a2[h] = s2;
    }
    float* lutf = reinterpret_cast<float*>(RES ? smem + L.part1 : lut);
    #pragma unroll (DSUB >= 8 ? 2 : 8)
    for (int c = warp; c < 256; c += NW) {
      #pragma unroll
      for (int h = 0; h < NH; ++h) {
        const float* p = A.cbt + ((size_t)c * MP + 32 * h + lane) * DSUB;
        float pv[DSUB];
        if constexpr (DSUB == 2) { const float2 t = __ldg(reinterpret_cast<const float2*>(p)); pv[0] = t.x; pv[1] = t.y; }
        else if constexpr (DSUB % 4 == 0) {
          #pragma unroll
          for (int i = 0; i < DSUB; i += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p + i));
            pv[i] = t.x; pv[i + 1] = t.y; pv[i + 2] = t.z; pv[i + 3] = t.w;
          }
        }
        else { for (int i = 0; i < DSUB; ++i) pv[i] = __ldg(p + i); }
        float dot = 0.f, b2 = 0.f;
        #pragma unroll
        for (int i = 0; i < DSUB; ++i) { dot = fmaf(xr[h][i], pv[i], dot); b2 = __fadd_rn(b2, __fmul_rn(pv[i], pv[i])); }
        float y = dot;
        if constexpr (RES) y = __fmul_rn(dot, 2.f);                    // part1 = 2 <x_m, p> (IVFPQIndex.py:377)
        else if (A.metric == TPQ_METRIC_EUCLIDEAN) y = __fsub_rn(__fsub_rn(__fmul_rn(dot, 2.f), a2[h]), b2);
        lutf[(h >> 1) * 16384 + c * 64 + (h & 1) * 32 + lane] = y;
      }
    }
  }
#ifdef TPQ_DEBUG_KNOBS
  auto phase_mark = [&](int i) {
    if (A.phase && tid == 0) { const long long t = clock64(); atomicAdd(A.phase + i, (unsigned long long)(t - t_ph)); t_ph = t; }
  };
#define TPQ_PHASE(i) phase_mark(i)
#else
#define TPQ_PHASE(i) ((void)0)
#endif
  Scanner<MP, NW> sc;
  sc.codes = A.codes; sc.valid = A.valid; sc.lut = lut;
  sc.seg_prefix = seg_prefix; sc.seg_blk0 = seg_blk0; sc.seg_addr0 = seg_addr0; sc.lane = lane;
  sc.next = reinterpret_cast<int*>(smem + L.lock) + 2;
  sc.tk.init(reinterpret_cast<uint64_t*>(smem + L.list), reinterpret_cast<unsigned long long*>(smem + L.thr),
             reinterpret_cast<int*>(smem + L.lock), reinterpret_cast<uint64_t*>(smem + L.bufs) + (size_t)warp * kStage,
             A.kp, A.k);
  // lane's slot offsets: byte r%4 of off[r/4] = ((lane + r) & 31) * 4
  #pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint32_t v = 0;
    #pragma unroll
    for (int b = 0; b < 4; ++b) v |= (uint32_t)(((lane + 4 * i + b) & 31) * 4) << (8 * b);
    sc.off[i] = v;
  }
  __syncthreads();
  TPQ_PHASE(0);

  const int total = seg_prefix[P];
  const int b_begin = (int)(((int64_t)total * slice) / A.S);
  const int b_end = (int)(((int64_t)total * (slice + 1)) / A.S);
  sc.thr = 0; sc.base = 0.f;
  sc.seg = 0; sc.seg_lo = 0; sc.seg_hi = seg_prefix[1]; sc.seg_b0 = seg_blk0[0]; sc.seg_a0 = seg_addr0[0];
  CtaTopK& tk = sc.tk;
  uint64_t* bufs = reinterpret_cast<uint64_t*>(smem + L.bufs);
  if constexpr (RES) {
    // query half of the LUT: the shared-memory table built above (DSUB > 0), or the staged copy in global memory
    const float4* p1 = DSUB > 0 ? reinterpret_cast<const float4*>(smem + L.part1)
                                : reinterpret_cast<const float4*>(A.lut_scan + (size_t)qi * MG * 16384);
    float4* dst = reinterpret_cast<float4*>(lut);
    for (int j = 0; j < P; ++j) {
      const int b0 = max(seg_prefix[j], b_begin), b1 = min(seg_prefix[j + 1], b_end);
      if (b1 <= b0) continue;                                          // skipped entry, empty cell, another shard's or slice's blocks
      __syncthreads();                                                 // every warp is done with the previous cell's LUT
      if (tid == 0) *sc.next = b0;
      const float4* p2 = reinterpret_cast<const float4*>(A.part2_scan + (size_t)seg_cell[j] * MG * 16384);
      #pragma unroll 4
      for (int i = tid; i < MG * 4096; i += NW * 32) {                 // store_precomputed_to_smem: part1 + part2 (:609-628)
        const float4 a = DSUB > 0 ? p1[i] : __ldg(p1 + i), b = __ldg(p2 + i);
        dst[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
      }
      __syncthreads();
      sc.base = A.base_sims[(size_t)q * A.n_probe + j];
      sc.run(b1);
    }
    tk.cta_drain(bufs, lane, warp);
    write_keys<NW>(A, tk.list, q, slice, tid);
    return;
  }
  // bootstrap: the first R blocks of every warp go to the list unfiltered through one register sort per warp and a
  // log2(NW)-level merge tree, which establishes the threshold (k-th best of the first R * NW * 32 vectors) without
  // R * NW lock-serialised flushes.
  const int R = A.boot_r;
  if (tid == 0) *sc.next = b_begin + R * NW;                    // the bootstrap's barriers publish it
  {
    #pragma unroll 1
    for (int r = 0; r < R; ++r) {
      const int b = b_begin + warp + r * NW;
      tk.buf[r * 32 + lane] = (b < b_end) ? sc.key_of_block(b) : 0ull;
    }
    tk.cnt = R * 32;
#ifdef TPQ_DEBUG_KNOBS
    if (A.boot_mode == 1) tk.cta_bootstrap(bufs, NW, R, lane, warp); else
#endif
    tk.cta_quickstart(reinterpret_cast<uint64_t*>(smem + L.rep), NW, R, lane, warp);
  }
  TPQ_PHASE(1);
  sc.run(b_end);
  TPQ_PHASE(2);
  tk.cta_drain(bufs, lane, warp);                          // leftovers of all warps: register sorts + merge tree, no lock
  TPQ_PHASE(3);
#ifdef TPQ_DEBUG_KNOBS
  if (A.phase && tid == 0) atomicAdd(A.phase + 4, 1ull);
#endif
  write_keys<NW>(A, tk.list, q, slice, tid);
}

// ----------------------------------------------------------------------------- merge + decode
// keys_in: n_parts sorted lists of k keys per query -> top-k, decoded.
// One warp per query.  (Cross-slice merge for small batches and cross-shard merge after the
// all-gather; with n_parts == 1 it is just the decode + get_id_by_address, BaseContainer.py:58-65.)
__global__ void __launch_bounds__(128)
merge_topk_kernel(const uint64_t* __restrict__ keys_in, int nq, int n_parts, int k, int kp,
                  int64_t part_stride, int64_t query_stride,
                  const int64_t* __restrict__ address2id, int64_t capacity,
                  float* __restrict__ values, int64_t* __restrict__ ids, int64_t* __restrict__ address,
                  uint64_t* __restrict__ keys_out) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * (blockDim.x >> 5) + warp;
  if (q >= nq) return;
  uint64_t* A = reinterpret_cast<uint64_t*>(smem) + (size_t)warp * 2 * kp;
  uint64_t* Bf = A + kp;
  const uint64_t* src = keys_in + (size_t)q * query_stride;
  for (int i = lane; i < kp; i += 32) A[i] = i < k ? src[i] : 0ull;
  __syncwarp();
  for (int p = 1; p < n_parts; ++p) {
    const uint64_t* sp = src + (size_t)p * part_stride;
    for (int i = lane; i < kp; i += 32) Bf[i] = i < k ? sp[i] : 0ull;
    __syncwarp();
    warp_merge_desc(A, kp, Bf, kp, lane);
  }
  for (int i = lane; i < k; i += 32) {
    const uint64_t key = A[i];
    const size_t o = (size_t)q * k + i;
    const int64_t a = key ? (int64_t)key_addr(key) : -1;
    if (values) values[o] = key ? key_score(key) : -INFINITY;
    if (address) address[o] = a;
    if (ids) ids[o] = (a >= 0 && a < capacity) ? address2id[a] : -1;
    if (keys_out) keys_out[o] = key;
  }
}

static int launch_merge(const uint64_t* keys_in, int nq, int n_parts, int k, int64_t part_stride, int64_t query_stride,
                        const int64_t* address2id, int64_t capacity, float* values, int64_t* ids, int64_t* address,
                        uint64_t* keys_out, cudaStream_t st) {
  const int kp = next_pow2(k < 32 ? 32 : k);
  const int wpb = 4;
  size_t smem = (size_t)wpb * 2 * kp * 8;
  TPQ_CUDA(cudaFuncSetAttribute(merge_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  merge_topk_kernel<<<(nq + wpb - 1) / wpb, wpb * 32, smem, st>>>(keys_in, nq, n_parts, k, kp, part_stride, query_stride,
                                                                 address2id, capacity, values, ids, address, keys_out);
  TPQ_LAUNCH_CHECK("merge_topk_kernel");
  return TPQ_OK;
}

// ----------------------------------------------------------------------------- host orchestration
constexpr int kQueryChunk = 16384;    // queries per scan launch (bounds the LUT workspace: 64 KB * MG each)

// optional CUDA-event timing of the scan launches (bench.py's roofline leg); off by default
constexpr int kProfMax = 1024;
static bool g_prof_on = false;
static int g_prof_n = 0;
static cudaEvent_t g_prof_start[kProfMax], g_prof_stop[kProfMax];
static bool g_prof_created = false;

#ifdef TPQ_DEBUG_KNOBS
static unsigned long long* g_phase = nullptr;     // device buffer set by tpq_debug_set_phase_buffer (debug build only)
#endif

static int pick_slices(const tpq_index* ix, int nq, int k) {
  // enough CTAs for ~2 waves of 2 CTAs/SM when the batch is small
  (void)k;
  const int want = 148 * 4;
  if (nq >= want) return 1;
  int s = (want + nq - 1) / nq;
  return s > 64 ? 64 : s;
}

// LUT source: built inside the scan CTA when d/M is 1, 2, 4 or 8 (the codebook slice a CTA reads from L2, M*dsub KB,
// is then small next to the codes it goes on to stream: C4, d/M = 8, reads 0.98 MB of codebook per 7.5 MB of codes and
// saves the 2 x 128 KB LUT round trip through HBM plus one launch); staged through HBM by lut_scan_kernel otherwise.
static int fused_dsub(const tpq_index* ix) {
  const int dsub = ix->d_vector / ix->n_subvectors;
#ifdef TPQ_DEBUG_KNOBS
  const char* mode = getenv("TPQ_LUT_MODE");
  if (mode && !strcmp(mode, "staged")) return 0;
#endif
  // residual IVFPQ keeps TWO tables when the query half is built in the CTA: that fits for M <= 64 and pays for d/M <= 4
  if (ix->residual && (ix->m_pad > 64 || dsub == 8)) return 0;
#ifdef TPQ_DEBUG_KNOBS
  { const char* rm = getenv("TPQ_RES_MODE"); if (ix->residual && rm && !strncmp(rm, "staged", 6)) return 0; }
#endif
  return (dsub == 1 || dsub == 2 || dsub == 4 || dsub == 8) ? dsub : 0;
}

struct PeerOut { uint64_t* keys[8]; int n, slot; };

struct SearchWs {
  size_t coarse, probe_sims, cells, npl, xnorm, lut, keys, total;
};
static SearchWs search_ws(const tpq_index* ix, int nq, int n_probe, int k) {
  SearchWs w; size_t off = 0;
  const int MG = (ix->m_pad + 63) / 64;
  const int chunk = nq < kQueryChunk ? nq : kQueryChunk;
  const int S = pick_slices(ix, nq, k);
  w.coarse = off;     off += align_up(tpq_coarse_workspace_bytes(ix->d_vector, nq, ix->n_cells), 256);
  w.probe_sims = off; off += align_up((size_t)nq * n_probe * 4, 256);
  w.cells = off;      off += align_up((size_t)nq * n_probe * 8, 256);
  w.npl = off;        off += align_up((size_t)nq * 8, 256);
  w.xnorm = off;      off += align_up((size_t)nq * ix->d_vector * 4, 256);
  w.lut = off;        off += fused_dsub(ix) ? 0 : align_up((size_t)chunk * MG * 65536, 256);
  w.keys = off;       off += align_up((size_t)nq * S * k * 8, 256);
  w.total = off;
  return w;
}

template <int MP, int NW, int MINB, int DSUB, bool RES = false>
static int launch_scan_d(const tpq_index* ix, const float* x, const int64_t* cells, const int64_t* npl,
                         int nq, int n_probe, int k, int S, float* lut_ws, uint64_t* keys, cudaStream_t st,
                         const float* base_sims = nullptr, const PeerOut* peers = nullptr) {
  const int kp = next_pow2(k < 32 ? 32 : k);
  ScanSmem L = scan_smem(MP, n_probe, NW, kp, RES, RES && DSUB > 0);
  if (L.total > 227 * 1024) {
    set_error("scan: M=%d n_probe=%d k=%d needs %zu B of shared memory (> 227 KB)", ix->n_subvectors, n_probe, k, L.total);
    return TPQ_ERR_UNSUPPORTED;
  }
  auto kern = ivfpq_scan_kernel<MP, NW, MINB, DSUB, RES>;
  TPQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
  const int MG = (MP + 63) / 64;
  const int dsub = ix->d_vector / ix->n_subvectors;
  for (int q0 = 0; q0 < nq; q0 += kQueryChunk) {
    const int n = (nq - q0) < kQueryChunk ? (nq - q0) : kQueryChunk;
    if (DSUB == 0) {
      size_t xs_bytes = (size_t)LS_QT * LS_MT * (dsub + 1) * 4;
      dim3 lgrid((n + LS_QT - 1) / LS_QT, (MP + LS_MT - 1) / LS_MT);
      lut_scan_kernel<<<lgrid, 256, xs_bytes, st>>>(x, ix->pq_codebook, ix->d_vector, ix->n_subvectors, MP,
                                                    nq, q0, n, RES ? kLutPart1 : ix->metric, lut_ws);
      TPQ_LAUNCH_CHECK("lut_scan_kernel");
    }
    ScanArgs A;
    A.x = x; A.cbt = ix->pq_codebook_t; A.M = ix->n_subvectors; A.metric = ix->metric;
    A.part2_scan = ix->part2_scan; A.base_sims = base_sims;
    A.codes = ix->codes_scan; A.valid = ix->block_valid; A.cell_block_start = ix->cell_block_start;
    A.cell_start = ix->cell_start; A.cell_size = ix->cell_size; A.lut_scan = lut_ws; A.cells = cells; A.n_probe_list = npl; A.keys_out = keys;
    A.nq = nq; A.q_base = q0; A.n_probe = n_probe; A.k = k; A.kp = kp; A.S = S;
    A.boot_r = 2; A.n_cells = ix->n_cells;
    A.n_peers = 0; A.peer_slot = 0;
    if (peers) { A.n_peers = peers->n; A.peer_slot = peers->slot; for (int p = 0; p < peers->n; ++p) A.peer_keys[p] = peers->keys[p]; }
#ifdef TPQ_DEBUG_KNOBS
    { const char* br = getenv("TPQ_BOOT_R"); if (br && atoi(br) == 1) A.boot_r = 1; }
    A.phase = g_phase;
    { const char* bm = getenv("TPQ_BOOT_MODE"); A.boot_mode = bm ? atoi(bm) : 1; }
#endif
    const bool prof = g_prof_on && g_prof_n < kProfMax;
    if (prof) cudaEventRecord(g_prof_start[g_prof_n], st);
    kern<<<n * S, NW * 32, L.total, st>>>(A, L);
    if (prof) cudaEventRecord(g_prof_stop[g_prof_n++], st);
    TPQ_LAUNCH_CHECK("ivfpq_scan_kernel");
    (void)MG; (void)dsub;
  }
  return TPQ_OK;
}

template <int MP, int NW, int MINB>
static int launch_scan(const tpq_index* ix, const float* x, const int64_t* cells, const int64_t* npl,
                       int nq, int n_probe, int k, int S, float* lut_ws, uint64_t* keys, cudaStream_t st, const PeerOut* peers) {
  switch (fused_dsub(ix)) {
    case 1:  return launch_scan_d<MP, NW, MINB, 1>(ix, x, cells, npl, nq, n_probe, k, S, lut_ws, keys, st, nullptr, peers);
    case 2:  return launch_scan_d<MP, NW, MINB, 2>(ix, x, cells, npl, nq, n_probe, k, S, lut_ws, keys, st, nullptr, peers);
    case 4:  return launch_scan_d<MP, NW, MINB, 4>(ix, x, cells, npl, nq, n_probe, k, S, lut_ws, keys, st, nullptr, peers);
    case 8:  return launch_scan_d<MP, NW, MINB, 8>(ix, x, cells, npl, nq, n_probe, k, S, lut_ws, keys, st, nullptr, peers);
    default: return launch_scan_d<MP, NW, MINB, 0>(ix, x, cells, npl, nq, n_probe, k, S, lut_ws, keys, st, nullptr, peers);
  }
}

static int check_index(const tpq_index* ix) {
  TPQ_REQUIRE(ix, "null index");
  TPQ_REQUIRE(ix->d_vector > 0 && ix->n_subvectors > 0 && ix->d_vector % ix->n_subvectors == 0,
              "d_vector=%d must be a multiple of n_subvectors=%d", ix->d_vector, ix->n_subvectors);
  TPQ_REQUIRE(ix->n_subvectors % 4 == 0, "n_subvectors=%d must be a multiple of 4", ix->n_subvectors);
  TPQ_REQUIRE(ix->metric == TPQ_METRIC_EUCLIDEAN || ix->metric == TPQ_METRIC_COSINE,
              "unsupported distance (reference supports euclidean and cosine only)");
  TPQ_REQUIRE(ix->vq_codebook && ix->pq_codebook && ix->cell_start && ix->cell_size && ix->address2id, "index is missing reference buffers");
  TPQ_REQUIRE(ix->codes_scan || ix->n_blocks == 0, "index has no scan layout (call tpq_relayout_codes first)");
  TPQ_REQUIRE(ix->block_valid || ix->n_blocks == 0, "index has no scan layout (block_valid)");
  TPQ_REQUIRE(ix->cell_block_start && ix->pq_codebook_t && ix->pq_norm_t, "index has no scan layout (plan / codebook)");
  TPQ_REQUIRE(ix->m_pad == (ix->n_subvectors + 31) / 32 * 32, "index.m_pad is inconsistent");
  TPQ_REQUIRE(ix->capacity >= 0 && ix->capacity <= 0xFFFFFFFFll, "capacity exceeds the 32-bit address range");
  return TPQ_OK;
}

static int scan_dispatch(const tpq_index* ix, const float* x, const int64_t* cells, const int64_t* npl,
                         int nq, int n_probe, int k, int S, float* lut_ws, uint64_t* keys, cudaStream_t st,
                         const float* base_sims = nullptr, const PeerOut* peers = nullptr) {
  if (ix->residual) {
    // residual IVFPQ.  M <= 64 and d/M in {1,2,4}: the query half of the LUT is built once into a second shared-memory
    // table (two 64 KB tables -> one 16-warp CTA per SM).  Otherwise (M up to 192, any d/M) the query half is staged in
    // global memory by lut_scan_kernel and added to the per-cell half straight from L2, so one table suffices.
    const int dsub = ix->d_vector / ix->n_subvectors;
    if (!ix->part2_scan || !base_sims) { set_error("residual search needs part2_scan and base_sims"); return TPQ_ERR_BAD_ARG; }
    if (ix->metric != TPQ_METRIC_EUCLIDEAN) { set_error("residual IVFPQ is defined for the euclidean metric"); return TPQ_ERR_UNSUPPORTED; }
#define TPQ_RES(MPv, DS) launch_scan_d<MPv, 16, 1, DS, true>(ix, x, cells, npl, nq, n_probe, k, S, lut_ws, keys, st, base_sims, peers)
#ifdef TPQ_DEBUG_KNOBS
    { const char* rm = getenv("TPQ_RES_MODE");                      // A/B: staged query half + 3 light CTAs/SM for M <= 64
      if (rm && !strcmp(rm, "staged3") && ix->m_pad == 64)
        return launch_scan_d<64, 8, 3, 0, true>(ix, x, cells, npl, nq, n_probe, k, S, lut_ws, keys, st, base_sims, peers);
      if (rm && !strcmp(rm, "staged2") && ix->m_pad == 64)
        return launch_scan_d<64, 8, 2, 0, true>(ix, x, cells, npl, nq, n_probe, k, S, lut_ws, keys, st, base_sims, peers); }
#endif
    switch (fused_dsub(ix)) {
      case 1: return ix->m_pad == 32 ? TPQ_RES(32, 1) : TPQ_RES(64, 1);
      case 2: return ix->m_pad == 32 ? TPQ_RES(32, 2) : TPQ_RES(64, 2);
      case 4: return ix->m_pad == 32 ? TPQ_RES(32, 4) : TPQ_RES(64, 4);
      default: break;
    }
    switch (ix->m_pad) {
      case 32:  return TPQ_RES(32, 0);
      case 64:  return TPQ_RES(64, 0);
      case 96:  return TPQ_RES(96, 0);
      case 128: return TPQ_RES(128, 0);
      case 160: return TPQ_RES(160, 0);
      case 192: return TPQ_RES(192, 0);
      default:
        set_error("residual IVFPQ: n_subvectors=%d (padded %d) > 192 is not supported", ix->n_subvectors, ix->m_pad);
        return TPQ_ERR_UNSUPPORTED;
    }
#undef TPQ_RES
  }
#define TPQ_SCAN_ARGS ix, x, cells, npl, nq, n_probe, k, S, lut_ws, keys, st, peers
#ifdef TPQ_DEBUG_KNOBS
  // tuning knob for experiments (scripts/sweep_scan.py): TPQ_SCAN_CFG = "<warps>x<min CTAs/SM>"
  const char* cfg = getenv("TPQ_SCAN_CFG");
  if (cfg && ix->m_pad == 64) {
    if (!strcmp(cfg, "8x2"))  return launch_scan<64, 8, 2>(TPQ_SCAN_ARGS);
    if (!strcmp(cfg, "8x3"))  return launch_scan<64, 8, 3>(TPQ_SCAN_ARGS);
    if (!strcmp(cfg, "16x1")) return launch_scan<64, 16, 1>(TPQ_SCAN_ARGS);
    if (!strcmp(cfg, "4x4"))  return launch_scan<64, 4, 4>(TPQ_SCAN_ARGS);
  }
#endif
  switch (ix->m_pad) {
    case 32:  return launch_scan<32, 8, 2>(TPQ_SCAN_ARGS);
    case 64: {
      // Long probe lists: 2 CTAs/SM with a deep prefetch (127 registers).  Short ones (small cells, or one shard of a
      // cell-sharded index): per-CTA prologue/epilogue weighs more, so 3 lighter CTAs/SM overlap them better
      // (measured on C3: 1 shard 11.0 vs 11.3 ms, 1/8 shard 2.57 vs 2.11 ms).
      const double blocks_per_cta = (double)ix->n_blocks / ix->n_cells * n_probe / S;
      return blocks_per_cta >= 2000.0 ? launch_scan<64, 8, 2>(TPQ_SCAN_ARGS) : launch_scan<64, 8, 3>(TPQ_SCAN_ARGS);
    }
    case 96:  return launch_scan<96, 16, 1>(TPQ_SCAN_ARGS);      // LUT >= 128 KB: one CTA per SM, so twice the warps
    case 128: return launch_scan<128, 16, 1>(TPQ_SCAN_ARGS);
    case 160: return launch_scan<160, 16, 1>(TPQ_SCAN_ARGS);
    case 192: return launch_scan<192, 16, 1>(TPQ_SCAN_ARGS);
    default:
      set_error("n_subvectors=%d (padded %d) > 192 is not supported by the scan-layout path", ix->n_subvectors, ix->m_pad);
      return TPQ_ERR_UNSUPPORTED;
  }
}

}  // namespace tpq

using namespace tpq;

#ifdef TPQ_DEBUG_KNOBS
extern "C" int tpq_debug_set_phase_buffer(unsigned long long* dev_buf8) { g_phase = dev_buf8; return TPQ_OK; }
#endif

extern "C" int tpq_profile_enable(int on) {
  if (on && !g_prof_created) {
    for (int i = 0; i < kProfMax; ++i) { TPQ_CUDA(cudaEventCreate(&g_prof_start[i])); TPQ_CUDA(cudaEventCreate(&g_prof_stop[i])); }
    g_prof_created = true;
  }
  g_prof_on = on != 0;
  g_prof_n = 0;
  return TPQ_OK;
}

extern "C" int tpq_profile_scan_ms(float* total_ms, int* n_launches) {
  TPQ_REQUIRE(total_ms && n_launches, "null pointer");
  float tot = 0.f;
  for (int i = 0; i < g_prof_n; ++i) {
    TPQ_CUDA(cudaEventSynchronize(g_prof_stop[i]));
    float ms = 0.f;
    TPQ_CUDA(cudaEventElapsedTime(&ms, g_prof_start[i], g_prof_stop[i]));
    tot += ms;
  }
  *total_ms = tot; *n_launches = g_prof_n;
  g_prof_n = 0;
  return TPQ_OK;
}

extern "C" size_t tpq_search_workspace_bytes(const tpq_index* ix, int nq, int n_probe, int k) {
  if (!ix || nq <= 0 || n_probe <= 0 || k <= 0) return 0;
  return search_ws(ix, nq, n_probe, k).total;
}

extern "C" int tpq_ivfpq_search_cells(const tpq_index* ix, const float* x_dn, const int64_t* cells, const float* base_sims,
                                      const int64_t* n_probe_list, int nq, int n_probe, int k,
                                      float* values, int64_t* ids, int64_t* address, uint64_t* keys_out,
                                      void* ws, size_t ws_bytes, void* stream) {
  if (int rc = check_index(ix)) return rc;
  TPQ_REQUIRE(0 < k && k <= 1024, "k must be in (0, 1024], got %d", k);        // IVFPQIndex.py:473
  TPQ_REQUIRE(nq >= 0 && n_probe >= 1, "bad n_query=%d / n_probe=%d", nq, n_probe);
  TPQ_REQUIRE(x_dn && cells && n_probe_list, "null pointer argument");
  if (nq == 0) return TPQ_OK;
  SearchWs W = search_ws(ix, nq, n_probe, k);
  if (!ws || ws_bytes < W.total) { set_error("workspace too small (%zu < %zu)", ws_bytes, W.total); return TPQ_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  uint8_t* w = reinterpret_cast<uint8_t*>(ws);
  const int S = pick_slices(ix, nq, k);
  uint64_t* keys = reinterpret_cast<uint64_t*>(w + W.keys);
  if (int rc = scan_dispatch(ix, x_dn, cells, n_probe_list, nq, n_probe, k, S,
                             reinterpret_cast<float*>(w + W.lut), keys, st, base_sims)) return rc;
  return launch_merge(keys, nq, S, k, k, (int64_t)S * k, ix->address2id, ix->capacity, values, ids, address, keys_out, st);
}

extern "C" int tpq_ivfpq_search(const tpq_index* ix, const float* x_dn, int nq, int n_probe, int k,
                                int smart, float temperature,
                                float* values, int64_t* ids, int64_t* address, uint64_t* keys_out,
                                void* ws, size_t ws_bytes, void* stream) {
  if (int rc = check_index(ix)) return rc;
  TPQ_REQUIRE(0 < k && k <= 1024, "k must be in (0, 1024], got %d", k);
  TPQ_REQUIRE(nq >= 0 && x_dn, "bad query batch");
  TPQ_REQUIRE(n_probe >= 1 && n_probe <= ix->n_cells, "n_probe=%d must be in [1, n_cells=%d]", n_probe, ix->n_cells);
  if (nq == 0) return TPQ_OK;
  SearchWs W = search_ws(ix, nq, n_probe, k);
  if (!ws || ws_bytes < W.total) { set_error("workspace too small (%zu < %zu)", ws_bytes, W.total); return TPQ_ERR_WORKSPACE; }
  uint8_t* w = reinterpret_cast<uint8_t*>(ws);
  const float* x = x_dn;
  if (ix->metric == TPQ_METRIC_COSINE) {                                        // IVFPQIndex.py:474-475
    float* xn = reinterpret_cast<float*>(w + W.xnorm);
    if (int rc = tpq_normalize_columns(x_dn, ix->d_vector, nq, xn, stream)) return rc;
    x = xn;
  }
  float* probe_sims = reinterpret_cast<float*>(w + W.probe_sims);
  int64_t* cells = reinterpret_cast<int64_t*>(w + W.cells);
  int64_t* npl = reinterpret_cast<int64_t*>(w + W.npl);
  if (int rc = tpq_coarse_probe(x, ix->vq_codebook, ix->d_vector, nq, ix->n_cells, n_probe, smart, temperature,
                                probe_sims, cells, npl, w + W.coarse, W.probe_sims - W.coarse, stream)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const int S = pick_slices(ix, nq, k);
  uint64_t* keys = reinterpret_cast<uint64_t*>(w + W.keys);
  if (int rc = scan_dispatch(ix, x, cells, npl, nq, n_probe, k, S, reinterpret_cast<float*>(w + W.lut), keys, st, probe_sims)) return rc;
  return launch_merge(keys, nq, S, k, k, (int64_t)S * k, ix->address2id, ix->capacity, values, ids, address, keys_out, st);
}

extern "C" int tpq_ivfpq_scan_push(const tpq_index* ix, const float* x_dn, const int64_t* cells, const float* base_sims,
                                   const int64_t* n_probe_list, int nq, int n_probe, int k, int smart, float temperature,
                                   uint64_t* const* peer_keys, int n_peers, int my_slot,
                                   void* ws, size_t ws_bytes, void* stream) {
  if (int rc = check_index(ix)) return rc;
  TPQ_REQUIRE(0 < k && k <= 1024, "k must be in (0, 1024], got %d", k);
  TPQ_REQUIRE(nq >= 0 && n_probe >= 1 && x_dn, "bad query batch");
  TPQ_REQUIRE(peer_keys && n_peers >= 1 && n_peers <= 8 && my_slot >= 0, "tpq_ivfpq_scan_push: bad peer list");
  TPQ_REQUIRE((cells == nullptr) == (n_probe_list == nullptr), "cells and n_probe_list go together");
  if (nq == 0) return TPQ_OK;
  if (pick_slices(ix, nq, k) != 1) { set_error("tpq_ivfpq_scan_push: batch of %d queries is sliced; use the gathered exchange", nq); return TPQ_ERR_UNSUPPORTED; }
  if (nq > kQueryChunk) { set_error("tpq_ivfpq_scan_push: at most %d queries per call", kQueryChunk); return TPQ_ERR_UNSUPPORTED; }
  SearchWs W = search_ws(ix, nq, n_probe, k);
  if (!ws || ws_bytes < W.total) { set_error("workspace too small (%zu < %zu)", ws_bytes, W.total); return TPQ_ERR_WORKSPACE; }
  uint8_t* w = reinterpret_cast<uint8_t*>(ws);
  const float* x = x_dn;
  if (!cells) {                                                       // the whole coarse stage runs here (replicated probe)
    TPQ_REQUIRE(n_probe <= ix->n_cells, "n_probe=%d must be in [1, n_cells=%d]", n_probe, ix->n_cells);
    if (ix->metric == TPQ_METRIC_COSINE) {
      float* xn = reinterpret_cast<float*>(w + W.xnorm);
      if (int rc = tpq_normalize_columns(x_dn, ix->d_vector, nq, xn, stream)) return rc;
      x = xn;
    }
    float* probe_sims = reinterpret_cast<float*>(w + W.probe_sims);
    int64_t* c = reinterpret_cast<int64_t*>(w + W.cells);
    int64_t* npl = reinterpret_cast<int64_t*>(w + W.npl);
    if (int rc = tpq_coarse_probe(x, ix->vq_codebook, ix->d_vector, nq, ix->n_cells, n_probe, smart, temperature,
                                  probe_sims, c, npl, w + W.coarse, W.probe_sims - W.coarse, stream)) return rc;
    cells = c; n_probe_list = npl; base_sims = probe_sims;
  }
  PeerOut po;
  po.n = n_peers; po.slot = my_slot;
  for (int p = 0; p < n_peers; ++p) { TPQ_REQUIRE(peer_keys[p], "null peer buffer"); po.keys[p] = peer_keys[p]; }
  return scan_dispatch(ix, x, cells, n_probe_list, nq, n_probe, k, 1, reinterpret_cast<float*>(w + W.lut),
                       reinterpret_cast<uint64_t*>(w + W.keys), (cudaStream_t)stream, base_sims, &po);
}

extern "C" int tpq_merge_topk(const uint64_t* keys_in, int nq, int n_parts, int k,
                              const int64_t* address2id, int64_t capacity,
                              float* values, int64_t* ids, int64_t* address, void* stream) {
  TPQ_REQUIRE(keys_in && nq >= 0 && n_parts >= 1 && 0 < k && k <= 1024, "tpq_merge_topk: bad argument");
  TPQ_REQUIRE(!ids || address2id, "tpq_merge_topk: ids requested without address2id");
  if (nq == 0) return TPQ_OK;
  // keys_in is [n_parts, nq, k] (the all-gather's natural layout): part stride nq*k, query stride k
  return launch_merge(keys_in, nq, n_parts, k, (int64_t)nq * k, k, address2id, capacity, values, ids, address, nullptr,
                      (cudaStream_t)stream);
}
