// Shared device/host helpers for the tpq_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/tpq_b200.h"

namespace tpq {

// ----------------------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
int  cuda_fail(cudaError_t e, const char* what);

#define TPQ_REQUIRE(cond, ...)                                   \
  do { if (!(cond)) { ::tpq::set_error(__VA_ARGS__); return TPQ_ERR_BAD_ARG; } } while (0)
#define TPQ_CUDA(call)                                           \
  do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return ::tpq::cuda_fail(e__, #call); } while (0)
#define TPQ_LAUNCH_CHECK(name)                                   \
  do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return ::tpq::cuda_fail(e__, name); } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int next_pow2(int x) { int p = 1; while (p < x) p <<= 1; return p; }

// ----------------------------------------------------------------------------- candidate keys
// A candidate is one 64-bit key: high word = order-preserving image of the fp32 score,
// low word = ~address.  Larger key == better candidate under the total order
// (score descending, address ascending).  Key 0 is "empty" (below every real key).
__host__ __device__ __forceinline__ uint32_t score_to_ord(float f) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f);
#else
  uint32_t b; memcpy(&b, &f, 4);
#endif
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord_to_score(uint32_t o) {
  uint32_t b = (o & 0x80000000u) ? (o ^ 0x80000000u) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  float f; memcpy(&f, &b, 4); return f;
#endif
}
__device__ __forceinline__ uint64_t make_key(float score, uint32_t addr) {
  return ((uint64_t)score_to_ord(score) << 32) | (uint64_t)(0xFFFFFFFFu - addr);
}
__device__ __forceinline__ float    key_score(uint64_t k) { return ord_to_score((uint32_t)(k >> 32)); }
__device__ __forceinline__ uint32_t key_addr(uint64_t k)  { return 0xFFFFFFFFu - (uint32_t)k; }

#ifdef __CUDACC__
// ----------------------------------------------------------------------------- warp-cooperative sorted lists in shared memory
// All routines are called by a full warp with warp-uniform arguments.

// Bitonic sort of a[0..n) (n a power of two >= 2), DESCENDING.
__device__ __forceinline__ void warp_sort_desc(uint64_t* a, int n, int lane) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < (n >> 1); t += 32) {
        int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int hi = lo + j;
        bool desc = ((lo & k) == 0);
        uint64_t x = a[lo], y = a[hi];
        if ((x < y) == desc) { a[lo] = y; a[hi] = x; }
      }
      __syncwarp();
    }
  }
}

// Sort a BITONIC sequence a[0..n) descending (the merge half of the network).
__device__ __forceinline__ void warp_bitonic_merge_desc(uint64_t* a, int n, int lane) {
  for (int j = n >> 1; j > 0; j >>= 1) {
    for (int t = lane; t < (n >> 1); t += 32) {
      int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
      int hi = lo + j;
      uint64_t x = a[lo], y = a[hi];
      if (x < y) { a[lo] = y; a[hi] = x; }
    }
    __syncwarp();
  }
}

// A[0..n) and B[0..nb) both sorted descending, nb <= n, n a power of two.
// A <- the n largest of A u B, sorted descending.  B is left unspecified.
__device__ __forceinline__ void warp_merge_desc(uint64_t* A, int n, const uint64_t* B, int nb, int lane) {
  for (int t = lane; t < nb; t += 32) {
    int i = n - nb + t;
    uint64_t x = A[i], y = B[nb - 1 - t];
    A[i] = x > y ? x : y;
  }
  __syncwarp();
  warp_bitonic_merge_desc(A, n, lane);
}

// Per-warp top-k selector: sorted list of KP keys + an unsorted staging buffer.
// Candidates are appended with push(); when the buffer cannot take another full
// warp's worth it is sorted and merged into the list.
constexpr int kTopkBuf = 64;
struct WarpTopK {
  uint64_t* list;   // [kp] descending
  uint64_t* buf;    // [kTopkBuf]
  int kp;           // power of two >= max(k, 32)... list capacity
  int k;            // wanted
  int cnt;          // entries in buf (warp-uniform)

  __device__ __forceinline__ void init(uint64_t* list_, uint64_t* buf_, int kp_, int k_, int lane) {
    list = list_; buf = buf_; kp = kp_; k = k_; cnt = 0;
    for (int i = lane; i < kp; i += 32) list[i] = 0;
    __syncwarp();
  }
  // k-th best so far (0 while fewer than k candidates were seen)
  __device__ __forceinline__ uint64_t kth() const { return list[k - 1]; }

  __device__ __forceinline__ void flush(int lane) {
    if (cnt == 0) return;
    for (int i = cnt + lane; i < kTopkBuf; i += 32) buf[i] = 0;
    __syncwarp();
    int nb = cnt <= 32 ? 32 : kTopkBuf;
    warp_sort_desc(buf, nb, lane);
    if (nb > kp) nb = kp;   // kp >= 32 always; only guards kp == 32 < 64
    warp_merge_desc(list, kp, buf, nb, lane);
    cnt = 0;
  }
  // pass: this lane has a candidate.  Returns true if a flush happened (threshold may have risen).
  __device__ __forceinline__ bool push(bool pass, uint64_t key, int lane) {
    unsigned m = __ballot_sync(0xffffffffu, pass);
    if (m == 0) return false;
    if (pass) buf[cnt + __popc(m & ((1u << lane) - 1))] = key;
    cnt += __popc(m);
    __syncwarp();
    if (cnt > kTopkBuf - 32) { flush(lane); return true; }
    return false;
  }
};

// Bitonic sort of a[0..n) (n a power of two), DESCENDING, by the whole CTA (every thread must call it).
__device__ __forceinline__ void cta_sort_desc(uint64_t* a, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
        int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        int hi = lo + j;
        bool desc = ((lo & k) == 0);
        uint64_t x = a[lo], y = a[hi];
        if ((x < y) == desc) { a[lo] = y; a[hi] = x; }
      }
      __syncthreads();
    }
  }
}

// CTA-wide top-k selector: ONE sorted list of kp keys shared by all warps of the CTA (so its
// k-th entry is the true running k-th best of everything merged so far), one staging buffer
// per warp.  A warp whose buffer is more than half full sorts it privately, then takes the
// CTA lock and merges it into the shared list; the other warps keep scanning against the
// (slightly stale, always valid) threshold in *thr.  Exact: a candidate is only ever dropped
// when k better ones are already in the list.
struct CtaTopK {
  uint64_t* list;               // [kp] descending, shared by the CTA
  unsigned long long* thr;      // k-th best of `list` (0 while fewer than k merged)
  int* lock;
  uint64_t* buf;                // [kTopkBuf] this warp's staging buffer
  int kp, k, cnt;

  // call from every thread of the CTA, then __syncthreads()
  __device__ __forceinline__ void init(uint64_t* list_, unsigned long long* thr_, int* lock_, uint64_t* buf_,
                                       int kp_, int k_) {
    list = list_; thr = thr_; lock = lock_; buf = buf_; kp = kp_; k = k_; cnt = 0;
    for (int i = threadIdx.x; i < kp; i += blockDim.x) list[i] = 0;
    if (threadIdx.x == 0) { *thr = 0ull; *lock = 0; }
  }
  __device__ __forceinline__ uint64_t threshold() const { return *reinterpret_cast<volatile unsigned long long*>(thr); }

  __device__ __forceinline__ void flush(int lane) {
    if (cnt == 0) return;
    for (int i = cnt + lane; i < kTopkBuf; i += 32) buf[i] = 0;
    __syncwarp();
    int nb = cnt <= 32 ? 32 : kTopkBuf;
    warp_sort_desc(buf, nb, lane);
    if (nb > kp) nb = kp;
    if (lane == 0) { while (atomicCAS(lock, 0, 1) != 0) __nanosleep(20); }
    __syncwarp();
    __threadfence_block();
    if (buf[0] > list[k - 1]) {                 // warp-uniform: anything left that still beats the k-th best?
      warp_merge_desc(list, kp, buf, nb, lane);
      if (lane == 0) *reinterpret_cast<volatile unsigned long long*>(thr) = list[k - 1];
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) atomicExch(lock, 0);
    cnt = 0;
  }
  __device__ __forceinline__ void push(bool pass, uint64_t key, int lane) {
    unsigned m = __ballot_sync(0xffffffffu, pass);
    if (m == 0) return;
    if (pass) buf[cnt + __popc(m & ((1u << lane) - 1))] = key;
    cnt += __popc(m);
    __syncwarp();
    if (cnt > kTopkBuf - 32) flush(lane);
  }
  // Whole-CTA flush (every thread calls it): the staging buffers of all warps (`bufs`, nw * kTopkBuf entries,
  // contiguous) are compacted, sorted together by the CTA and merged into the list once, instead of nw
  // lock-serialised warp flushes.  Used to bootstrap the threshold and to drain the buffers at the end.  Only the
  // smallest power of two that holds the live entries is sorted.  `scratch` is an int[nw + 1] in shared memory.
  __device__ __forceinline__ void cta_flush(uint64_t* bufs, int* scratch, int nw, int lane, int warp) {
    // compaction: every thread picks its (up to two) entries up into registers, then writes them at the warp's offset
    const uint64_t e0 = lane < cnt ? buf[lane] : 0ull;
    const uint64_t e1 = lane + 32 < cnt ? buf[lane + 32] : 0ull;
    if (lane == 0) scratch[warp + 1] = cnt;
    if (threadIdx.x == 0) scratch[0] = 0;
    __syncthreads();
    if (warp == 0) {                                                 // inclusive scan of the nw counts (nw <= 32)
      int v = lane < nw ? scratch[lane + 1] : 0;
      #pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
      if (lane < nw) scratch[lane + 1] = v;
    }
    __syncthreads();
    const int total = scratch[nw], off0 = scratch[warp];
    int nb = 32;
    while (nb < total) nb <<= 1;                                     // <= nw * kTopkBuf by construction
    if (lane < cnt) bufs[off0 + lane] = e0;
    if (lane + 32 < cnt) bufs[off0 + lane + 32] = e1;
    __syncthreads();
    for (int i = total + threadIdx.x; i < nb; i += blockDim.x) bufs[i] = 0;
    __syncthreads();
    if (total > 0) {                                                 // CTA-uniform
      cta_sort_desc(bufs, nb);
      if (warp == 0) {
        const int take = nb < kp ? nb : kp;
        warp_merge_desc(list, kp, bufs, take, lane);
        if (lane == 0) *reinterpret_cast<volatile unsigned long long*>(thr) = list[k - 1];
      }
    }
    cnt = 0;
    __syncthreads();
  }
};
#endif  // __CUDACC__

}  // namespace tpq
