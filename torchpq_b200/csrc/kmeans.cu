// The three small neighbours of the search path that the north-star names:
//
//   tpq_max_sim            MaxSimCuda / max_sim_tn     torchpq/kernels/MaxSimCuda.py:184-238,
//                          (MultiKMeans assignment)    kernels/cuda/max_sim.cu:182-309 (thread_nseuclidean :78-98,
//                                                      reduce_dim_2 :152-180); caller clustering/MultiKMeans.py:314-333
//   tpq_compute_centroids  ComputeCentroidsCuda        torchpq/kernels/ComputeCentroidsCuda.py:43-81,
//                          (k-means update)            kernels/cuda/compute_centroids.cu:9-86
//   tpq_pq_decode          PQDecodeCuda                torchpq/kernels/PQDecodeCuda.py:38-65, kernels/cuda/pq_decode.cu:7-53
//
// Round-1 versions: exact fp32 arithmetic in the reference's own order (parity first).  max_sim is a
// single pass with the arg-max kept in registers across centroid tiles -- no float atomicMax + racy label
// write as in the reference (max_sim.cu:173-178), so labels are deterministic (lowest index wins ties).
#include "common.cuh"

namespace tpq {

// ----------------------------------------------------------------------------- assignment: argmax_j sim(x_i, c_j)
// data [l, d, n], cent [l, d, k]; sim = sum_e -(x-c)^2 (euclidean, fmaf(-dif, dif, acc), e ascending) or sum_e x*c.
constexpr int MS_B = 128, MS_K = 8;
template <bool EUC>
__global__ void __launch_bounds__(256)
max_sim_kernel(const float* __restrict__ data, const float* __restrict__ cent, int d, int n, int k,
               float* __restrict__ maxsims, int64_t* __restrict__ labels) {
  __shared__ __align__(16) float As[MS_K][MS_B];
  __shared__ __align__(16) float Bs[MS_K][MS_B];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int l = blockIdx.y;
  const int i0 = blockIdx.x * MS_B;
  const float* X = data + (size_t)l * d * n;
  const float* Cn = cent + (size_t)l * d * k;
  float best[8]; int besti[8];
  #pragma unroll
  for (int i = 0; i < 8; ++i) { best[i] = -INFINITY; besti[i] = 0; }

  for (int j0 = 0; j0 < k; j0 += MS_B) {
    float acc[8][8];
    #pragma unroll
    for (int i = 0; i < 8; ++i)
      #pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int e0 = 0; e0 < d; e0 += MS_K) {
      #pragma unroll
      for (int r = 0; r < 4; ++r) {                         // 8 x 128 per operand / 256 threads = 4 each
        const int el = r * 256 + tid, ee = el >> 7, col = el & 127, ge = e0 + ee;
        As[ee][col] = (ge < d && i0 + col < n) ? X[(size_t)ge * n + i0 + col] : 0.f;
        Bs[ee][col] = (ge < d && j0 + col < k) ? Cn[(size_t)ge * k + j0 + col] : 0.f;
      }
      __syncthreads();
      const int kk_end = min(MS_K, d - e0);
      for (int kk = 0; kk < kk_end; ++kk) {
        const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        #pragma unroll
        for (int i = 0; i < 8; ++i)
          #pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (EUC) { const float dif = a[i] - b[j]; acc[i][j] = fmaf(-dif, dif, acc[i][j]); }
            else acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
          }
      }
      __syncthreads();
    }
    // per point: best over this thread's 8 centroids, then over the 16 threads (tx) of the row group
    #pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = -INFINITY; int vi = 0x7fffffff;
      #pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int cj = j0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
        if (cj < k && (acc[i][j] > v || (acc[i][j] == v && cj < vi))) { v = acc[i][j]; vi = cj; }
      }
      #pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, vi, o);
        if (ov > v || (ov == v && oi < vi)) { v = ov; vi = oi; }
      }
      if (v > best[i]) { best[i] = v; besti[i] = vi; }      // earlier tile (lower index) keeps ties
    }
  }
  if (tx == 0) {
    #pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int p = i0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
      if (p < n) { maxsims[(size_t)l * n + p] = best[i]; labels[(size_t)l * n + p] = besti[i]; }
    }
  }
}

// ----------------------------------------------------------------------------- centroid update
// sums[l, e, j] += data[l, e, i] for labels[l, i] == j ; counts[l, j] += 1.  One CTA = (tile of points, chunk of
// features, l); each warp owns whole features, so two warps never touch the same shared-memory row.
__global__ void __launch_bounds__(256)
centroid_accumulate_kernel(const float* __restrict__ data, const int64_t* __restrict__ labels,
                           int d, int n, int k, int de, int tile, float* __restrict__ sums, float* __restrict__ counts) {
  extern __shared__ __align__(16) float sh[];              // [de][k] sums, then [k] counts
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  const int l = blockIdx.z, e0 = blockIdx.y * de, n0 = blockIdx.x * tile;
  const int ne = min(de, d - e0), np = min(tile, n - n0);
  float* cnt = sh + (size_t)de * k;
  for (int i = tid; i < de * k + k; i += blockDim.x) sh[i] = 0.f;
  __syncthreads();
  const int64_t* lab = labels + (size_t)l * n + n0;
  for (int e = warp; e < ne; e += nw) {
    const float* x = data + ((size_t)l * d + e0 + e) * n + n0;
    float* row = sh + (size_t)e * k;
    for (int p = lane; p < np; p += 32) {
      const int j = (int)lab[p];
      if (j >= 0 && j < k) atomicAdd(row + j, x[p]);
    }
  }
  if (blockIdx.y == 0) {
    for (int p = tid; p < np; p += blockDim.x) {
      const int j = (int)lab[p];
      if (j >= 0 && j < k) atomicAdd(cnt + j, 1.f);
    }
  }
  __syncthreads();
  for (int i = tid; i < ne * k; i += blockDim.x) {
    const float v = sh[i];
    if (v != 0.f) atomicAdd(sums + ((size_t)l * d + e0 + i / k) * k + (i % k), v);
  }
  if (blockIdx.y == 0)
    for (int j = tid; j < k; j += blockDim.x) if (cnt[j] != 0.f) atomicAdd(counts + (size_t)l * k + j, cnt[j]);
}

// Streaming update for k <= 1024 (the path C5 takes): one CTA owns (l, a block of CG_E * CG_G = 16 features) for ALL
// points, walks the points in tiles of CG_T, counting-sorts every tile by label once (shared-memory permutation), and
// for each group of 4 features stages the 4 rows as one float4 per point; thread j then walks the segment of cluster
// j (and j + 256, ...) with one LDS.128 gather per member and keeps the running sums of its clusters IN REGISTERS
// across the whole point stream.  No shared-memory atomics on the data, no global atomics at all, no memset and no
// finalize pass: the CTA knows its clusters' counts and writes sum / count itself.  (Round 1's centroid_sorted_kernel
// issued one global atomicAdd per (tile, feature, cluster) -- 128 M at C5 -- and two barriers per feature row.)
constexpr int CG_T = 4096, CG_E = 4, CG_G = 4;
template <int KJ>
__global__ void __launch_bounds__(256, 2)
centroid_stream_kernel(const float* __restrict__ data, const int64_t* __restrict__ labels,
                       int d, int n, int k, float* __restrict__ cent) {
  extern __shared__ __align__(16) uint8_t sm[];
  float4* xs = reinterpret_cast<float4*>(sm);                       // [CG_T] 4 features of one point
  uint16_t* perm = reinterpret_cast<uint16_t*>(xs + CG_T);          // [CG_T] point (within tile) at sorted position
  uint16_t* lab16 = perm + CG_T;                                    // [CG_T]
  int* seg = reinterpret_cast<int*>(lab16 + CG_T);                  // [k + 1] segment starts
  int* cur = seg + (k + 1);                                         // [k] histogram / scatter cursors
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int l = blockIdx.y, e_base = blockIdx.x * (CG_E * CG_G);
  float acc[KJ][CG_G][CG_E];
  int cnt[KJ];
  #pragma unroll
  for (int a = 0; a < KJ; ++a) {
    cnt[a] = 0;
    #pragma unroll
    for (int g = 0; g < CG_G; ++g)
      #pragma unroll
      for (int e = 0; e < CG_E; ++e) acc[a][g][e] = 0.f;
  }
  const int64_t* lab_l = labels + (size_t)l * n;
  constexpr int LPT = CG_T / 256;                                    // labels per thread per tile
  int lreg[LPT];                                                     // this tile's labels, loaded one tile ahead
  auto load_labels = [&](int n0) {
    #pragma unroll
    for (int u = 0; u < LPT; ++u) {
      const int p = n0 + u * 256 + tid;
      int j = 0xFFFF;
      if (p < n) { const long long v = __ldcs(reinterpret_cast<const long long*>(lab_l + p)); j = (v >= 0 && v < k) ? (int)v : 0xFFFF; }
      lreg[u] = j;                                                   // out-of-range labels are ignored (as the reference does)
    }
  };
  load_labels(0);
  for (int n0 = 0; n0 < n; n0 += CG_T) {
    const int np = min(CG_T, n - n0);
    for (int j = tid; j < k; j += 256) cur[j] = 0;
    __syncthreads();
    #pragma unroll
    for (int u = 0; u < LPT; ++u) {
      const int p = u * 256 + tid;
      if (p < np) { lab16[p] = (uint16_t)lreg[u]; if (lreg[u] != 0xFFFF) atomicAdd(&cur[lreg[u]], 1); }
    }
    if (n0 + CG_T < n) load_labels(n0 + CG_T);                       // in flight while this tile is summed
    __syncthreads();
    if (warp == 0) {                                                 // exclusive scan of the histogram
      int carry = 0;
      for (int j0 = 0; j0 < k; j0 += 32) {
        const int j = j0 + lane;
        const int c = j < k ? cur[j] : 0;
        int incl = c;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (j < k) seg[j] = carry + incl - c;
        carry += __shfl_sync(0xffffffffu, incl, 31);
      }
      if (lane == 0) seg[k] = carry;
    }
    __syncthreads();
    #pragma unroll
    for (int a = 0; a < KJ; ++a) {
      const int j = tid + 256 * a;
      if (j < k) { cnt[a] += cur[j]; cur[j] = seg[j]; }
    }
    __syncthreads();
    for (int p = tid; p < np; p += 256) {
      const int j = lab16[p];
      if (j != 0xFFFF) perm[atomicAdd(&cur[j], 1)] = (uint16_t)p;
    }
    // (the first staging barrier below also publishes perm)
    #pragma unroll
    for (int g = 0; g < CG_G; ++g) {
      const int e0 = e_base + g * CG_E;
      if (e0 >= d) break;                                            // CTA-uniform
      const float* x0 = data + ((size_t)l * d + e0) * n + n0;
      // 4 coalesced row reads -> one float4 per point.  All 32 loads of a batch are issued before the first shared-memory
      // store (ncu on the first version: the loop body's 4 loads were the only ones in flight, 1.3 TB/s, 46 % of the
      // samples waiting on them at the STS).
      constexpr int SU = 8;
      const bool h1 = e0 + 1 < d, h2 = e0 + 2 < d, h3 = e0 + 3 < d;
      #pragma unroll 1
      for (int p0 = 0; p0 < np; p0 += 256 * SU) {
        float r[SU][4];
        #pragma unroll
        for (int u = 0; u < SU; ++u) {
          const int p = p0 + u * 256 + tid;
          const bool in = p < np;
          r[u][0] = in ? __ldcs(x0 + p) : 0.f;
          r[u][1] = (in && h1) ? __ldcs(x0 + (size_t)n + p) : 0.f;
          r[u][2] = (in && h2) ? __ldcs(x0 + 2 * (size_t)n + p) : 0.f;
          r[u][3] = (in && h3) ? __ldcs(x0 + 3 * (size_t)n + p) : 0.f;
        }
        #pragma unroll
        for (int u = 0; u < SU; ++u) {
          const int p = p0 + u * 256 + tid;
          if (p < np) xs[p] = make_float4(r[u][0], r[u][1], r[u][2], r[u][3]);
        }
      }
      __syncthreads();
      #pragma unroll
      for (int a = 0; a < KJ; ++a) {
        const int j = tid + 256 * a;
        if (j < k) {
          const int s1 = seg[j + 1];
          float4 t = make_float4(acc[a][g][0], acc[a][g][1], acc[a][g][2], acc[a][g][3]);
          for (int i = seg[j]; i < s1; ++i) {
            const float4 v = xs[perm[i]];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
          }
          acc[a][g][0] = t.x; acc[a][g][1] = t.y; acc[a][g][2] = t.z; acc[a][g][3] = t.w;
        }
      }
      __syncthreads();
    }
  }
  #pragma unroll
  for (int a = 0; a < KJ; ++a) {
    const int j = tid + 256 * a;
    if (j >= k) continue;
    const float c = (float)cnt[a];
    #pragma unroll
    for (int g = 0; g < CG_G; ++g)
      #pragma unroll
      for (int e = 0; e < CG_E; ++e) {
        const int ee = e_base + g * CG_E + e;
        if (ee < d) cent[((size_t)l * d + ee) * k + j] = cnt[a] == 0 ? 0.f : __fdiv_rn(acc[a][g][e], c);   // compute_centroids.cu:80
      }
  }
}

// centroids = count == 0 ? 0 : sum / count   (compute_centroids.cu:80)
__global__ void centroid_finalize_kernel(float* __restrict__ cent, const float* __restrict__ counts, int d, int k, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t l = i / ((size_t)d * k);
  const float c = counts[l * k + (i % k)];
  cent[i] = c == 0.f ? 0.f : __fdiv_rn(cent[i], c);
}

// ----------------------------------------------------------------------------- PQ decode
// out[m*dsub + e, i] = codebook[m, e, code[m, i]]   (pq_decode.cu:7-53).  Write-bound: 4*d bytes per vector.
__global__ void __launch_bounds__(256)
pq_decode_kernel(const float* __restrict__ codebook, const uint8_t* __restrict__ code, int dsub, int64_t n,
                 bool vec, float* __restrict__ out) {
  extern __shared__ __align__(16) float tab[];             // [dsub][256]
  const int m = blockIdx.y;
  for (int i = threadIdx.x; i < dsub * 256; i += blockDim.x) tab[i] = codebook[(size_t)m * dsub * 256 + i];
  __syncthreads();
  const uint8_t* cm = code + (size_t)m * n;
  float* om = out + (size_t)m * dsub * n;
  for (int64_t i4 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i4 < n; i4 += (int64_t)gridDim.x * blockDim.x * 4) {
    if (vec) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(cm + i4);
      const int c0 = w & 0xff, c1 = (w >> 8) & 0xff, c2 = (w >> 16) & 0xff, c3 = w >> 24;
      for (int e = 0; e < dsub; ++e) {
        const float* t = tab + e * 256;
        __stcs(reinterpret_cast<float4*>(om + (size_t)e * n + i4), make_float4(t[c0], t[c1], t[c2], t[c3]));
      }
    } else {
      for (int u = 0; u < 4 && i4 + u < n; ++u) {
        const int c = cm[i4 + u];
        for (int e = 0; e < dsub; ++e) om[(size_t)e * n + i4 + u] = tab[e * 256 + c];
      }
    }
  }
}

}  // namespace tpq

namespace tpq {
bool assign_tc_supported(int l, int d, long long n, int k, const void* data, const void* cent);
int launch_assign_tc(const float* data, const float* cent, int l, int d, long long n, int k, int exact_values,
                     float* maxsims, int64_t* labels, cudaStream_t st);
}

using namespace tpq;

extern "C" int tpq_max_sim(const float* data, const float* centroids, int l, int d, int64_t n, int k, int metric,
                           int exact, float* maxsims, int64_t* labels, void* stream) {
  TPQ_REQUIRE(data && centroids && maxsims && labels, "tpq_max_sim: null pointer");
  TPQ_REQUIRE(l > 0 && d > 0 && n >= 0 && k > 0, "tpq_max_sim: bad sizes l=%d d=%d n=%lld k=%d", l, d, (long long)n, k);
  TPQ_REQUIRE(n <= 0x7fffffffll - MS_B && l <= 65535, "tpq_max_sim: n or l too large");
  TPQ_REQUIRE(metric == TPQ_METRIC_EUCLIDEAN || metric == TPQ_METRIC_COSINE, "tpq_max_sim: unsupported metric %d", metric);
  if (n == 0) return TPQ_OK;
  if (exact != 1 && metric == TPQ_METRIC_EUCLIDEAN && assign_tc_supported(l, d, n, k, data, centroids))
    return launch_assign_tc(data, centroids, l, d, n, k, exact == 2 ? 1 : 0, maxsims, labels, (cudaStream_t)stream);
  dim3 grid((unsigned)((n + MS_B - 1) / MS_B), l);
  if (metric == TPQ_METRIC_EUCLIDEAN)
    max_sim_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(data, centroids, d, (int)n, k, maxsims, labels);
  else
    max_sim_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(data, centroids, d, (int)n, k, maxsims, labels);
  TPQ_LAUNCH_CHECK("max_sim_kernel");
  return TPQ_OK;
}

extern "C" size_t tpq_compute_centroids_workspace_bytes(int l, int k) { return align_up((size_t)l * k * 4, 256); }

extern "C" int tpq_compute_centroids(const float* data, const int64_t* labels, int l, int d, int64_t n, int k,
                                     float* centroids, void* ws, size_t ws_bytes, void* stream) {
  TPQ_REQUIRE(data && labels && centroids, "tpq_compute_centroids: null pointer");
  TPQ_REQUIRE(l > 0 && d > 0 && n >= 0 && k > 0 && n <= 0x7fffffffll && l <= 65535, "tpq_compute_centroids: bad sizes");
  if (!ws || ws_bytes < tpq_compute_centroids_workspace_bytes(l, k)) {
    set_error("tpq_compute_centroids: workspace too small"); return TPQ_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* counts = reinterpret_cast<float*>(ws);
  if (n > 0 && k <= 1024) {
    // streaming kernel: sums in registers, one launch, deterministic ownership of every output element
    const size_t smem = (size_t)CG_T * 16 + (size_t)CG_T * 2 * 2 + (size_t)(2 * k + 1) * 4;
    dim3 grid((unsigned)((d + CG_E * CG_G - 1) / (CG_E * CG_G)), l);
#define TPQ_CSTREAM(KJ)                                                                                              \
    do {                                                                                                             \
      TPQ_CUDA(cudaFuncSetAttribute(centroid_stream_kernel<KJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      centroid_stream_kernel<KJ><<<grid, 256, smem, st>>>(data, labels, d, (int)n, k, centroids);                    \
    } while (0)
    if (k <= 256) TPQ_CSTREAM(1); else if (k <= 512) TPQ_CSTREAM(2); else TPQ_CSTREAM(4);
#undef TPQ_CSTREAM
    TPQ_LAUNCH_CHECK("centroid_stream_kernel");
    return TPQ_OK;
  }
  TPQ_CUDA(cudaMemsetAsync(centroids, 0, (size_t)l * d * k * 4, st));
  TPQ_CUDA(cudaMemsetAsync(counts, 0, (size_t)l * k * 4, st));
  if (n > 0) {
    int de = (int)((160 * 1024 / 4 - k) / k);               // features per CTA so that (de+1)*k floats fit 160 KB
    if (de < 1) { set_error("tpq_compute_centroids: k=%d too large for one shared-memory row set", k); return TPQ_ERR_UNSUPPORTED; }
    if (de > d) de = d;
    const int tile = 8192;
    size_t smem = (size_t)(de + 1) * k * 4;
    TPQ_CUDA(cudaFuncSetAttribute(centroid_accumulate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)((n + tile - 1) / tile), (d + de - 1) / de, l);
    centroid_accumulate_kernel<<<grid, 256, smem, st>>>(data, labels, d, (int)n, k, de, tile, centroids, counts);
    TPQ_LAUNCH_CHECK("centroid_accumulate_kernel");
  }
  const size_t total = (size_t)l * d * k;
  centroid_finalize_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(centroids, counts, d, k, total);
  TPQ_LAUNCH_CHECK("centroid_finalize_kernel");
  return TPQ_OK;
}

extern "C" int tpq_pq_decode(const float* codebook, const uint8_t* code, int M, int dsub, int64_t n, float* out, void* stream) {
  TPQ_REQUIRE(codebook && code && out, "tpq_pq_decode: null pointer");
  TPQ_REQUIRE(M > 0 && dsub > 0 && n >= 0 && M <= 65535, "tpq_pq_decode: bad sizes");
  TPQ_REQUIRE((size_t)dsub * 1024 <= 200 * 1024, "tpq_pq_decode: d_subvector=%d too large", dsub);
  if (n == 0) return TPQ_OK;
  size_t smem = (size_t)dsub * 1024;
  TPQ_CUDA(cudaFuncSetAttribute(pq_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t bx = (n + 1023) / 1024;
  if (bx > 148 * 8) bx = 148 * 8;
  const bool vec = (n & 3) == 0 && (reinterpret_cast<uintptr_t>(code) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  pq_decode_kernel<<<dim3((unsigned)bx, M), 256, smem, (cudaStream_t)stream>>>(codebook, code, dsub, n, vec, out);
  TPQ_LAUNCH_CHECK("pq_decode_kernel");
  return TPQ_OK;
}
