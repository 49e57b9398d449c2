// ADC look-up table in the reference's `precomputed` layout [M, nq, 256] fp32.
//
// Replaces PQCodec.precompute_adc (torchpq/codec/PQCodec.py:62-75) ->
// MultiKMeans.sim (torchpq/clustering/MultiKMeans.py:211-223):
//   euclidean: euc_sim   LUT[m,q,c] = ((2 * <x_m, p_mc>) - |x_m|^2) - |p_mc|^2   (MultiKMeans.py:183-209)
//   cosine   : cos_sim(normalize=False) = <x_m, p_mc>                            (MultiKMeans.py:155-181,220-221)
// The reference does this with a batched cuBLAS sgemm (K = d/M, i.e. 2..16) plus four
// elementwise passes.  Here: one launch, exact fp32 FMA in ascending-i order.
// (The tuned search path does not use this table at all -- it builds the LUT inside
// the scan CTA's shared memory, see scan.cu; this entry point exists for the
// reference-layout drop-in tpq_ivfpq_topk and for parity tests of the LUT itself.)
#include "common.cuh"

namespace tpq {

constexpr int LUT_QT = 16;   // queries per CTA

__global__ void __launch_bounds__(256)
build_lut_kernel(const float* __restrict__ x, const float* __restrict__ cb,
                 int d, int M, int nq, int metric, float* __restrict__ lut) {
  const int m = blockIdx.x;
  const int q0 = blockIdx.y * LUT_QT;
  const int c = threadIdx.x;                 // code
  const int dsub = d / M;
  const float* p = cb + (size_t)m * dsub * 256 + c;     // p[i * 256]
  float b2 = 0.f;
  if (metric == TPQ_METRIC_EUCLIDEAN)
    for (int i = 0; i < dsub; ++i) { float v = p[i * 256]; b2 = __fadd_rn(b2, __fmul_rn(v, v)); }
  for (int qq = 0; qq < LUT_QT; ++qq) {
    const int q = q0 + qq;
    if (q >= nq) break;
    const float* xq = x + (size_t)m * dsub * nq + q;    // xq[i * nq]
    float dot = 0.f, a2 = 0.f;
    for (int i = 0; i < dsub; ++i) {
      float xv = __ldg(xq + (size_t)i * nq);
      dot = fmaf(xv, p[i * 256], dot);
      a2 = __fadd_rn(a2, __fmul_rn(xv, xv));
    }
    float y = dot;
    if (metric == TPQ_METRIC_EUCLIDEAN) y = __fsub_rn(__fsub_rn(__fmul_rn(dot, 2.f), a2), b2);
    lut[((size_t)m * nq + q) * 256 + c] = y;
  }
}

}  // namespace tpq

using namespace tpq;

extern "C" int tpq_build_lut(const float* x_dn, const float* pq_codebook, int d, int M, int nq, int metric,
                             float* lut_mqk, void* stream) {
  TPQ_REQUIRE(x_dn && pq_codebook && lut_mqk, "tpq_build_lut: null pointer");
  TPQ_REQUIRE(d > 0 && M > 0 && d % M == 0, "d_vector=%d must be a multiple of n_subvectors=%d", d, M);
  TPQ_REQUIRE(metric == TPQ_METRIC_EUCLIDEAN || metric == TPQ_METRIC_COSINE, "unsupported metric %d", metric);
  TPQ_REQUIRE(nq >= 0, "bad n_query");
  if (nq == 0) return TPQ_OK;
  dim3 grid(M, (nq + LUT_QT - 1) / LUT_QT);
  build_lut_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x_dn, pq_codebook, d, M, nq, metric, lut_mqk);
  TPQ_LAUNCH_CHECK("build_lut_kernel");
  return TPQ_OK;
}
