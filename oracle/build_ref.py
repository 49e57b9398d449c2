"""Compile the REFERENCE's own CUDA kernels, unmodified, into oracle/_ref/ (test infrastructure).

Runs only where /root/reference exists (the build container); the GPU box uses the prebuilt
shared objects that travel with the snapshot.  Nothing is copied into the repository: the
reference source is read where it lies, specialised exactly the way the reference's Python
wrapper specialises it -- textual token substitution, torchpq/kernels/IVFPQTopkCuda.py:32-42 --
written to a scratch file under /tmp together with a small host launcher of ours, and compiled
with the reference's own NVRTC options (--maxrregcount=255 --use_fast_math,
IVFPQTopkCuda.py:44-53) for sm_100a.  Only the resulting .so files land in oracle/_ref/
(git-ignored, NOT gpurun-ignored).

The launcher reproduces IVFPQTopkCuda.topk's launch (grid = n_query, block = tpb,
dynamic smem = M * 1024; IVFPQTopkCuda.py:121-141).
"""
import os
import subprocess
import sys
import tempfile

REF = "/root/reference/torchpq/kernels/cuda"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

LAUNCHER = r'''
#include <cuda_runtime.h>
extern "C" int ref_ivfpq_topk_launch(const void* data, const float* precomputed, const unsigned char* is_empty,
                                     const long long* cell_start, const long long* cell_size,
                                     const long long* tot_size, const long long* n_probe_list,
                                     float* values, long long* indices,
                                     int n_data, int n_query, int n_probe, int n_cand_pow2, void* stream) {
  cudaError_t e = cudaFuncSetAttribute(ivfpq_topk, cudaFuncAttributeMaxDynamicSharedMemorySize, %(SMEM)d);
  if (e != cudaSuccess) return (int)e;
  ivfpq_topk<<<n_query, %(TPB)d, %(SMEM)d, (cudaStream_t)stream>>>(
      (const uint8n_t*)data, precomputed, is_empty, cell_start, cell_size, tot_size, n_probe_list,
      values, indices, n_data, n_query, n_probe, n_cand_pow2);
  return (int)cudaGetLastError();
}
// launch of IVFPQTopkCuda.topk_residual_precomputed (IVFPQTopkCuda.py:212-283)
extern "C" int ref_ivfpq_topk_residual_precomputed_launch(
    const void* data, const float* part1, const float* part2, const long long* cells, const float* base_sims,
    const unsigned char* is_empty, const long long* cell_start, const long long* cell_size, const long long* tot_size,
    const long long* n_probe_list, float* values, long long* indices,
    int n_data, int n_query, int n_probe, int n_cand_pow2, void* stream) {
  cudaError_t e = cudaFuncSetAttribute(ivfpq_topk_residual_precomputed, cudaFuncAttributeMaxDynamicSharedMemorySize, %(SMEM)d);
  if (e != cudaSuccess) return (int)e;
  ivfpq_topk_residual_precomputed<<<n_query, %(TPB)d, %(SMEM)d, (cudaStream_t)stream>>>(
      (const uint8n_t*)data, part1, part2, cells, base_sims, is_empty, cell_start, cell_size, tot_size, n_probe_list,
      values, indices, n_data, n_query, n_probe, n_cand_pow2);
  return (int)cudaGetLastError();
}
'''


def build_ivfpq_topk(m, tpb=256, n_cs=4, stack_capacity=2):
    src = open(os.path.join(REF, "ivfpq_topk.cu")).read()
    varnames = ", ".join(f"d{i}" for i in range(n_cs))
    code = (src.replace("_VARNAMES_", varnames).replace("_M_", str(m)).replace("_K_", "256")
            .replace("_TPB_", str(tpb)).replace("_NCS_", str(n_cs)).replace("_STACKCAP_", str(stack_capacity)))
    code += LAUNCHER % {"SMEM": m * 1024, "TPB": tpb}
    os.makedirs(OUT, exist_ok=True)
    out = os.path.join(OUT, f"libref_ivfpq_topk_m{m}_tpb{tpb}.so")
    with tempfile.TemporaryDirectory() as td:
        cu = os.path.join(td, f"ref_ivfpq_topk_m{m}.cu")
        open(cu, "w").write(code)
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-maxrregcount=255", "-use_fast_math",
               "-shared", "-Xcompiler", "-fPIC", "-w", "-o", out, cu]
        subprocess.check_call(cmd)
    return out


def _compile(name, code, extra=()):
    os.makedirs(OUT, exist_ok=True)
    out = os.path.join(OUT, f"lib{name}.so")
    with tempfile.TemporaryDirectory() as td:
        cu = os.path.join(td, f"{name}.cu")
        open(cu, "w").write(code)
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", *extra,
                               "-shared", "-Xcompiler", "-fPIC", "-w", "-o", out, cu])
    return out


MAX_SIM_LAUNCHER = r'''
#include <cuda_runtime.h>
extern "C" int ref_max_sim_tn_launch(const float* A, const float* B, float* C, long long* D,
                                     int M, int N, int K, int DIM, int L, void* stream) {
  dim3 grid = DIM == 1 ? dim3((N + 127) / 128, (M + 127) / 128, L) : dim3((M + 127) / 128, (N + 127) / 128, L);
  max_sim_tn<<<grid, 256, 0, (cudaStream_t)stream>>>(A, B, C, D, M, N, K, DIM);
  return (int)cudaGetLastError();
}
'''


def build_max_sim(distfn="thread_nseuclidean", tag="euclidean"):
    """MaxSimCuda.__init__ (kernels/MaxSimCuda.py:17-68): m/n/k/dim left dynamic, options --maxrregcount=128
    --use_fast_math; launch as _call_tn (:184-238)."""
    src = open(os.path.join(REF, "max_sim.cu")).read()
    code = (src.replace("_M_", "M").replace("_N_", "N").replace("_K_", "K").replace("_DIM_", "DIM")
            .replace("_DISTFN_", distfn))
    return _compile(f"ref_max_sim_{tag}", code + MAX_SIM_LAUNCHER, ("-maxrregcount=128", "-use_fast_math"))


CENTROIDS_LAUNCHER = r'''
#include <cuda_runtime.h>
#undef int64_t
extern "C" int ref_compute_centroids_launch(const float* data, const long long* labels, float* C,
                                            int m, int n, int e, int k, void* stream) {
  dim3 grid(m, (k + %(DK)d - 1) / %(DK)d, (e + %(DE)d - 1) / %(DE)d);
  int smem = %(DK)d * (%(DE)d + 1) * 4;
  cudaFuncSetAttribute(compute_centroids, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  compute_centroids<<<grid, 256, smem, (cudaStream_t)stream>>>(data, labels, C, m, n, e, k);
  return (int)cudaGetLastError();
}
'''


def build_compute_centroids(dk, de=1, tpb=256):
    """ComputeCentroidsCuda.__init__/__call__ (kernels/ComputeCentroidsCuda.py:23-81); de=1, dk=min(k,4096)
    as MultiKMeans constructs it (clustering/MultiKMeans.py:67-79)."""
    import math
    src = open(os.path.join(REF, "compute_centroids.cu")).read()
    code = (src.replace("_DE_", str(de)).replace("_DK_", str(dk)).replace("_TPB_", str(tpb))
            .replace("_NITERS_", str(math.ceil(dk / tpb))))
    # NVRTC (what the reference uses) has no <stdint.h>, so the source typedefs int64_t itself; under nvcc
    # that clashes with the system typedef.  A one-line preprocessor prelude renames the source's own
    # typedef -- the reference text itself is untouched.
    prelude = "#include <cuda_runtime.h>\n#define int64_t ref_int64_t\n"
    return _compile(f"ref_compute_centroids_dk{dk}", prelude + code + CENTROIDS_LAUNCHER % {"DK": dk, "DE": de})


DECODE_LAUNCHER = r'''
#include <cuda_runtime.h>
extern "C" int ref_pq_decode_launch(const float* codebook, const unsigned char* code, float* result,
                                    int M, int D, int N, void* stream) {
  dim3 grid((M + 1) / 2, (D + 7) / 8);
  int smem = 8 * 2 * 256 * 4;
  pq_decode<<<grid, 256, smem, (cudaStream_t)stream>>>(codebook, code, result, M, D, N);
  return (int)cudaGetLastError();
}
'''


def build_pq_decode(tm=2, td=8, tpb=256):
    """PQDecodeCuda (kernels/PQDecodeCuda.py:8-65; tm=2, td=8 as codec/PQCodec.py:34 constructs it)."""
    src = open(os.path.join(REF, "pq_decode.cu")).read()
    code = src.replace("_TD_", str(td)).replace("_TM_", str(tm)).replace("_TPB_", str(tpb))
    return _compile("ref_pq_decode", code + DECODE_LAUNCHER)


PLACE_LAUNCHER = r'''
#include <cuda_runtime.h>
// GetIOACuda.__call__ (kernels/GetIOACuda.py:34-60): grid = ceil(n_unique / tpb); GetWriteAddressV2Cuda.__call__
// (kernels/GetWriteAddressV2Cuda.py:33-66): grid = ceil(n_labels / tpb)
extern "C" int ref_get_ioa_launch(const long long* labels, const long long* unique_labels, long long* ioa,
                                  int n_labels, int n_unique, void* stream) {
  get_ioa<<<(n_unique + %(TPB)d - 1) / %(TPB)d, %(TPB)d, 0, (cudaStream_t)stream>>>(labels, unique_labels, ioa, n_labels, n_unique);
  return (int)cudaGetLastError();
}
extern "C" int ref_get_write_address_launch(const unsigned char* is_empty, const long long* div_start, const long long* div_size,
                                            const long long* labels, const long long* ioa, long long* write_adr,
                                            int n_slots, int n_labels, void* stream) {
  get_write_address<<<(n_labels + %(TPB)d - 1) / %(TPB)d, %(TPB)d, 0, (cudaStream_t)stream>>>(
      is_empty, div_start, div_size, labels, ioa, write_adr, n_slots, n_labels);
  return (int)cudaGetLastError();
}
'''


def build_placement(tpb=256):
    """get_ioa.cu + get_write_address_v2.cu (container/CellContainer.py:89-99 constructs both with tpb=256).  The two
    sources repeat the same typedefs, which nvcc accepts when they agree."""
    a = open(os.path.join(REF, "get_ioa.cu")).read().replace("_TPB_", str(tpb))
    b = open(os.path.join(REF, "get_write_address_v2.cu")).read().replace("_TPB_", str(tpb))
    return _compile("ref_placement", a + "\n" + b + PLACE_LAUNCHER % {"TPB": tpb})


def main():
    if not os.path.isdir(REF):
        print("oracle/_ref: /root/reference not present; keeping prebuilt files")
        return 0
    outs = [build_ivfpq_topk(m) for m in (8, 16, 32, 64, 120)]
    outs += [build_max_sim("thread_nseuclidean", "euclidean"), build_max_sim("thread_matmul", "inner")]
    outs += [build_compute_centroids(dk) for dk in (16, 256)]
    outs += [build_pq_decode()]
    outs += [build_placement()]
    for out in outs:
        print("built", os.path.relpath(out, HERE))
    return 0


if __name__ == "__main__":
    sys.exit(main())
