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


def main():
    if not os.path.isdir(REF):
        print("oracle/_ref: /root/reference not present; keeping prebuilt files")
        return 0
    for m in (8, 16, 32, 64, 120):
        out = build_ivfpq_topk(m)
        print("built", os.path.relpath(out, HERE))
    return 0


if __name__ == "__main__":
    sys.exit(main())
