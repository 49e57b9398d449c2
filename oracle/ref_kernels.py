"""Launch the reference's own CUDA kernels (prebuilt into oracle/_ref/ by build_ref.py) on torch tensors.

TEST INFRASTRUCTURE: used by tests/ to pin the CPU oracle (and the product) against the real
reference kernel on the GPU box, and by bench.py to report the reference kernel's own time on
B200 beside ours.  Host logic restates IVFPQTopkCuda.topk (torchpq/kernels/IVFPQTopkCuda.py:81-142):
tot_size, -inf / 0 initialised outputs padded to 2*nextpow2(ceil(k/2)), slice [:, :k].
"""
import ctypes as C
import math
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_libs = {}


def lib_path(m, tpb=256):
    return os.path.join(_REF, f"libref_ivfpq_topk_m{m}_tpb{tpb}.so")


def available(m, tpb=256) -> bool:
    return os.path.exists(lib_path(m, tpb))


def _load(m, tpb=256):
    key = (m, tpb)
    if key not in _libs:
        lib = C.CDLL(lib_path(m, tpb))
        lib.ref_ivfpq_topk_launch.restype = C.c_int
        lib.ref_ivfpq_topk_launch.argtypes = [C.c_void_p] * 9 + [C.c_int] * 4 + [C.c_void_p]
        _libs[key] = lib
    return _libs[key]


def ivfpq_topk(data, precomputed, cell_start, cell_size, is_empty, n_probe_list, k, tpb=256):
    """Reference kernel `ivfpq_topk` (kernels/cuda/ivfpq_topk.cu:822-971), 1 < k <= tpb.
    Returns (values [nq,k] f32, address [nq,k] i64) exactly as IVFPQTopkCuda.topk does."""
    m = data.shape[0] * 4
    lib = _load(m, tpb)
    n_data = data.shape[1]
    n_query, n_probe = cell_start.shape
    assert precomputed.shape == (m, n_query, 256) and 1 < k <= tpb
    n_pow2 = 2 * (1 if math.ceil(k / 2) == 0 else 2 ** math.ceil(math.log2(math.ceil(k / 2))))
    dev = data.device
    tot_size = cell_size.sum(dim=1)
    values = torch.empty(n_query, n_pow2, device=dev, dtype=torch.float32).fill_(float("-inf"))
    indices = torch.zeros(n_query, n_pow2, device=dev, dtype=torch.int64)
    args = [t.contiguous() for t in (data, precomputed, is_empty, cell_start, cell_size, tot_size, n_probe_list)]
    rc = lib.ref_ivfpq_topk_launch(*[C.c_void_p(t.data_ptr()) for t in args],
                                   C.c_void_p(values.data_ptr()), C.c_void_p(indices.data_ptr()),
                                   n_data, n_query, n_probe, n_pow2,
                                   C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc != 0:
        raise RuntimeError(f"reference kernel launch failed: cudaError {rc}")
    return values[:, :k], indices[:, :k]


# ---------------------------------------------------------------------------------- k-means / codec kernels
def _so(name):
    return os.path.join(_REF, f"lib{name}.so")


def aux_available() -> bool:
    return all(os.path.exists(_so(n)) for n in ("ref_max_sim_euclidean", "ref_max_sim_inner", "ref_pq_decode",
                                                "ref_compute_centroids_dk256", "ref_compute_centroids_dk16"))


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def max_sim_tn(A, B, distance="euclidean", dim=2):
    """MaxSimCuda._call_tn (kernels/MaxSimCuda.py:184-238): A [l, k, m], B [l, k, n] -> (vals, inds)."""
    lib = C.CDLL(_so("ref_max_sim_" + ("euclidean" if distance == "euclidean" else "inner")))
    lib.ref_max_sim_tn_launch.argtypes = [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_void_p]
    l, k, m = A.shape
    n = B.shape[2]
    size = m if dim == 2 else n
    vals = torch.full([l, size], float("-inf"), device=A.device, dtype=torch.float32)
    inds = torch.empty([l, size], device=A.device, dtype=torch.long)
    A, B = A.contiguous(), B.contiguous()
    rc = lib.ref_max_sim_tn_launch(A.data_ptr(), B.data_ptr(), vals.data_ptr(), inds.data_ptr(), m, n, k, dim, l, _stream(A.device))
    assert rc == 0, rc
    return vals, inds


def compute_centroids(data, labels, k):
    """ComputeCentroidsCuda.__call__ (kernels/ComputeCentroidsCuda.py:43-81): data [m, d, n], labels [m, n]."""
    dk = 256 if k >= 256 else 16
    assert k in (16, 256)
    lib = C.CDLL(_so(f"ref_compute_centroids_dk{dk}"))
    lib.ref_compute_centroids_launch.argtypes = [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p]
    m, d, n = data.shape
    cent = torch.zeros(m, d, k, device=data.device, dtype=torch.float32)
    data, labels = data.contiguous(), labels.contiguous()
    rc = lib.ref_compute_centroids_launch(data.data_ptr(), labels.data_ptr(), cent.data_ptr(), m, n, d, k, _stream(data.device))
    assert rc == 0, rc
    return cent


def pq_decode(codebook, code):
    """PQDecodeCuda.__call__ (kernels/PQDecodeCuda.py:38-65)."""
    lib = C.CDLL(_so("ref_pq_decode"))
    lib.ref_pq_decode_launch.argtypes = [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_void_p]
    m, d, _ = codebook.shape
    n = code.shape[1]
    res = torch.ones(m, d, n, device=codebook.device, dtype=torch.float32)
    codebook, code = codebook.contiguous(), code.contiguous()
    rc = lib.ref_pq_decode_launch(codebook.data_ptr(), code.data_ptr(), res.data_ptr(), m, d, n, _stream(codebook.device))
    assert rc == 0, rc
    return res.reshape(m * d, n)


def ivfpq_topk_residual_precomputed(data, part1, part2, cells, base_sims, cell_start, cell_size, is_empty,
                                    n_probe_list, k, tpb=256):
    """Reference kernel `ivfpq_topk_residual_precomputed` (kernels/cuda/ivfpq_topk.cu:1039-1207) with the host logic
    of IVFPQTopkCuda.topk_residual_precomputed (IVFPQTopkCuda.py:212-283)."""
    m = data.shape[0] * 4
    lib = _load(m, tpb)
    fn = lib.ref_ivfpq_topk_residual_precomputed_launch
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 12 + [C.c_int] * 4 + [C.c_void_p]
    n_data = data.shape[1]
    n_query, n_probe = cell_start.shape
    n_pow2 = 2 * (2 ** math.ceil(math.log2(math.ceil(k / 2))))
    dev = data.device
    tot_size = cell_size.sum(dim=1)
    values = torch.empty(n_query, n_pow2, device=dev, dtype=torch.float32).fill_(float("-inf"))
    indices = torch.zeros(n_query, n_pow2, device=dev, dtype=torch.int64)
    args = [t.contiguous() for t in (data, part1, part2, cells, base_sims, is_empty, cell_start, cell_size, tot_size, n_probe_list)]
    rc = fn(*[C.c_void_p(t.data_ptr()) for t in args], C.c_void_p(values.data_ptr()), C.c_void_p(indices.data_ptr()),
            n_data, n_query, n_probe, n_pow2, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    if rc != 0:
        raise RuntimeError(f"reference kernel launch failed: cudaError {rc}")
    return values[:, :k], indices[:, :k]


def placement_available() -> bool:
    return os.path.exists(_so("ref_placement"))


def get_ioa(labels):
    """GetIOACuda.__call__ (kernels/GetIOACuda.py:34-60) on the reference kernel get_ioa.cu:8-47."""
    lib = C.CDLL(_so("ref_placement"))
    lib.ref_get_ioa_launch.argtypes = [C.c_void_p] * 3 + [C.c_int] * 2 + [C.c_void_p]
    labels = labels.contiguous()
    uniq = torch.unique(labels)
    ioa = torch.zeros_like(labels) - 2
    rc = lib.ref_get_ioa_launch(labels.data_ptr(), uniq.data_ptr(), ioa.data_ptr(), labels.shape[0], uniq.shape[0], _stream(labels.device))
    assert rc == 0, rc
    return ioa


def get_write_address(is_empty, div_start, div_size, labels, ioa):
    """GetWriteAddressV2Cuda.__call__ (kernels/GetWriteAddressV2Cuda.py:33-66) on get_write_address_v2.cu:9-41;
    CellContainer.get_write_address passes (_is_empty, _cell_start, _cell_capacity, cells, ioa) (CellContainer.py:165-170)."""
    lib = C.CDLL(_so("ref_placement"))
    lib.ref_get_write_address_launch.argtypes = [C.c_void_p] * 6 + [C.c_int] * 2 + [C.c_void_p]
    args = [t.contiguous() for t in (is_empty, div_start, div_size, labels, ioa)]
    out = torch.zeros_like(labels) - 1
    rc = lib.ref_get_write_address_launch(*[t.data_ptr() for t in args], out.data_ptr(), is_empty.shape[0], labels.shape[0],
                                          _stream(labels.device))
    assert rc == 0, rc
    return out
