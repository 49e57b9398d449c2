"""Generate golden vectors for the ADC-scan + top-k from the REFERENCE'S OWN KERNEL.

Run on the GPU box (needs oracle/_ref/, built by oracle/build_ref.py from /root/reference):
    gpurun -- python tests/golden/make_golden.py gpurun_out/golden
then copy gpurun_out/golden/*.npz into tests/golden/.  Each file holds the exact inputs of
fn.IVFPQTopk.topk (reference layouts) and the (values, address) the reference kernel
`ivfpq_topk` (torchpq/kernels/cuda/ivfpq_topk.cu:822-971, launched as IVFPQTopkCuda.py:121-141)
returned for them on a B200, plus `tie_free`: rows whose top-(k+1) scores are pairwise distinct.
tests/test_golden.py checks the CPU oracle against these files (no GPU needed).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from oracle import ivfpq_oracle as O, build_state as B, ref_kernels as R

CASES = [  # name, M, d, n_cells, n, n_probe, k, nq, kind, smart
    ("int_m8", 8, 32, 16, 3000, 6, 10, 24, "integer", False),
    ("int_m64", 64, 128, 16, 3000, 5, 100, 12, "integer", True),
    ("randn_m16", 16, 64, 32, 4000, 8, 20, 24, "randn", True),
    ("randn_m64_holes", 64, 128, 16, 3000, 6, 50, 12, "randn", False),
]


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    for name, M, d, C, n, n_probe, k, nq, kind, smart in CASES:
        if kind == "integer":
            st, queries = B.integer_state(d, M, C, n, seed=len(name), lo=-12, hi=13)
            x = queries(nq)
        else:
            torch.manual_seed(len(name))
            st = B.build_state(torch.randn(d, n), M, C, vq_iters=2, pq_iters=1)
            x = torch.randn(d, nq)
        if name.endswith("holes"):
            live = np.flatnonzero(st.is_empty == 0)
            st.is_empty[np.random.default_rng(0).choice(live, 300, replace=False)] = 1
        st.n_probe, st.use_smart_probing = n_probe, smart
        xx, sims, cells, npl = O.coarse_probe(st, x)
        lut = O.precompute_adc(xx, torch.from_numpy(st.pq_codebook), st.distance)
        cn = cells.numpy()
        cs, cz = st.cell_start[cn], st.cell_size[cn]
        g = lambda a: torch.as_tensor(a).cuda()
        rv, ra = R.ivfpq_topk(g(st.storage), lut.cuda(), g(cs), g(cz), g(st.is_empty), npl.cuda(), k)
        rv1, _ = R.ivfpq_topk(g(st.storage), lut.cuda(), g(cs), g(cz), g(st.is_empty), npl.cuda(), k + 1)
        rv, ra, rv1 = rv.cpu().numpy(), ra.cpu().numpy(), rv1.cpu().numpy()
        tie_free = np.all(np.diff(rv1, axis=1) != 0, axis=1)
        np.savez_compressed(os.path.join(out_dir, f"ivfpq_topk_{name}.npz"),
                            storage=st.storage, lut=lut.numpy(), is_empty=st.is_empty, cell_start=cs, cell_size=cz,
                            n_probe_list=npl.numpy(), k=np.int64(k), ref_values=rv, ref_address=ra, tie_free=tie_free)
        print(name, "rows", nq, "tie-free", int(tie_free.sum()))


def residual_and_placement(out_dir):
    """Round 2: golden vectors from the reference's `ivfpq_topk_residual_precomputed` kernel (ivfpq_topk.cu:1039-1207)
    and from its placement kernels get_ioa.cu / get_write_address_v2.cu."""
    g = lambda a: torch.as_tensor(a).cuda()
    for name, M, d, C, n, n_probe, k, nq, smart in [("res_m16", 16, 64, 16, 4000, 6, 20, 24, True),
                                                     ("res_m64", 64, 128, 16, 3000, 5, 50, 12, False)]:
        torch.manual_seed(len(name) + M)
        st = B.build_state_residual(torch.randn(d, n), M, C, vq_iters=2, pq_iters=1)
        st.n_probe, st.use_smart_probing = n_probe, smart
        x = torch.randn(d, nq)
        xx, sims, cells, npl = O.coarse_probe(st, x)
        p1, p2 = O.residual_parts(xx, torch.from_numpy(st.vq_codebook), torch.from_numpy(st.pq_codebook))
        cn = cells.numpy()
        cs, cz = st.cell_start[cn], st.cell_size[cn]
        rv, ra = R.ivfpq_topk_residual_precomputed(g(st.storage), p1.cuda(), p2.cuda(), cells.cuda(), sims.cuda(), g(cs), g(cz),
                                                   g(st.is_empty), npl.cuda(), k)
        rv1, _ = R.ivfpq_topk_residual_precomputed(g(st.storage), p1.cuda(), p2.cuda(), cells.cuda(), sims.cuda(), g(cs), g(cz),
                                                   g(st.is_empty), npl.cuda(), k + 1)
        rv, ra, rv1 = rv.cpu().numpy(), ra.cpu().numpy(), rv1.cpu().numpy()
        tie_free = np.all(np.diff(rv1, axis=1) != 0, axis=1)
        np.savez_compressed(os.path.join(out_dir, f"residual_{name}.npz"), storage=st.storage, part1=p1.numpy(), part2=p2.numpy(),
                            cells=cn, base_sims=sims.numpy(), is_empty=st.is_empty, cell_start=cs, cell_size=cz,
                            n_probe_list=npl.numpy(), k=np.int64(k), ref_values=rv, ref_address=ra, tie_free=tie_free)
        print("residual", name, "tie-free", int(tie_free.sum()))
    rng = np.random.default_rng(12)
    C, cap_cell, n = 23, 97, 900
    is_empty = (rng.random(C * cap_cell) < 0.55).astype(np.uint8)
    start = np.arange(C, dtype=np.int64) * cap_cell
    free = np.array([int(is_empty[s:s + cap_cell].sum()) for s in start], dtype=np.int64)
    cells = np.repeat(np.arange(C), np.minimum(free, rng.integers(5, 60, C)))
    rng.shuffle(cells)
    ioa = R.get_ioa(g(cells))
    w = R.get_write_address(g(is_empty), g(start), g(np.zeros(C, np.int64) + cap_cell), g(cells), ioa)
    np.savez_compressed(os.path.join(out_dir, "placement.npz"), cells=cells, is_empty=is_empty, cell_start=start,
                        cell_capacity=np.zeros(C, np.int64) + cap_cell, ref_ioa=ioa.cpu().numpy(), ref_write_address=w.cpu().numpy())
    print("placement", cells.shape[0], "items")


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
    if "--round2-only" not in sys.argv:
        main(out)
    os.makedirs(out, exist_ok=True)
    residual_and_placement(out)
