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


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
