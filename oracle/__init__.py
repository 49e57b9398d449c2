"""CPU oracle for the IVFPQ search hot path (TEST INFRASTRUCTURE ONLY).

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs (and its ``reference_gpu_kernel`` column: the reference's own kernel from
``oracle/_ref`` timed beside ours) may import it, and only as the checker or as a
timed baseline.

Parity status: **unpinned by reference fixtures** -- the reference ships no golden
vectors, known-answer tests or fixtures for ``IVFPQIndex.search`` (its ``tests/``
cover the three container classes only), and ``import torchpq``
needs CuPy + a CUDA device, so it cannot run in the build container.  The oracle
is pinned instead against the reference's *own CUDA kernel sources* compiled
unmodified into ``oracle/_ref/`` (see ``oracle/build_ref.py``) and executed on the
GPU box (``tests/test_ref_kernels_gpu.py``, ``tests/test_build_gpu.py``), against golden
vectors those kernels produced on a B200 (``tests/golden/``: plain and residual scans,
index of appearance, write addresses; checked on CPU by ``tests/test_golden.py``), and
against the container test cases the reference does ship (``tests/test_container_oracle.py``).
"""
