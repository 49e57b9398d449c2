"""CPU oracle for the IVFPQ search hot path (TEST INFRASTRUCTURE ONLY).

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it, and only as the checker or as the timed CPU baseline.

Parity status: **unpinned by reference fixtures** -- the reference ships no golden
vectors, known-answer tests or fixtures for ``IVFPQIndex.search`` (its ``tests/``
cover the three container classes only), and ``import torchpq``
needs CuPy + a CUDA device, so it cannot run in the build container.  The oracle
is pinned instead against the reference's *own CUDA kernel sources* compiled
unmodified into ``oracle/_ref/`` (see ``oracle/build_ref.py``) and executed on the
GPU box (``tests/test_ref_kernels_gpu.py``), and against the container test cases
the reference does ship (``tests/test_container_oracle.py``).
"""
