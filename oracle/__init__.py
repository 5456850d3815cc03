"""CPU oracle for the GCC pretraining hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import this package.  ``gcc_b200`` (the product)
never imports it and has no CPU fallback.

Parity status: the reference's arithmetic for sampling,
induction, aggregation and pooling lives in DGL 0.4.3, which is absent here, so
those parts are **parity unpinned** against DGL; the reference-owned code
(budget formula, node ordering, eigen-decomposition call, GIN/MLP/BN structure,
MemoryMoCo, NCE criteria, LR schedule) is pinned by golden vectors generated
from the real reference modules (tests/golden/make_golden.py).
"""
