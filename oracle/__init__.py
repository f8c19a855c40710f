"""CPU oracle for the MOBODY rollout / Q-weighted-BC hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline.  The product path (``mobody_b200``) never imports it
and fails loudly when the CUDA extension is missing.

Parity status: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the reference itself, generated in
the build container by ``oracle/make_golden.py`` (imports ``/root/reference``)
and committed under ``tests/golden/``.  ``tests/test_oracle_golden.py`` checks
the oracle against those vectors on every run (rollout, train step, classifier,
refresh block, dynamics fitting step).  ``oracle/ingest_oracle.py`` (dataset walk,
batched evaluator) restates reference modules that cannot be imported here (gym /
d4rl / h5py at module level): parity unpinned for those two host-side helpers.
"""
from .mobody_oracle import *  # noqa: F401,F403
from .philox import philox4x32_10, philox_uniform, philox_normal_pairs, recipe_fill  # noqa: F401
