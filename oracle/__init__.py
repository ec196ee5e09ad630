"""CPU oracle for the ConsistentNeRF per-ray hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``consistentnerf_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or as the
timed CPU baseline -- never as the shipped path.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the unmodified reference
(``/root/reference/nerf-pytorch-master``) in the build container, runs it on seeded
inputs and commits its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks every function here against those vectors.
"""
from .nerf_oracle import *  # noqa: F401,F403
