"""CPU oracle for the set-abstraction / size-adaptive clustering hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``3dsad-main_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
or the timed CPU baseline -- never as the product path.

PARITY UNPINNED with respect to the reference: ``/root/reference`` holds a
two-line README.md and nothing else (README.md:1 ``# 3DSAD-main``, README.md:2
``Size Adaptive Clustering for 3D object detection in Point Clouds``), so there
are no golden vectors, tests or fixtures to pin against.  As BASELINE.json's
north_star instructs for this case, the oracle is a straightforward NumPy /
torch-CPU implementation of the ops from their definitions (SURVEY.md section
8(a), arithmetic order fixed by section 7 H1/H2).  Its own correctness is pinned
by hand-derived known-answer vectors in ``tests/test_oracle.py`` and the
committed fixtures under ``tests/golden/``.
"""
from .sad_oracle import *  # noqa: F401,F403
