"""CPU oracle for the WMF raster-drainage hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may
import this package; the product (``wmf_heavy_b200``) never does.  See the header of
``wmf_oracle.c`` for what is restated, from which reference lines, and which parts are
pinned by the reference's own golden vectors (family A, ``wmf_py``) and which are
"parity unpinned" (family B, the Fortran modules that cannot be compiled here).
"""
from .oracle import *  # noqa: F401,F403
