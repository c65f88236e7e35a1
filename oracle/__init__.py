"""CPU oracle for the climatrix IDW / ordinary-kriging reconstruction hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``climatrix_b200/`` may import this
package: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline / ``--impl reference`` legs use it, and only as the checker or the
timed CPU arm -- never as a fallback for the CUDA path.

Parity status
-------------
* IDW (``idw_oracle``): restates ``src/climatrix/reconstruct/idw.py:135-180``
  with the *real* ``scipy.spatial.cKDTree``.  PINNED: ``oracle/make_golden.py``
  imports the unmodified reference ``IDWReconstructor`` from ``/root/reference``
  (third-party imports that are absent in this image are stubbed) and the
  outputs are committed under ``tests/golden/``; ``tests/test_oracle.py``
  checks this restatement against them bit for bit.
* Ordinary kriging (``ok_oracle``): the climatrix wrapper
  (``src/climatrix/reconstruct/kriging.py:147-250``) is pinned the same way
  (the real reference class drives it), but the arithmetic lives in the
  third-party ``pykrige`` package (un-vendored, un-pinned in
  ``pyproject.toml:65-67``; upstream GeoStat-Framework/PyKrige 1.7.x), which is
  not installable here.  ``pykrige_restated`` restates its published
  algorithm.  For that part: **parity unpinned**.
"""

from .idw_oracle import idw_oracle, idw_oracle_batched, knn_bruteforce, knn_ckdtree  # noqa: F401
from .ok_oracle import ok_oracle  # noqa: F401
from .pykrige_restated import OrdinaryKriging  # noqa: F401
