"""CPU oracle for the symmetrized-contraction hot path of Eike-Flath/symtensor.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import it, and only as the checker / the CPU arm -- never as a fallback for the CUDA path.

Parity status: PINNED.  Every function is checked (``tests/test_oracle_*.py``) against

* the reference's own golden vectors (``symtensor/tests/test_permcls_numpy.py:159-176``,
  ``symtensor/testing/api.py:76-82, 186-193, 235-240, 308-328``, ``symtensor/tests/test_utils.py:79-88``), and
* outputs of the unmodified reference itself, generated in the build container by
  ``oracle/gen_golden.py`` (which imports ``/root/reference`` through ``oracle/ref_shim``) and
  committed under ``tests/golden/``.

Modules
-------
index_oracle   integer bookkeeping: class order, sizes, multiplicities, storage order, rank/unrank
dense_oracle   the reference's *actual* algorithm (densify -> NumPy op -> r! symmetrize -> repack)
packed_oracle  the same four ops computed in packed space, NumPy
c_oracle       ctypes loader for ``symoracle.c`` (OpenMP, long-double accumulation) for full-size checks
"""
