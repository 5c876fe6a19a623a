"""CPU oracle for the orphics flat-sky Fourier hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It restates, in plain numpy/scipy, the reference algorithm for the path named
by BASELINE.json's north_star (MapGen.get_map -> FourierCalc.power2d ->
stats.bin2D.bin -> lensing.qest), every function citing the reference
file:line it follows.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
Nothing under ``orphics_b200/`` imports it; the product path fails loudly when
the CUDA library is missing.

Parity pinning status (see DESIGN.md "Oracle"):

* ``stats_np.bin2D`` / ``Statistics`` / ``mpi_distribute`` -- PINNED against the
  reference's own importable ``orphics.stats`` / ``orphics.mpi`` through the
  golden vectors in ``tests/golden/`` (made by ``tests/golden/make_golden.py``
  with ``PYTHONPATH=/root/reference``).
* ``enmap_np`` / ``maps_np`` (MapGen, FourierCalc, helpers) -- PARITY UNPINNED:
  the arithmetic lives in the third-party package ``pixell`` (unpinned, not in
  the reference's requirements.txt, not installed, no network).  Restated from
  pixell's published behaviour; anchored on the reference's call sites
  (maps.py:1553-1677) and on its notebooks' known answers (geometry printout
  tutorials/demo-grf.ipynb:97; binned/theory -> 1).
* ``qe_np`` (lensing.qest) -- PARITY UNPINNED: the class is absent from the
  reference snapshot (only call sites tutorials/tt_verification.ipynb:81,608,610
  and lensing.py:973-976).  Restated from the historical estimator and checked
  physically (brute-force normalisation, unit response).
"""
