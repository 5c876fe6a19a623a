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
* ``maps_np`` (MapGen, FourierCalc, tapers, masks, beam, filter_map, split_calc,
  noise_from_splits, silc/cilc) and ``lensing_np`` (flat_taylens, kappa_to_phi,
  SplitLensing) -- the logic orphics itself owns is PINNED: the reference's own
  class / function bodies are cut out of /root/reference with ``ast`` and executed
  unmodified over the pixell stand-in below; their outputs are the golden vectors
  ``tests/golden/{maps_refbody,lensing_refbody,split_callers,ilc}.npz``
  (``tests/golden/make_golden_{maps,lensing,callers,ilc}.py``).
* ``enmap_np`` (the pixell layer: geometry, lmap, fft normalisations,
  rand_gauss_harm, map_mul, spec2flat, queb_rotmat, harm2map) -- PARITY UNPINNED:
  third-party package ``pixell`` (unpinned, not in the reference's
  requirements.txt, not installed, no network).  Restated from pixell's
  published behaviour; anchored on the reference's call sites
  (maps.py:1553-1677) and on its notebooks' known answers (geometry printout
  tutorials/demo-grf.ipynb:97; binned/theory -> 1).
* ``qe_np`` (lensing.qest) -- PARITY UNPINNED: the class is absent from the
  reference snapshot (only call sites tutorials/tt_verification.ipynb:81,608,610
  and lensing.py:973-976).  Restated from the historical estimator and checked
  physically (brute-force normalisation, unit response).
"""
