"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

``oracle/`` holds the CPU checker for the per-timestep particle loop of Nano-kappa
(``/root/reference/classes/Population.py:1724`` ``run_timestep`` and the seams it calls in
``Mesh``, ``Geometry.SubvolClassifier`` and ``Phonon``):

* ``ref_harness.py``  loads the UNMODIFIED reference sources where they lie under
  ``/root/reference`` (only possible in the build container) with the import shims the survey
  lists (stubs for matplotlib/trimesh/shapely/h5py/phonopy, ``ndarray.ptp`` rewrite, NumPy-1
  ``linalg.solve`` broadcasting).  Used to pin the restatement and to generate ``tests/golden``.
* ``nk_oracle.py``    NumPy restatement of the hot path, each function citing the reference
  file:line it follows.  This one travels to the GPU box.  Pinned bit-for-bit against the
  reference run here (``tests/test_oracle_pin.py`` + fixtures made by ``gen_golden.py``).
* ``gen_golden.py``   regenerates ``tests/golden/*.npz`` by executing the reference.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this package.  ``nanokappa_b200`` never does: the product path fails loudly when
the CUDA library is missing instead of falling back to anything in here.
"""
