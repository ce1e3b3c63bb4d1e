"""CPU oracle for the MagellanMapper blob-detection hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import it.  Nothing under ``magellanmapper_b200/``
imports it, and the product path raises if its CUDA library is missing.

What it is
----------
A float64 numpy + scipy restatement of the reference's algorithm for the path
named by ``BASELINE.json:north_star``:

* ``skimage_restated``  - the scikit-image 0.25.2 functions the reference calls
  (``feature.blob_log`` with ``peak_local_max`` and ``_prune_blobs``,
  ``filters.gaussian``, ``morphology.erosion(octahedron(1))``, ``img_as_float``).
  scikit-image is a third-party dependency pinned at
  ``/root/reference/envs/requirements.txt:46`` and is NOT installed in this
  image, so these are restated from the published algorithm.  All filtering in
  them is delegated to ``scipy.ndimage`` / ``scipy.spatial`` exactly as
  scikit-image itself does, so the arithmetic is the same library code.
* ``magmap_restated``   - the reference's own Python on the path
  (``magmap/cv/chunking.py``, ``magmap/cv/detector.py``,
  ``magmap/cv/stack_detect.py``, ``magmap/plot/plot_3d.py:24-172``),
  re-expressed with explicit parameters instead of global config.

Parity pinning status
---------------------
* The reference's OWN Python (chunk geometry, block setup, blob table layout,
  seam pruning, sub-ROI orchestration, saturate/denoise glue) IS pinned:
  ``oracle/make_golden.py`` imports the unmodified reference from
  ``/root/reference`` in this container (missing GUI/IO wheels mocked, see
  ``oracle/ref_shim.py``), runs it, and commits the outputs under
  ``tests/golden/``; ``tests/test_oracle_vs_reference.py`` checks this package
  against those vectors bit for bit.
* The scikit-image internals are "PARITY UNPINNED": the reference's tests hold
  no golden blob values (``magmap/tests/test_image_stack_integration.py:62-72``
  only asserts ``nblobs > 0``) and scikit-image cannot be run here.  The scipy
  calls they reduce to are real.
"""
