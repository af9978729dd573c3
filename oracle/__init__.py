"""CPU oracle for the OCT metric suite -- TEST INFRASTRUCTURE ONLY.

This package restates, in numpy, the arithmetic of the reference's ``Metrics/``
modules (ZhangHH233/Retinal_OCT_Image_Segmentation_via_Deep_Learning).  It is
the checker the CUDA path is compared against.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import
it; the product package never does (``tests/test_no_oracle_in_product.py``
enforces that).

Pinning status
--------------
* ``metrics_oracle`` (confusion / region / pixel-error / biomarker formulas):
  PINNED -- checked bit-for-bit against the unmodified reference modules
  executed in the build container (``oracle/make_golden.py`` imports them
  from ``/root/reference/Metrics`` and writes ``tests/golden/*.npz``; the
  reference ships no tests or golden vectors of its own).
* ``contours_oracle.find_contours`` (scikit-image marching squares, version
  unpinned by the reference, scikit-image absent from this image, so it
  cannot be executed): PINNED TO PUBLISHED VECTORS -- the known-answer
  vectors scikit-image itself publishes (docstring doctest, ``test_binary``,
  ``test_float``; ``tests/golden/skimage_published_vectors.json``) are
  reproduced vertex for vertex, in order; squared distances are cross-checked
  against ``scipy.ndimage.distance_transform_edt`` on the doubled lattice.
* boundary extraction, 3-D surface distances: build-defined extensions with no
  reference counterpart.
"""
