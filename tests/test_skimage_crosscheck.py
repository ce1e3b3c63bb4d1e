"""Cross-check of the restated scikit-image layer against a REAL scikit-image, whenever one
is importable (it is in neither this image nor the offline wheelhouse, so these tests skip
here; they are the pin the oracle's skimage layer otherwise lacks)."""
import numpy as np
import pytest

skimage = pytest.importorskip("skimage")

from oracle import skimage_restated as ski            # noqa: E402
from magellanmapper_b200 import synth                 # noqa: E402


def test_blob_log_equals_skimage():
    from skimage.feature import blob_log
    vol, _ = synth.make_volume((24, 60, 56), seed=3, density=1 / 1500.0)
    for lo, hi in ((3, 5), (4, 10)):
        want = blob_log(vol, min_sigma=lo, max_sigma=hi, num_sigma=10, threshold=0.1, overlap=0.5)
        got = ski.blob_log(vol, lo, hi, 10, 0.1, 0.5)
        assert sorted(map(tuple, got)) == sorted(map(tuple, want))


def test_filters_morphology_transform_equal_skimage():
    from skimage import filters, morphology, transform
    rng = np.random.default_rng(5)
    img = rng.random((20, 31, 27))
    np.testing.assert_array_equal(ski.filters_gaussian(img, 8), filters.gaussian(img, 8))
    np.testing.assert_array_equal(ski.erosion_octahedron1(img),
                                  morphology.erosion(img, morphology.octahedron(1)))
    np.testing.assert_array_equal(ski.ball(2), morphology.ball(2))
    lab = rng.integers(-1, 40, size=(12, 14, 13))
    np.testing.assert_array_equal(ski.dilation(lab, ski.ball(2)),
                                  morphology.dilation(lab, morphology.ball(2)))
    u16 = rng.integers(0, 65535, size=(9, 20, 22), dtype=np.uint16)
    for out_shape in ((27, 20, 22), (13, 12, 16)):
        np.testing.assert_array_equal(
            ski.transform_resize(u16, out_shape, mode="reflect", preserve_range=True),
            transform.resize(u16, out_shape, mode="reflect", preserve_range=True))
