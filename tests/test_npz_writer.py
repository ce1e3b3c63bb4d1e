"""The archive writer behind ``Blobs.save_archive`` against ``numpy.savez``: same members,
dtypes and values through ``numpy.load``, valid checksums, and the ZIP64 directory forms."""
import os
import zipfile

import numpy as np
import pytest

from magellanmapper_b200.io import npz_writer


def _same(path_a, path_b):
    with np.load(path_a, allow_pickle=True) as a, np.load(path_b, allow_pickle=True) as b:
        assert a.files == b.files
        for k in a.files:
            assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape, k
            if a[k].dtype == object:
                assert a[k].tolist() == b[k].tolist(), k
            else:
                np.testing.assert_array_equal(a[k], b[k], err_msg=k)


def _archive(rows, rng):
    table = rng.random((rows, 8))
    return {
        "ver": 5, "segments": table, "resolutions": np.array([[1.0, 0.5, 0.5]]),
        "basename": "sample", "offset": None, "roi_size": (3, 4, 5), "colocs": None,
        "cols": ["z", "y", "x", "radius"], "empty": np.zeros((0, 8)),
        "strided": np.asfortranarray(rng.integers(0, 9, (700, 900)).astype(np.uint16)),
        "big_u16": rng.integers(0, 65535, (3, 1500, 1200)).astype(np.uint16),
        "big_strided": table[:, ::2], "bools": table[:, 0] > 0.5,
        "swapped": table[:70000].astype(">f8"),
        "record": np.zeros(3, dtype=[("a", "<i4"), ("b", "<f8")]),
    }


@pytest.mark.parametrize("rows", [0, 7, 600000])
def test_equals_numpy_savez(tmp_path, rows):
    arc = _archive(rows, np.random.default_rng(rows))
    want, got = str(tmp_path / "want.npz"), str(tmp_path / "got.npz")
    np.savez(want, **arc)
    npz_writer.savez(got, arc)
    assert os.path.getsize(want) == os.path.getsize(got)
    with zipfile.ZipFile(got) as z:
        assert z.testzip() is None
        assert [i.filename for i in z.infolist()] == [k + ".npy" for k in arc]
        assert all(i.compress_type == zipfile.ZIP_STORED for i in z.infolist())
    _same(want, got)
    # members that need no pickle load without it, as the reference's reader expects
    with np.load(got) as z:
        np.testing.assert_array_equal(z["segments"], arc["segments"])
        with pytest.raises(ValueError):
            z["offset"]


def test_suffix_rule_and_overwrite(tmp_path):
    arc = {"a": np.arange(5)}
    npz_writer.savez(str(tmp_path / "plain"), arc)
    assert os.path.exists(tmp_path / "plain.npz")
    npz_writer.savez(str(tmp_path / "exact.bin"), arc, add_suffix=False)
    assert os.path.exists(tmp_path / "exact.bin")
    # a shorter archive over a longer one leaves no tail behind
    big = {"a": np.arange(2 << 20)}
    npz_writer.savez(str(tmp_path / "plain"), big)
    npz_writer.savez(str(tmp_path / "plain"), arc)
    with np.load(tmp_path / "plain.npz") as z:
        np.testing.assert_array_equal(z["a"], arc["a"])
    assert os.path.getsize(tmp_path / "plain.npz") < 1000


def test_zip64_directory_records(tmp_path, monkeypatch):
    """Offsets and sizes past the 32-bit forms: the limits are lowered so that a 10 MB
    archive takes the ZIP64 records a 4 GB one would."""
    rng = np.random.default_rng(3)
    arc = {"first": rng.random((700000,)), "second": rng.random((600000,)), "n": 3}
    monkeypatch.setattr(npz_writer, "ZIP64_LIMIT", 1 << 20)
    monkeypatch.setattr(zipfile, "ZIP64_LIMIT", 1 << 20)
    got = str(tmp_path / "big.npz")
    npz_writer.savez(got, arc)
    with zipfile.ZipFile(got) as z:
        assert z.testzip() is None
        assert z.getinfo("second.npy").header_offset > (1 << 20)
    with np.load(got) as z:
        for k in arc:
            np.testing.assert_array_equal(z[k], arc[k])
    monkeypatch.setattr(npz_writer, "ZIP_FILECOUNT_LIMIT", 2)
    many = {f"k{i}": np.arange(i) for i in range(5)}
    npz_writer.savez(got, many)
    with np.load(got) as z:
        assert z.files == list(many)
        np.testing.assert_array_equal(z["k4"], np.arange(4))


def test_plain_write_when_the_file_cannot_be_mapped(tmp_path, monkeypatch):
    def refuse(*a, **k):
        raise OSError("no fallocate here")
    monkeypatch.setattr(os, "posix_fallocate", refuse)
    arc = _archive(600000, np.random.default_rng(5))
    want, got = str(tmp_path / "want.npz"), str(tmp_path / "got.npz")
    np.savez(want, **arc)
    npz_writer.savez(got, arc)
    with zipfile.ZipFile(got) as z:
        assert z.testzip() is None
    _same(want, got)
