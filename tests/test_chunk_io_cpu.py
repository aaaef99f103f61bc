"""spsg_b200.chunk_io.load_sdf (numpy.frombuffer reader) against what the reference's own data_util.load_sdf returned
for tests/golden/tiny_chunk.sdf (tests/golden/make_golden_chunk_io.py), plus a write -> read round trip at chunk size
and the error path.  CPU only."""
import os
import time

import numpy as np

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_reader_matches_reference_golden():
    from spsg_b200 import chunk_io
    g = np.load(os.path.join(GOLD, "chunk_io_ref.npz"))
    path = os.path.join(GOLD, "tiny_chunk.sdf")
    (locs, sdf), dims, w2g, known, color, sem = chunk_io.load_sdf(path, load_sparse=True, load_known=False, load_color=True)
    assert locs.dtype == np.int32 and sdf.dtype == np.float32 and known is None and sem is None
    np.testing.assert_array_equal(locs, g["sp_locs"])
    np.testing.assert_array_equal(sdf, g["sp_sdf"])
    assert list(dims) == list(g["sp_dims"])
    np.testing.assert_array_equal(w2g, g["sp_w2g"])
    np.testing.assert_array_equal(color, g["sp_color"])
    dense, w2g, known, color, sem = chunk_io.load_sdf(path, load_sparse=False, load_known=True, load_color=True, load_semantic=True)
    np.testing.assert_array_equal(dense, g["de_sdf"])          # -inf where absent
    np.testing.assert_array_equal(known, g["de_known"])
    np.testing.assert_array_equal(color, g["de_color"])
    np.testing.assert_array_equal(sem, g["de_sem"])


def test_round_trip_chunk_size_and_speed(tmp_path):
    from spsg_b200 import chunk_io, synthetic as S
    sdf_vox, _ = S.sdf_volume(2)
    zz, yy, xx = np.nonzero(np.abs(sdf_vox) < S.TRUNCATION)
    locs_xyz = np.stack([xx, yy, zz], 1)
    path = str(tmp_path / "chunk.sdf")
    dz, dy, dx = S.DIMS_ZYX
    rng = np.random.default_rng(0)
    color = rng.integers(0, 256, (dz, dy, dx, 3), dtype=np.uint8)
    known = np.ones((dz, dy, dx), np.uint8)
    chunk_io.write_chunk_file(path, (dx, dy, dz), S.VOXELSIZE, np.eye(4), locs_xyz, sdf_vox[zz, yy, xx] * S.VOXELSIZE,
                              known=known, color=color)
    t0 = time.perf_counter()
    (locs, sdf), dims, w2g, _, col, _ = chunk_io.load_sdf(path, load_sparse=True, load_known=False, load_color=True)
    dt = time.perf_counter() - t0
    assert list(dims) == [dz, dy, dx] and locs.shape == (len(zz), 3)
    np.testing.assert_array_equal(locs, np.stack([zz, yy, xx], 1))
    np.testing.assert_allclose(sdf, sdf_vox[zz, yy, xx], rtol=1e-6)
    np.testing.assert_array_equal(col, color)
    assert dt < 0.5, "reading one chunk took %.3f s" % dt   # the struct.unpack reader needs seconds for this


def test_unreadable_file_returns_nones(tmp_path, capsys):
    from spsg_b200 import chunk_io
    out = chunk_io.load_sdf(str(tmp_path / "missing.sdf"), load_sparse=True, load_known=False, load_color=False)
    assert out == (None, None, None, None, None)
    assert "failed to read file" in capsys.readouterr().out
    short = tmp_path / "short.sdf"
    short.write_bytes(b"\x01\x02\x03")
    assert chunk_io.load_sdf(str(short), True, False, False) == (None, None, None, None, None)
