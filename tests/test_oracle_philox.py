"""Philox oracle (oracle/philox.py) against the Random123 known-answer vectors and basic statistics."""
import numpy as np

from oracle.philox import KAT, normals, philox4x32_10, uniforms


def test_philox4x32_10_known_answers():
    for ctr, key, want in KAT:
        out = philox4x32_10(np.array([ctr], dtype=np.uint32), np.array(key, dtype=np.uint32))[0]
        assert tuple(int(v) for v in out) == want


def test_uniforms_stay_in_unit_interval():
    u = uniforms(np.array([0, 1, 2 ** 31, 2 ** 32 - 1], dtype=np.uint32))
    assert u.dtype == np.float32 and (u > 0).all() and (u <= 1).all()


def test_normal_field_statistics_and_determinism():
    z = normals(1 << 18, seed=1234, step=7)
    assert z.dtype == np.float32 and abs(float(z.mean())) < 0.01 and abs(float(z.std()) - 1.0) < 0.01
    assert abs(float((z ** 3).mean())) < 0.03 and abs(float((z ** 4).mean()) - 3.0) < 0.1
    assert np.array_equal(z, normals(1 << 18, seed=1234, step=7))
    assert not np.array_equal(z, normals(1 << 18, seed=1234, step=8))
    assert not np.array_equal(z, normals(1 << 18, seed=1235, step=7))
    assert np.array_equal(normals(10, 5, 0), normals(12, 5, 0)[:10])   # ragged tail = prefix of the next group
