"""Shared helpers for the parity tests."""
import numpy as np

from oracle import pyoracle as po

f32 = np.float32
REL_TOL = 1e-4  # BASELINE.json north_star: per-particle density / pressure / acceleration within 1e-4 relative (f32)


def uniform_points(n, density, seed):
    """neighborhood_search.rs:535-538 / benches/benchmarks/neighborhood_search.rs:14-17 (SmallRng restated)."""
    r = np.zeros(2 * n, np.float32)
    po.lib().yo_rng_fill(seed, r.ctypes.data_as(po.C.POINTER(po.C.c_float)), 2 * n)
    return (r.reshape(n, 2) * f32(np.sqrt(f32(n) / f32(density)))).astype(np.float32)


def assert_close(got, ref, what, rel=REL_TOL, floor_frac=1e-3):
    """|got - ref| <= rel * max(|ref|, floor), floor = floor_frac * typical magnitude."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    if got.size == 0:
        return
    floor = floor_frac * max(np.abs(ref).mean(), 1e-30)
    err = np.abs(got - ref) / np.maximum(np.abs(ref), floor)
    worst = int(np.argmax(err))
    assert err.flat[worst] <= rel, "%s: max rel err %.3e at %d (got %r, ref %r)" % (what, err.flat[worst], worst, got.flat[worst], ref.flat[worst])


def assert_lists_equal(gpu_lists, ora_lists):
    gcd, gct, gl = gpu_lists
    ocd, oct_, ol = ora_lists
    assert np.array_equal(gcd, ocd), "count_dynamic differs at %s" % np.nonzero(gcd != ocd)[0][:10]
    assert np.array_equal(gct, oct_), "count_total differs at %s" % np.nonzero(gct != oct_)[0][:10]
    k = np.arange(gl.shape[1])[None, :]
    mask = k < gct[:, None]
    bad = np.nonzero(((gl != ol) & mask).any(axis=1))[0]
    assert len(bad) == 0, "neighbour lists differ for particles %s: gpu %s oracle %s" % (bad[:5], gl[bad[0], : gct[bad[0]]], ol[bad[0], : oct_[bad[0]]])
