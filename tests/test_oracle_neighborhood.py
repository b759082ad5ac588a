"""The reference's neighbourhood test (src/sph/neighborhood_search.rs:529-556) restated against the oracle,
plus checks of the compact-cell traversal (ns.rs:169-259) against brute force."""
import numpy as np
import pytest

from oracle import pyoracle as po

f32 = np.float32


def uniform_points(n, density, seed):
    """ns.rs:535-538 / benches/benchmarks/neighborhood_search.rs:14-17 (SmallRng restated, deviation D3)."""
    r = np.zeros(2 * n, np.float32)
    po.lib().yo_rng_fill(seed, r.ctypes.data_as(po.C.POINTER(po.C.c_float)), 2 * n)
    return (r.reshape(n, 2) * f32(np.sqrt(f32(n) / f32(density)))).astype(np.float32)


def brute_force(pos, radius, others=None):
    """{j != i : d2 <= r2} in ascending j, d2 evaluated in f32 exactly as cgmath distance2 (no FMA)."""
    src = pos if others is None else others
    r2 = f32(radius) * f32(radius)
    out = []
    for i in range(len(pos)):
        dx = src[:, 0] - pos[i, 0]
        dy = src[:, 1] - pos[i, 1]
        d2 = dx * dx + dy * dy  # numpy f32 elementwise: mul, mul, add, each rounded
        m = (d2 <= r2) & (d2 > f32(1e-10))
        out.append(np.nonzero(m)[0].astype(np.uint32))
    return out


def test_neighbors_contains_neighbors():  # ns.rs:529-556: 1000 points, density 10, radius 1, seed 123456789
    pos = uniform_points(1000, 10.0, 123456789)
    w = po.World(h=1.0)
    w.set_particles(pos)
    w.update_neighborhood()
    sp = w.positions()
    cd, ct, lists = w.neighbors()
    bf = brute_force(sp, 1.0)
    for i in range(len(sp)):
        assert ct[i] == cd[i]
        assert np.array_equal(lists[i, : cd[i]], bf[i])


def test_sort_is_stable_morton_order():
    pos = uniform_points(3000, 10.0, 7)
    w = po.World(h=0.75)
    w.set_particles(pos)
    w.update_neighborhood()
    perm = w.last_sorting()
    keys = np.array([po.lib().yo_position_to_cidx(f32(0.75), p[0], p[1]) for p in pos], np.uint32)
    assert np.array_equal(perm, np.argsort(keys, kind="stable").astype(np.uint32))
    assert np.array_equal(w.positions(), pos[perm])
    fp, ci = w.cells()
    sk = keys[perm]
    heads = np.nonzero(np.concatenate([[True], sk[1:] != sk[:-1]]))[0]
    assert np.array_equal(fp[:-1], heads.astype(np.uint32)) and np.array_equal(ci[:-1], sk[heads])
    assert fp[-1] == len(pos) and ci[-1] == 0xFFFFFFFF  # sentinel, ns.rs:161-164


def test_runs_cover_exactly_the_3x3_box():
    """get_particle_runs_in_neighborbox (ns.rs:191-259): <=5 ascending runs == all particles of the 9 cells."""
    pos = uniform_points(4000, 10.0, 99)
    h = 1.0
    w = po.World(h=h)
    w.set_particles(pos)
    w.update_neighborhood()
    sp = w.positions()
    L = po.lib()
    keys = np.array([L.yo_position_to_cidx(f32(h), p[0], p[1]) for p in sp], np.uint32)
    cx = np.array([L.yo_morton_decode_x(int(k)) for k in keys])
    cy = np.array([L.yo_morton_decode_y(int(k)) for k in keys])
    fp, ci = w.cells()
    for c in range(0, len(ci) - 1, 7):
        x, y = L.yo_morton_decode_x(int(ci[c])), L.yo_morton_decode_y(int(ci[c]))
        expect = np.nonzero((abs(cx - x) <= 1) & (abs(cy - y) <= 1))[0]
        runs = w.runs(ci[c])
        got = np.concatenate([np.arange(a, b) for a, b in runs if b > a])
        assert np.array_equal(got, expect)


def test_static_neighbors_and_cap_64():
    """dynamic first then static, ascending; cap at 64 total (ns.rs:353-381); deviation D4 counted."""
    rng = np.random.default_rng(5)
    pos = (rng.random((600, 2)) * 4.0).astype(np.float32)
    bnd = (rng.random((900, 2)) * 4.0).astype(np.float32)
    w = po.World(h=1.0)
    w.set_particles(pos)
    w.set_boundary(bnd)
    w.update_neighborhood()
    sp, sb = w.positions(), w.boundary()
    cd, ct, lists = w.neighbors()
    bd, bs = brute_force(sp, 1.0), brute_force(sp, 1.0, sb)
    stats = w.neighbor_stats()
    assert stats["capped"] > 0  # density 37/unit^2 -> mean 118 dynamic candidates: the cap is exercised
    for i in range(len(sp)):
        d = bd[i][:64]
        s = bs[i][: 64 - len(d)]
        assert cd[i] == len(d) and ct[i] == len(d) + len(s)
        assert np.array_equal(lists[i, : cd[i]], d)
        assert np.array_equal(lists[i, cd[i] : ct[i]], s)


def test_coincident_and_empty():
    w = po.World(h=1.0)
    pos = np.array([[1.5, 1.5], [1.5, 1.5], [1.6, 1.5], [50.0, 50.0]], np.float32)
    w.set_particles(pos)
    w.update_neighborhood()
    cd, ct, lists = w.neighbors()
    sp = w.positions()
    # coincident particles (d2 <= 1e-10) are excluded like self (ns.rs:323,357)
    for i in range(4):
        expect = [j for j in range(4) if 1e-10 < ((sp[j] - sp[i]) ** 2).sum() <= 1.0]
        assert list(lists[i, : cd[i]]) == expect
    w2 = po.World(h=1.0)
    w2.set_particles(np.zeros((0, 2), np.float32))
    w2.update_neighborhood()
    assert w2.n == 0 and w2.cells()[0].tolist() == [0]


def test_dam_break_scene_counts():
    """main.rs:177-196 with world (2.0, 10000, 100) main.rs:85-89 -> 4050 fluid, ~6840 boundary particles."""
    w = po.dam_break_scene(po.World())
    p = w.props()
    assert w.n == 45 * 90
    assert abs(w.m - 6840) <= 16
    assert p["h"] == f32(0.02) and p["mass"] == f32(0.01) and p["radius"] == f32(0.005)
    pos = w.positions()
    step = f32(0.5) / f32(45)
    assert pos[:, 0].min() >= 0.1 + 0.5 * 0.05 * step - 1e-6 and pos[:, 0].max() <= 0.1 + 44 * step + 0.05 * step + 1e-6
