"""GPU neighbourhood search vs the oracle, through the C ABI.  Bit-exact: sorted positions, permutation, every list."""
import numpy as np
import pytest

import yasph2d_b200 as y
from oracle import pyoracle as po
from util import assert_lists_equal, uniform_points

pytestmark = pytest.mark.gpu
capi = y.capi


def run_pair(pos, radius, boundary=None, **knobs):
    ns = y.NeighborhoodSearch(radius, max_particles=max(len(pos), 1), max_boundary=max(1, 0 if boundary is None else len(boundary)), **knobs)
    w = po.World(h=radius)
    w.set_particles(pos)
    if boundary is not None:
        sb = ns.update_static(boundary)
        w.set_boundary(boundary)
    w.update_neighborhood()
    if boundary is not None:
        assert np.array_equal(sb, w.boundary()), "sorted boundary differs"
    spos, _ = ns.update_dynamic(pos)
    assert np.array_equal(spos, w.positions()), "sorted positions differ"
    assert np.array_equal(ns.ctx.field(capi.FIELD_SORT_PERMUTATION), w.last_sorting())
    nl = ns.neighbor_lists()
    assert_lists_equal((nl.count_dynamic, nl.count_total, nl.lists), w.neighbors())
    st = w.neighbor_stats()
    rep = ns.last_report
    assert rep.total_neighbors == st["total"] and rep.neighbors_capped == st["capped"] and rep.neighbors_dropped == st["static_drops"]
    assert rep.num_cells == len(w.cells()[0]) - 1
    return ns, w


def test_reference_neighborhood_test():
    """neighborhood_search.rs:529-556: 1000 points, density 10, radius 1, seed 123456789 -- equals brute force in order."""
    pos = uniform_points(1000, 10.0, 123456789)
    ns, w = run_pair(pos, 1.0)
    sp = w.positions()
    nl = ns.neighbor_lists()
    for i in range(len(sp)):
        d = sp - sp[i]
        d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]
        bf = np.nonzero((d2 <= np.float32(1.0)) & (np.arange(len(sp)) != i))[0]
        assert np.array_equal(nl.neighbors_dynamic(i), bf)
        assert nl.num_neighbors(i) == len(bf) and len(nl.neighbors_static(i)) == 0


@pytest.mark.parametrize("n,radius,seed", [(20000, 1.0, 123456789), (20000, 0.5, 1), (50000, 0.75, 2), (30000, 1.25, 3), (257, 1.0, 4), (1, 1.0, 5)])
def test_uniform_points(n, radius, seed):
    """benches/benchmarks/neighborhood_search.rs:9-17 inputs (20000 points, density 10) and the config-5 radii."""
    run_pair(uniform_points(n, 10.0, seed), radius)


def test_static_neighbors_and_cap():
    """dynamic first then static, cap 64 total, static overflow dropped (ns.rs:353-381, deviation D4)."""
    rng = np.random.default_rng(5)
    pos = (rng.random((3000, 2)) * 9.0).astype(np.float32)
    bnd = (rng.random((4000, 2)) * 9.0).astype(np.float32)
    ns, w = run_pair(pos, 1.0, bnd)
    assert w.neighbor_stats()["capped"] > 0


def test_coincident_points_and_far_outlier():
    pos = np.array([[1.5, 1.5], [1.5, 1.5], [1.6, 1.5], [50.0, 50.0], [-99.99, -99.99], [1.5, 1.5]], np.float32)
    run_pair(pos, 1.0)


def test_far_apart_clusters_need_the_top_radix_digit():
    """Cell coordinates beyond 2^13.5: the sort's fourth pass (key bits 27..31), normally skipped, has to run.  One point sits in the
    cell whose key has all 27 low bits set (cell (16383, 8191) from grid_min = -100)."""
    rng = np.random.default_rng(17)
    a = (rng.random((6000, 2)) * 12.0).astype(np.float32)
    b = (rng.random((6000, 2)) * 12.0 + np.array([30000.0, 21000.0])).astype(np.float32)
    c = (rng.random((3000, 2)) * 6.0 + np.array([16280.0, 8088.0])).astype(np.float32)
    tie = np.array([[16283.5, 8091.5]], np.float32)
    pos = np.concatenate([a, b, c, tie])[rng.permutation(15001)]
    ns, w = run_pair(pos, 1.0)
    keys = ns.ctx.field(capi.FIELD_CELL_KEY)
    assert keys.max() >= (1 << 27) and (keys & 0x07FFFFFF == 0x07FFFFFF).any() and np.all(np.diff(keys.astype(np.int64)) >= 0)


def test_dam_break_scene_lists():
    """The application's scene (main.rs:177-196): 4050 fluid + 6840 boundary particles, h = 0.02."""
    ow = po.dam_break_scene(po.World())
    pos, bnd = ow.positions(), ow.boundary()
    run_pair(pos, float(ow.props()["h"]), bnd)


def test_clustered_points_many_per_cell():
    rng = np.random.default_rng(11)
    centers = rng.random((40, 2)) * 30.0
    pos = (centers[rng.integers(0, 40, 20000)] + rng.normal(0, 0.8, (20000, 2))).astype(np.float32)
    run_pair(pos, 1.0)


def test_oversized_tile_capacity_is_rejected():
    with pytest.raises(capi.YasphError) as e:
        y.NeighborhoodSearch(1.0, max_particles=16, tile_dynamic_capacity=16384, tile_static_capacity=16384)
    assert e.value.status == 3


def test_tiles_too_large_to_stage_fall_back_to_global_memory():
    """1250 particles per cell: far more candidates per tile than the configured staging capacity (and, without that knob, than
    the shared memory of an SM holds).  The reference accepts any density up to its 64-neighbour cap
    (neighborhood_search.rs:353-381); such tiles are processed unstaged, from global memory -- the lists equal the oracle's."""
    rng = np.random.default_rng(3)
    pos = (rng.random((5000, 2)) * 2.0).astype(np.float32)
    ns = y.NeighborhoodSearch(1.0, max_particles=5000, tile_dynamic_capacity=256)
    spos, _ = ns.update_dynamic(pos)
    w = po.World(h=1.0)
    w.set_particles(pos)
    w.update_neighborhood()
    assert np.array_equal(spos, w.positions())
    assert_lists_equal(ns.ctx.neighbors(True), w.neighbors())
    assert ns.last_report.neighbors_capped == w.neighbor_stats()["capped"] > 0  # every list is cut at 64 here


def test_capacity_error_is_loud():
    """What stays an error: more than 65535 staged candidates in one tile (neighbour slots are 16 bit)."""
    rng = np.random.default_rng(3)
    pos = (rng.random((70000, 2)) * 2.0).astype(np.float32)
    ns = y.NeighborhoodSearch(1.0, max_particles=70000)
    with pytest.raises(capi.YasphError) as e:
        ns.update_dynamic(pos)
    assert e.value.status == 3  # YASPH_ERR_CAPACITY
