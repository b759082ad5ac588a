"""Host side of the 1-D slab decomposition (multi-GPU): partitioning of a scene over cell columns and the per-rank driver.

The reference is single-process; this is the host logic that sits between its `Solver` surface and the
per-GPU contexts: every rank owns the particles of a contiguous range of cell columns (the x cell index of
src/sph/neighborhood_search.rs:45-64), uploads only those, and steps in lock step with the other ranks.  Everything in
here is array bookkeeping on the host -- particle physics happens in libyasph_gpu.so only.
"""
import ctypes as C

import numpy as np

from . import _capi as capi
from .host import GpuContext, _f32p

f32 = np.float32
GHOST_COLUMNS = 12  # default width of the ghost layer (cell columns): no per-pass halo exchange in a DFSPH step of up to ~4 Jacobi iterations + warm start


def cell_columns(x, smoothing_length, grid_min_x=-100.0):
    """Cell column of every x coordinate: ((x - grid_min) * (1 / cell_size)) as u16 in f32 arithmetic, saturating like
    Rust's `as u16` (neighborhood_search.rs:52-58, :475) -- the same arithmetic the device uses (yasph_cell_column)."""
    x = np.asarray(x, np.float32)
    inv = f32(1.0) / f32(smoothing_length)
    c = (x - f32(grid_min_x)) * inv
    c = np.where(np.isnan(c), f32(0.0), c)
    return np.clip(np.trunc(c), 0, 65535).astype(np.uint32)


def partition_columns(columns, world, min_width=2):
    """Splits the column axis into `world` adjacent ranges [lo, hi) holding about the same number of particles each.

    Returns a list of (lo, hi); the first range starts at 0 and the last ends at 65536 so that every position has an
    owner.  Cuts are only placed between columns (a cell column is never shared) and every range keeps at least
    `min_width` occupied-span columns.
    """
    columns = np.asarray(columns, np.uint32)
    if world < 1:
        raise ValueError("world must be >= 1")
    if world == 1:
        return [(0, 65536)]
    if columns.size == 0:
        edges = [int(round(65536 * r / world)) for r in range(world + 1)]
        return list(zip(edges[:-1], edges[1:]))
    cmin, cmax = int(columns.min()), int(columns.max())
    if cmax - cmin + 1 < min_width * world:
        raise ValueError("cannot split the %d occupied columns [%d, %d] into %d slabs of >= %d columns" % (cmax - cmin + 1, cmin, cmax, world, min_width))
    cum = np.cumsum(np.bincount(columns - cmin, minlength=cmax - cmin + 1).astype(np.int64))
    total = int(cum[-1])
    cuts, prev = [], cmin
    for r in range(1, world):
        k = int(np.searchsorted(cum, total * r / world, side="left"))  # first column whose cumulative count reaches the target
        cut = cmin + k + 1                                             # cut after that column
        cut = max(cut, prev + min_width)                               # this slab keeps min_width columns ...
        cut = min(cut, cmax + 1 - min_width * (world - r))             # ... and so do the ones to the right
        cuts.append(cut)
        prev = cut
    bounds = [0] + cuts + [65536]
    return list(zip(bounds[:-1], bounds[1:]))


def owned_mask(columns, lo, hi):
    columns = np.asarray(columns)
    return (columns >= lo) & (columns < hi)


def comm_unique_id():
    """ncclGetUniqueId through the C ABI (call on one rank, broadcast the bytes over any host channel)."""
    buf = (C.c_ubyte * capi.COMM_ID_BYTES)()
    capi.check(capi.lib().yasph_comm_unique_id(buf, capi.COMM_ID_BYTES))
    return bytes(buf)


def broadcast_unique_id(dist, src=0):
    """Creates the NCCL id on rank `src` and broadcasts it with torch.distributed (any backend)."""
    obj = [comm_unique_id() if dist.get_rank() == src else None]
    dist.broadcast_object_list(obj, src=src)
    return obj[0]


class LoopbackFabric:
    """All ranks in one process (one thread per rank): the transport used to test the slab logic on a single GPU."""

    def __init__(self, world):
        self.world = world
        h = C.c_void_p()
        capi.check(capi.lib().yasph_loopback_create(world, C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            capi.lib().yasph_loopback_destroy(self.h)
            self.h = None


class SlabContext(GpuContext):
    """A GpuContext that owns one slab of the domain.

    comm: bytes (NCCL unique id shared by all ranks) or a LoopbackFabric.
    """

    def __init__(self, cfg, rank, world, comm, col_range, n_global, id_base=0, track_ids=True):
        if track_ids:  # ids travel with the particles (tests, verification); a production run does not need them
            cfg.flags |= capi.FLAG_TRACK_IDS
        super().__init__(cfg)
        self.rank, self.world = rank, world
        if isinstance(comm, LoopbackFabric):
            self._ck(capi.lib().yasph_comm_init_loopback(self.h, comm.h, rank))
        elif world > 1:
            buf = (C.c_ubyte * capi.COMM_ID_BYTES).from_buffer_copy(comm)
            self._ck(capi.lib().yasph_comm_init(self.h, rank, world, buf, capi.COMM_ID_BYTES))
        self._ck(capi.lib().yasph_slab_set(self.h, int(col_range[0]), int(col_range[1]), int(n_global), int(id_base)))

    def info(self):
        out = capi.SlabInfo()
        self._ck(capi.lib().yasph_slab_get(self.h, C.byref(out)))
        return out

    def step_host_slab(self, pos, vel, dens, n_in, input_unchanged=False):
        """pos/vel/dens: host arrays with spare capacity; returns (report, n_out).  input_unchanged: the arrays still hold what the
        previous call handed back (the upload is skipped)."""
        rep = capi.StepReport()
        n_out = C.c_uint32(0)
        self._ck(capi.lib().yasph_step_host_slab_ex(self.h, _f32p(pos), _f32p(vel), _f32p(dens), int(n_in), len(pos),
                                                    capi.HOST_INPUT_UNCHANGED if input_unchanged else 0, C.byref(n_out), C.byref(rep)))
        return rep, n_out.value

    def local_field(self, field, dtype, width=1):
        n = self.info().n_local
        out = np.empty((n, width) if width > 1 else n, dtype)
        self._ck(capi.lib().yasph_download_field(self.h, field | capi.FIELD_LOCAL_BIT, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def ids(self):
        n, _ = self.counts()
        out = np.empty(n, np.uint32)
        self._ck(capi.lib().yasph_download_field(self.h, capi.FIELD_ID, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def local_neighbors(self):
        """Neighbour lists of the rank's local array (owned + ghosts), local indexing."""
        n = self.info().n_local
        cd = np.zeros(n, np.uint16)
        ct = np.zeros(n, np.uint16)
        lists = np.zeros((n, capi.MAX_NEIGHBORS), np.uint32)
        self._ck(capi.lib().yasph_neighbors_download(self.h, cd.ctypes.data_as(C.POINTER(C.c_uint16)), ct.ctypes.data_as(C.POINTER(C.c_uint16)),
                                                     lists.ctypes.data_as(C.POINTER(C.c_uint32))))
        return cd, ct, lists


def scatter_scene(positions, smoothing_length, grid_min_x, world, ranges=None):
    """Partition of a scene every rank holds in full: returns (ranges, own) with own[r] = indices of rank r's particles."""
    positions = np.asarray(positions, np.float32).reshape(-1, 2)
    cols = cell_columns(positions[:, 0], smoothing_length, grid_min_x)
    if ranges is None:
        ranges = partition_columns(cols, world)
    own = [np.nonzero(owned_mask(cols, lo, hi))[0] for lo, hi in ranges]
    assert sum(len(o) for o in own) == len(positions)
    return ranges, own


def make_slab_context(base_cfg, rank, world, comm, positions, velocities, boundary, ranges=None, track_ids=True):
    """Partitions a scene (identical arrays on every rank), creates this rank's SlabContext and uploads its particles.

    Particle ids (YASPH_FIELD_ID) are positions in the concatenation own[0] ++ own[1] ++ ...; the returned `id_to_global`
    maps them back to indices of `positions`.  Returns (ctx, ranges, id_to_global).
    """
    positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 2)
    ranges, own = scatter_scene(positions, base_cfg.smoothing_length, base_cfg.grid_min[0], world, ranges)
    cfg = capi.Config.from_buffer_copy(base_cfg)
    if cfg.ghost_columns == 0:
        # as wide a ghost layer as the narrowest slab between two others allows (W <= width - 1: a migrant that arrives from one side must not
        # belong to the ghost layer of the OTHER side), at most GHOST_COLUMNS (see yasph_config.ghost_columns)
        inner = [hi - lo for lo, hi in ranges[1:-1]]
        cfg.ghost_columns = max(1, min([GHOST_COLUMNS] + [w - 1 for w in inner]))
    id_base = int(sum(len(o) for o in own[:rank]))
    ctx = SlabContext(cfg, rank, world, comm, ranges[rank], len(positions), id_base, track_ids)
    ctx.set_boundary(boundary)
    mine = own[rank]
    vel = None if velocities is None else np.ascontiguousarray(velocities, np.float32).reshape(-1, 2)[mine]
    ctx.upload_particles(positions[mine], vel)
    return ctx, ranges, np.concatenate(own)
