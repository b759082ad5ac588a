"""Scene / solver-state files and per-step trajectory logs.

The reference keeps everything in memory (`Particles`, fluidparticleworld.rs:11-23; solver arrays dfsph.rs:36-41,
wscsph.rs:22) and has no file format.  These two formats exist so that a run of the oracle, of the GPU path or of an
external build of the Rust reference can be compared offline, and so that a GPU run can be checkpointed and resumed
bit-exactly (`GpuContext.solver_state`, `yasph_solver_state_set`, `yasph_upload_field`).

State file (little endian):   b"YSPH2D01" | u32 header_bytes | header (UTF-8 JSON) | arrays
    header = {"params": {...}, "arrays": [{"name", "dtype" ("f4"), "shape"} ...] in file order, "solver": {...}}
    arrays = raw C-order data, each padded to a multiple of 8 bytes.
    Well-known arrays: positions (N,2), velocities (N,2), boundary (M,2), densities (N), kappa (N), stiffness (N),
    accelerations (N,2) -- the first three are what `add_fluid_rect` / `add_boundary_*` produce (fluidparticleworld.rs:140-195).
Trajectory log: JSON lines, one object per simulation step (`TrajectoryRecorder.FIELDS`), first line = {"header": ...}.

Pure host code: nothing here touches the GPU library unless a GpuContext is passed in.
"""
import json
import struct

import numpy as np

from . import _capi as capi

MAGIC = b"YSPH2D01"


def save_state(path, arrays, params=None, solver=None):
    """arrays: dict name -> float32 ndarray; params / solver: JSON-serialisable dicts."""
    metas, blobs = [], []
    for name, a in arrays.items():
        a = np.ascontiguousarray(a, np.float32)
        metas.append({"name": name, "dtype": "f4", "shape": list(a.shape)})
        b = a.tobytes()
        blobs.append(b + b"\0" * (-len(b) % 8))
    header = json.dumps({"params": params or {}, "solver": solver or {}, "arrays": metas}).encode("utf-8")
    header += b" " * (-len(header) % 8)
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(header)))
        f.write(b"\0\0\0\0")
        f.write(header)
        for b in blobs:
            f.write(b)


def load_state(path):
    """Returns (arrays, params, solver)."""
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError("%s: not a yasph2d state file" % path)
        (hb,) = struct.unpack("<I", f.read(4))
        f.read(4)
        header = json.loads(f.read(hb).decode("utf-8"))
        arrays = {}
        for m in header["arrays"]:
            if m["dtype"] != "f4":
                raise ValueError("%s: unsupported dtype %s" % (path, m["dtype"]))
            count = int(np.prod(m["shape"])) if m["shape"] else 1
            nbytes = 4 * count
            raw = f.read(nbytes + (-nbytes % 8))
            if len(raw) < nbytes:
                raise ValueError("%s: truncated array %s" % (path, m["name"]))
            arrays[m["name"]] = np.frombuffer(raw[:nbytes], np.float32).reshape(m["shape"]).copy()
    return arrays, header["params"], header["solver"]


def checkpoint(ctx, path, boundary=None, params=None):
    """Everything a GpuContext needs to continue bit-exactly: sorted positions, velocities, the solver's carried arrays
    (DFSPH warm starts / WCSPH accelerations), iteration counts, the current time step and the total simulated time (the
    TargetFrameLength rule of timemanager.rs:268-272 depends on it)."""
    pos, vel, dens = ctx.download_particles()
    st = ctx.solver_state()
    arrays = {"positions": pos, "velocities": vel, "densities": dens}
    if ctx.cfg.solver == capi.SOLVER_DFSPH:
        if st.initialized:
            arrays["kappa"] = ctx.field(capi.FIELD_KAPPA)
            arrays["stiffness"] = ctx.field(capi.FIELD_STIFFNESS)
    else:
        arrays["accelerations"] = ctx.field(capi.FIELD_ACCELERATION)
    arrays["boundary"] = ctx.field(capi.FIELD_BOUNDARY) if boundary is None else np.asarray(boundary, np.float32)
    solver = {"kind": int(ctx.cfg.solver), "step_ns": int(st.step_ns), "iters_density": int(st.iters_density),
              "iters_divergence": int(st.iters_divergence), "initialized": int(st.initialized), "total_simulated_ns": int(st.total_simulated_ns)}
    save_state(path, arrays, params, solver)


def resume(ctx, path):
    """Loads a checkpoint into a fresh GpuContext of the same configuration; returns (arrays, params, solver)."""
    arrays, params, solver = load_state(path)
    if int(solver.get("kind", ctx.cfg.solver)) != int(ctx.cfg.solver):
        raise ValueError("checkpoint of solver kind %s loaded into a context of kind %s" % (solver.get("kind"), ctx.cfg.solver))
    ctx.set_boundary(arrays["boundary"])
    ctx.upload_particles(arrays["positions"], arrays["velocities"])
    ctx.set_solver_state(solver["step_ns"], solver["iters_density"], solver["iters_divergence"], bool(solver["initialized"]),
                         solver.get("total_simulated_ns", 0))
    if "kappa" in arrays:
        ctx.upload_field(capi.FIELD_KAPPA, arrays["kappa"])
        ctx.upload_field(capi.FIELD_STIFFNESS, arrays["stiffness"])
    if "accelerations" in arrays:
        ctx.upload_field(capi.FIELD_ACCELERATION, arrays["accelerations"])
    return arrays, params, solver


class TrajectoryRecorder:
    """Per-step scalars of a run (BASELINE north star: iteration counts, density-error and kinetic-energy trajectories)."""

    FIELDS = ("step", "time_ns", "dt_ns", "iters_density", "iters_divergence", "avg_density_error", "avg_divergence", "max_velocity",
              "kinetic_energy")

    def __init__(self, path=None, header=None, particle_mass=None):
        self.rows, self.path, self.mass, self.time_ns = [], path, particle_mass, 0
        self._f = open(path, "w") if path else None
        if self._f:
            self._f.write(json.dumps({"header": header or {}, "fields": list(self.FIELDS)}) + "\n")

    def record(self, report, velocities=None):
        """report: any object with the step-report fields (GPU `yasph_step_report` or the oracle's); velocities: optional
        (N,2) array -> kinetic energy sum 1/2 m |v|^2 accumulated in float64."""
        self.time_ns += int(report.dt_ns)
        ek = None
        if velocities is not None and self.mass is not None:
            v = np.asarray(velocities, np.float64)
            ek = 0.5 * float(self.mass) * float((v * v).sum())
        row = {"step": len(self.rows), "time_ns": self.time_ns, "dt_ns": int(report.dt_ns), "iters_density": int(report.iters_density),
               "iters_divergence": int(report.iters_divergence), "avg_density_error": float(report.avg_density_error),
               "avg_divergence": float(report.avg_divergence), "max_velocity": float(report.max_velocity), "kinetic_energy": ek}
        self.rows.append(row)
        if self._f:
            self._f.write(json.dumps(row) + "\n")
        return row

    def close(self):
        if self._f:
            self._f.close()
            self._f = None


def load_trajectory(path):
    with open(path) as f:
        lines = [json.loads(l) for l in f if l.strip()]
    return lines[0], lines[1:]


def compare_trajectories(a, b, rel=1e-2, iters_slack=1, abs_tol=None):
    """The north-star bars between two runs (lists of rows): iteration counts within `iters_slack` per solve, density-error and
    kinetic-energy trajectories within `rel` (plus an absolute floor per field, default 1e-3 density units -- 1e-5 of the rest
    density of the application's fluid -- and 0 for the energy).  Returns a list of violations (empty = within bounds)."""
    abs_tol = {"avg_density_error": 1e-3, "kinetic_energy": 0.0, **(abs_tol or {})}
    bad = []
    for ra, rb in zip(a, b):
        s = ra["step"]
        if abs(ra["iters_density"] - rb["iters_density"]) > iters_slack or abs(ra["iters_divergence"] - rb["iters_divergence"]) > iters_slack:
            bad.append((s, "iterations", (ra["iters_density"], ra["iters_divergence"]), (rb["iters_density"], rb["iters_divergence"])))
        for k in ("avg_density_error", "kinetic_energy"):
            x, y = ra.get(k), rb.get(k)
            if x is None or y is None:
                continue
            if abs(x - y) > rel * max(abs(x), abs(y)) + abs_tol.get(k, 0.0):
                bad.append((s, k, x, y))
    if len(a) != len(b):
        bad.append((min(len(a), len(b)), "length", len(a), len(b)))
    return bad
