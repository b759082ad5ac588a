"""yasph2d_b200 -- B200-native (sm_100a) implementation of yasph2d's per-step SPH hot path.

The product is libyasph_gpu.so (CUDA, C ABI in include/yasph_gpu.h).  This package holds its sources (csrc/), the ctypes
binding (_capi) and the host-side mirror of the reference's Solver / FluidParticleWorld / NeighborhoodSearch surface (host).
"""
from . import _capi as capi  # noqa: F401
from . import stateio  # noqa: F401
from .host import (  # noqa: F401
    ConstantFluidProperties,
    DFSPHSolver,
    FluidParticleWorld,
    GpuContext,
    NeighborhoodSearch,
    NeighborLists,
    Particles,
    PhysicalViscosityModel,
    Rect,
    SimulationStepConfig,
    Solver,
    TimeManager,
    WCSPHSolver,
    XSPHViscosityModel,
    dam_break_scene,
    tank_scene,
)
