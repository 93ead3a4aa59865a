"""B200-native Euler residual + RK4: a drop-in for PDESolver.jl's hot path.

``evalResidual`` / ``rk4`` / ``EulerData`` mirror the reference's physics-module
API (src/solver/euler/euler.jl:111-175, src/NonlinearSolvers/rk4.jl:404-410) on
top of the C ABI in ``include/pdes_euler_b200.h``; ``sbp`` and ``mesh`` are
host-side stand-ins for the un-vendored SummationByParts.jl / PumiInterface.jl
inputs (synthetic structured meshes, SBP operators).
"""
from . import dump, mesh, sbp  # noqa: F401
from .euler import (calcEnstrophy, calcEntropyIntegral, calcKineticEnergy, calcKineticEnergydt, contractResEntropyVars,  # noqa: F401
                    diagnostics, integrateQ)
from .euler import (EulerData, ParamType, PDESolverError, PhysicsError,  # noqa: F401
                    createObjects, evaldRdqProduct, evalResidual, linearSolve, lserk54, newton, rk4)
from .mesh import structured_mesh, two_element_mesh  # noqa: F401
from .sbp import build_operator  # noqa: F401

__all__ = ["sbp", "mesh", "EulerData", "ParamType", "PDESolverError", "PhysicsError", "createObjects",
           "evalResidual", "evaldRdqProduct", "rk4", "lserk54", "newton", "linearSolve", "diagnostics", "calcEntropyIntegral", "contractResEntropyVars", "integrateQ",
           "calcKineticEnergy", "calcKineticEnergydt", "calcEnstrophy", "structured_mesh", "two_element_mesh", "build_operator"]
