"""Stand-ins for the diffrax objects the reference evaluators take as constructor arguments
(``solver=diffrax.Euler()``, ``stepsize_controller=diffrax.ConstantStepSize()``; the notebooks pass
``diffrax.Dopri5()`` and ``diffrax.PIDController(atol, rtol, dtmin=0.001)``).  Real diffrax objects
are accepted too (matched by class name and attributes); these classes exist so that user code
runs where diffrax is not installed.  They carry settings only -- the integration itself is the
CUDA rollout kernel.
"""
from __future__ import annotations

from dataclasses import dataclass

from . import abi


class Euler:
    pass


class Dopri5:
    pass


class ConstantStepSize:
    pass


@dataclass
class PIDController:
    """diffrax.PIDController: defaults are an I-controller (pcoeff 0, icoeff 1, dcoeff 0)."""
    rtol: float
    atol: float
    pcoeff: float = 0.0
    icoeff: float = 1.0
    dcoeff: float = 0.0
    dtmin: float | None = None
    dtmax: float | None = None
    force_dtmin: bool = True
    factormin: float = 0.2
    factormax: float = 10.0
    safety: float = 0.9


def solver_settings(solver, stepsize_controller) -> dict:
    """Maps (solver, stepsize_controller) objects or names to RolloutData keyword arguments."""
    def name(o):
        return (o if isinstance(o, str) else type(o).__name__).lower()

    out = {}
    s = "euler" if solver is None else name(solver)
    if s == "euler":
        out["solver"] = abi.SOLVER_EULER
    elif s == "dopri5":
        out["solver"] = abi.SOLVER_DOPRI5
    else:
        raise NotImplementedError(f"solver {s!r}: the CUDA rollout kernels implement Euler and Dopri5 (no CPU fallback)")
    c = "constantstepsize" if stepsize_controller is None else name(stepsize_controller)
    if c in ("constantstepsize", "constant"):
        out["controller"] = abi.CONTROLLER_CONSTANT
    elif c == "pidcontroller":
        pc = stepsize_controller
        if out["solver"] != abi.SOLVER_DOPRI5:
            raise ValueError("PIDController needs a solver with an embedded error estimate (Dopri5); diffrax raises too")
        if getattr(pc, "dtmax", None) is not None:
            raise NotImplementedError("PIDController(dtmax=...) is not supported")
        if not getattr(pc, "force_dtmin", True):
            raise NotImplementedError("PIDController(force_dtmin=False) is not supported")
        out.update(controller=abi.CONTROLLER_PID, rtol=float(pc.rtol), atol=float(pc.atol),
                   dtmin=float(pc.dtmin) if getattr(pc, "dtmin", None) is not None else 0.0,
                   pid=(float(getattr(pc, "pcoeff", 0.0)), float(getattr(pc, "icoeff", 1.0)),
                        float(getattr(pc, "dcoeff", 0.0))),
                   pid_limits=(float(getattr(pc, "safety", 0.9)), float(getattr(pc, "factormin", 0.2)),
                               float(getattr(pc, "factormax", 10.0))))
    else:
        raise NotImplementedError(f"step-size controller {c!r} is not supported")
    return out
