"""diffrax is third-party and absent: only the names the reference's modules touch at import / construction time."""


class _Opaque:
    def __init__(self, *a, **k):
        self.args, self.kwargs = a, k


class Euler(_Opaque): pass
class Dopri5(_Opaque): pass
class Tsit5(_Opaque): pass
class ConstantStepSize(_Opaque): pass
class PIDController(_Opaque): pass
class SaveAt(_Opaque): pass
class Event(_Opaque): pass
class ODETerm(_Opaque): pass
class ControlTerm(_Opaque): pass
class MultiTerm(_Opaque): pass
class VirtualBrownianTree(_Opaque): pass
class UnsafeBrownianPath(_Opaque): pass
class EulerHeun(_Opaque): pass


class DirectAdjoint(_Opaque): pass
class AbstractStepSizeController(_Opaque): pass
class AbstractSolver(_Opaque): pass
class SpaceTimeLevyArea(_Opaque): pass


class _Solution:
    def __init__(self, ts, ys):
        self.ts, self.ys = ts, ys


# diffrax.diffeqsolve is third-party and is NOT emulated.  The recorder INJECTS the saved states a solve would
# return (taken from the oracle's own solve) so that the reference's code AROUND the solve -- the arguments it
# passes, the post-pass over the save points, the fitness function -- runs unmodified; CALLS logs what it passed.
INJECT = []
CALLS = []


def diffeqsolve(terms, solver, t0, t1, dt0, y0, saveat=None, args=None, **kw):
    import numpy as _np
    import jax.numpy as jnp
    if not INJECT:
        raise NotImplementedError("refshim: diffrax.diffeqsolve is third-party; nothing was injected")
    CALLS.append(dict(t0=float(t0), t1=float(t1), dt0=float(dt0), y0=_np.asarray(y0, dtype=_np.float32).copy(),
                      max_steps=kw.get("max_steps"), throw=kw.get("throw"), has_event=kw.get("event") is not None,
                      vector_field=getattr(terms, "args", (None,))[0]))
    ts = saveat.kwargs.get("ts") if saveat is not None else None
    return _Solution(ts, jnp.asarray(INJECT.pop(0)))
