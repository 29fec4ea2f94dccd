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


def diffeqsolve(*a, **k):
    raise NotImplementedError("refshim: diffrax.diffeqsolve is third-party; the oracle restates it")
