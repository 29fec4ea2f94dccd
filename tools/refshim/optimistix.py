"""optimistix is third-party and absent: the names SR_evaluator.py touches (a root finder that is constructed, never used)."""


class Newton:
    def __init__(self, *a, **k):
        pass


def rms_norm(x):
    raise NotImplementedError
