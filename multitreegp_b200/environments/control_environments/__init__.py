from .acrobot import Acrobot
from .harmonic_oscillator import HarmonicOscillator
from .reactor import StirredTankReactor

__all__ = ["Acrobot", "HarmonicOscillator", "StirredTankReactor"]
