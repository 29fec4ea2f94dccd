"""Generates the constants used by include/mtgp_math.h (run once; output pasted into the header).

Everything is derived from first principles with mpmath (no table is copied from any library):
  * 2/pi bits for the Payne-Hanek reduction,
  * pi/2 three-way float split for the Cody-Waite reduction,
  * near-minimax (Chebyshev) polynomial coefficients for sin, cos, exp, log, tanh kernels.
"""
import mpmath as mp
import numpy as np

mp.mp.prec = 400


def f32(x):
    return float(np.float32(float(x)))


def hexf(x):
    return float(np.float32(x)).hex()


def fit(fn, a, b, n):
    c = mp.chebyfit(fn, [a, b], n)  # highest power first
    return [f32(v) for v in c]


print("// 2/pi, 32-bit words, most significant first")
x = 2 / mp.pi
words = []
for i in range(8):
    x = x * (1 << 32)
    w = int(mp.floor(x))
    x -= w
    words.append(w)
print(", ".join("0x%08Xu" % w for w in words))

hi = mp.mpf(f32(mp.pi / 2))
mid = mp.mpf(f32(mp.pi / 2 - hi))
lo = mp.mpf(f32(mp.pi / 2 - hi - mid))
print("pio2 hi/mid/lo", hexf(hi), hexf(mid), hexf(lo), float(hi), float(mid), float(lo))
print("2/pi", hexf(2 / mp.pi), "pi/2 double", repr(float(mp.pi / 2)))

lim = (mp.pi / 4 + mp.mpf("0.02")) ** 2
eps = mp.mpf("1e-40")
S = fit(lambda s: (mp.sin(mp.sqrt(s + eps)) / mp.sqrt(s + eps) - 1) / (s + eps), 0, lim, 4)
print("sin P(s) hi..lo", [hexf(v) for v in S], S)
C = fit(lambda s: (mp.cos(mp.sqrt(s + eps)) - 1 + (s + eps) / 2) / (s + eps) ** 2, 0, lim, 4)
print("cos Q(s) hi..lo", [hexf(v) for v in C], C)

l2 = mp.log(2)
a = l2 / 2 + mp.mpf("0.01")
def eexp(r):
    if abs(r) < mp.mpf("1e-30"):
        return mp.mpf(1) / 2
    return (mp.exp(r) - 1 - r) / r ** 2
E = fit(eexp, -a, a, 6)
print("exp E(r) hi..lo", [hexf(v) for v in E], E)
l2hi = mp.mpf(float(np.float32(float(l2)).view(np.uint32) & np.uint32(0xFFFFF000)).hex() if False else 0)
# ln2 split: hi has 12 trailing zero bits so n*hi is exact for |n| < 2^12
bits = np.float32(float(l2)).view(np.uint32) & np.uint32(0xFFFFF000)
l2hi = mp.mpf(float(bits.view(np.float32)))
l2lo = mp.mpf(f32(l2 - l2hi))
print("ln2 hi/lo", hexf(l2hi), hexf(l2lo), "log2e", hexf(1 / l2))

# log: R(z) with log((1+s)/(1-s)) = 2s + s*R(z), z = s^2, s in [-0.1716, 0.1716]
zmax = (mp.mpf("0.1716") + mp.mpf("0.002")) ** 2
def rlog(z):
    z = z + eps
    s = mp.sqrt(z)
    return (mp.log((1 + s) / (1 - s)) - 2 * s) / s
# R(z) = z*(2/3 + 2/5 z + ...): fit R(z)/z
L = fit(lambda z: rlog(z) / (z + eps), 0, zmax, 4)
print("log Lg(z) hi..lo", [hexf(v) for v in L], L)

# tanh small: tanh(x) = x + x*s*T(s), s=x^2, |x| <= 0.55
tl = mp.mpf("0.56") ** 2
T = fit(lambda s: (mp.tanh(mp.sqrt(s + eps)) / mp.sqrt(s + eps) - 1) / (s + eps), 0, tl, 5)
print("tanh T(s) hi..lo", [hexf(v) for v in T], T)
print("sqrt2", hexf(mp.sqrt(2)), "sqrt(1/2)", hexf(mp.sqrt(mp.mpf(1) / 2)))
