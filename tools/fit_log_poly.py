"""Coefficients Lg1..Lg7 of csrc/kem_math.cuh::log.

log(1+f) = 2s + s R(z),  s = f/(2+f),  z = s^2,  R(z) = z (Lg1 + Lg2 z + ... + Lg7 z^6):
Chebyshev-node interpolation of R(z)/z on [0, 1.02 * 0.1716^2] (|s| <= 0.1716 for the
mantissa range [sqrt(1/2), sqrt(2))), 60-digit arithmetic.  Prints hex-float literals and
the maximum error contribution relative to log(1+f) (6.2e-18).
"""
import mpmath as mp

mp.mp.dps = 60
zmax = (mp.mpf("0.1716") ** 2) * mp.mpf("1.02")
n = 7
nodes = [zmax / 2 * (1 + mp.cos(mp.pi * (2 * k + 1) / (2 * n))) for k in range(n)]


def target(z):
    s = mp.sqrt(z)
    return (mp.log((1 + s) / (1 - s)) / s - 2) / z


A, b = mp.matrix(n, n), mp.matrix(n, 1)
for i, z in enumerate(nodes):
    for j in range(n):
        A[i, j] = z ** j
    b[i] = target(z)
coef = [float(x) for x in mp.lu_solve(A, b)]
worst = 0
for k in range(1, 4001):
    z = zmax * k / 4000
    p = sum(mp.mpf(coef[j]) * z ** j for j in range(n)) * z
    worst = max(worst, abs(p - target(z) * z) / 2)
for j, cj in enumerate(coef):
    print(f"Lg{j + 1} = {cj!r:24s} {cj.hex()}")
print("max error relative to log(1+f):", mp.nstr(worst, 5))
