"""Coefficients of the degree-11 polynomial in csrc/kem_math.cuh::exp.

Chebyshev-node interpolation (near-minimax) of q(r) = (e^r - 1 - r)/r^2 by a degree-9
polynomial on [-ln2/2, ln2/2] (with a 1e-4 margin for the rounding of k), in 60-digit
arithmetic; exp(r) ~ 1 + r + r^2 q(r), so c0 = c1 = 1 exactly.  Prints the coefficients
as C hex-float literals and the maximum relative error of the double-rounded polynomial
evaluated exactly (1.6e-17, i.e. 0.14 ulp, before the rounding of the Horner steps).
"""
import mpmath as mp

mp.mp.dps = 60
a = mp.log(2) / 2 * mp.mpf("1.0001")
n = 10
nodes = [a * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
A, b = mp.matrix(n, n), mp.matrix(n, 1)
for i, x in enumerate(nodes):
    for j in range(n):
        A[i, j] = x ** j
    b[i] = (mp.e ** x - 1 - x) / x ** 2
c = mp.lu_solve(A, b)
coef = [1.0, 1.0] + [float(c[j]) for j in range(n)]
worst = 0
for k in range(-3000, 3001):
    x = a * k / 3000
    if x == 0:
        continue
    p = sum(mp.mpf(coef[j]) * x ** j for j in range(len(coef)))
    worst = max(worst, abs(p - mp.e ** x) / mp.e ** x)
for j, cj in enumerate(coef):
    print(f"c{j:<2d} = {cj!r:26s} {float(cj).hex()}")
print("max relative error (exact evaluation):", mp.nstr(worst, 5), " 2^-53 =", 2.0 ** -53)
