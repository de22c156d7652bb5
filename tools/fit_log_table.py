"""Table and polynomial of the table-assisted log in csrc/kem_math.cuh.

x = 2^k m with m in [0.70703125, 1.4140625) (the halving threshold sits on a table boundary,
mantissa 0x6a000, instead of exactly sqrt 2).  With t = (high word of x) - 0x3fe6a000 the
exponent is k = t >> 20 and the table index j = (t >> 12) & 255, i.e. the entries are ordered by
m, from 0.70703125 (j = 0) through 1 (j = 150) to 1.4140625; entry j holds (invc_j, logc_j): invc_j ~ 1/centre of the interval, logc_j = -log(invc_j) of the ROUNDED invc_j,
so that log(m) = logc_j + log1p(r) holds exactly for r = m invc_j - 1 (one FMA, exact up to its
own rounding).  The two intervals next to 1 (j = 149 and j = 150) use invc = 1, logc = 0: r = m - 1,
which keeps the result accurate relative to itself as x -> 1.  |r| <= 2^-8.
log1p(r) = r - r^2/2 + r^3 q(r), q of degree 4 fitted at Chebyshev nodes in 60-digit arithmetic.
Prints the C initialisers and the worst error of the polynomial relative to log1p(r).
"""
import mpmath as mp

mp.mp.dps = 60
N, FIRST = 256, 0x6a          # entry j covers the mantissa bits (j + FIRST) mod 256; j < N - FIRST: halved
a = mp.mpf(2) ** -8 * mp.mpf("1.002")
n = 5
nodes = [a * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]


def q_exact(r):
    if abs(r) < mp.mpf(10) ** -25:
        return mp.mpf(1) / 3
    return (mp.log1p(r) - r + r * r / 2) / r ** 3


A, b = mp.matrix(n, n), mp.matrix(n, 1)
for i, x in enumerate(nodes):
    for j in range(n):
        A[i, j] = x ** j
    b[i] = q_exact(x)
c = [float(v) for v in mp.lu_solve(A, b)]
worst = 0
for k in range(-3000, 3001):
    r = a * k / 3000
    if r == 0:
        continue
    p = r - r * r / 2 + r ** 3 * sum(mp.mpf(c[j]) * r ** j for j in range(n))
    worst = max(worst, abs(p - mp.log1p(r)) / abs(mp.log1p(r)))
rows = []
for j in range(N):
    b = (j + FIRST) % N
    lo, hi = 1 + mp.mpf(b) / N, 1 + mp.mpf(b + 1) / N
    if j < N - FIRST:
        lo, hi = lo / 2, hi / 2
    if b in (0, N - 1):
        invc, logc = 1.0, 0.0
    else:
        invc = float(1 / ((lo + hi) / 2))
        logc = float(-mp.log(mp.mpf(invc)))
    rmax = max(abs(lo * mp.mpf(invc) - 1), abs(hi * mp.mpf(invc) - 1))
    assert rmax <= a, (j, rmax)
    rows.append((invc, logc))
for j, cj in enumerate(c):
    print(f"q{j} = {cj!r:26s} {cj.hex()}")
print("polynomial: max error relative to log1p(r):", mp.nstr(worst, 4), "; 2^-53 =", 2.0 ** -53)
print("table (invc, logc):")
for j in range(0, N, 2):
    print("    " + ", ".join(f"{{{rows[i][0].hex()}, {rows[i][1].hex()}}}" for i in range(j, j + 2)) + ",")
