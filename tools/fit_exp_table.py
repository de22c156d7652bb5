"""Table and polynomial of the table-assisted exp in csrc/kem_math.cuh.

exp(x) = 2^k * T[j] * exp(r),  x = (256 k + j) ln2/256 + r,  |r| <= ln2/512,
T[j] = 2^(j/256) rounded to double, exp(r) = 1 + r + r^2 q(r) with q of degree 2
(Chebyshev-node interpolation of (e^r - 1 - r)/r^2 in 60-digit arithmetic).
Prints the C initialisers and the maximum relative error of T[j] * (1 + r + r^2 q(r))
evaluated exactly with the double-rounded table and coefficients.
"""
import mpmath as mp

mp.mp.dps = 60
J = 256
a = mp.log(2) / (2 * J) * mp.mpf("1.001")
n = 3
nodes = [a * mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
A, b = mp.matrix(n, n), mp.matrix(n, 1)
for i, x in enumerate(nodes):
    for j in range(n):
        A[i, j] = x ** j
    b[i] = (mp.e ** x - 1 - x) / x ** 2 if abs(x) > mp.mpf(10) ** -30 else mp.mpf(1) / 2
c = [float(v) for v in mp.lu_solve(A, b)]
table = [float(mp.mpf(2) ** (mp.mpf(j) / J)) for j in range(J)]
worst_poly = 0
for k in range(-2000, 2001):
    x = a * k / 2000
    if x == 0:
        continue
    p = 1 + x + x * x * sum(mp.mpf(c[j]) * x ** j for j in range(n))
    worst_poly = max(worst_poly, abs(p - mp.e ** x) / mp.e ** x)
worst_tab = max(abs(mp.mpf(table[j]) - mp.mpf(2) ** (mp.mpf(j) / J)) / mp.mpf(2) ** (mp.mpf(j) / J) for j in range(J))
inv = float(J / mp.log(2))
hi = float(mp.log(2) / J)
lo = float(mp.log(2) / J - mp.mpf(hi))
print(f"J/ln2      = {inv.hex()}")
print(f"-ln2/J hi  = {(-hi).hex()}")
print(f"-ln2/J lo  = {(-lo).hex()}")
for j, cj in enumerate(c):
    print(f"c{j + 2} = {cj!r:26s} {cj.hex()}")
print("polynomial: max relative error", mp.nstr(worst_poly, 4), "; table rounding", mp.nstr(worst_tab, 4),
      "; 2^-53 =", 2.0 ** -53)
print("table:")
for j in range(0, J, 4):
    print("    " + ", ".join(table[i].hex() for i in range(j, j + 4)) + ",")
