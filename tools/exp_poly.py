"""Coefficients of kExpPoly in csrc/pmvs_device.cuh: degree-11 interpolant of exp at the Chebyshev nodes of
[-ln2/2, ln2/2] (slightly widened), computed with mpmath at 60 digits and rounded to double."""
import mpmath as mp

mp.mp.dps = 60
a = mp.log(2) / 2 * mp.mpf("1.0001")
n = 11
nodes = [a * mp.cos(mp.pi * (2 * k + 1) / (2 * (n + 1))) for k in range(n + 1)]
V = mp.matrix(n + 1, n + 1)
for i, x in enumerate(nodes):
    for j in range(n + 1):
        V[i, j] = x ** j
c = mp.lu_solve(V, mp.matrix([mp.e ** x for x in nodes]))
coef = [float(c[j]) for j in range(n + 1)]
xs = [-a + 2 * a * k / 2000 for k in range(2001)]
err = max(abs(sum(mp.mpf(coef[j]) * x ** j for j in range(n + 1)) - mp.e ** x) / mp.e ** x for x in xs)
print("max relative error with rounded coefficients:", mp.nstr(err, 5))
for j, cj in enumerate(coef):
    print(j, cj.hex())
