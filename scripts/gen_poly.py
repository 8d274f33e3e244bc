#!/usr/bin/env python3
"""Derives the polynomial coefficients used by option-pricing-ffn-lbfgs_b200/csrc/dhj_fastmath.cuh with
mpmath (60 digits): near-minimax fits by Chebyshev-node interpolation followed by a few Remez exchanges.
Prints C initialisers and the achieved maximum error.  Run: python scripts/gen_poly.py"""
import mpmath as mp

mp.mp.dps = 60


def remez(f, a, b, deg, weight=None, iters=12):
    """Minimax polynomial sum c_i x^i for f on [a,b] (absolute error, optionally weighted)."""
    w = weight or (lambda x: mp.mpf(1))
    n = deg + 2
    xs = [(a + b) / 2 + (b - a) / 2 * mp.cos(mp.pi * (n - 1 - i) / (n - 1)) for i in range(n)]
    for _ in range(iters):
        A = mp.matrix(n, n)
        rhs = mp.matrix(n, 1)
        for i, x in enumerate(xs):
            for j in range(deg + 1):
                A[i, j] = x ** j
            A[i, deg + 1] = (-1) ** i / w(x)
            rhs[i] = f(x)
        sol = mp.lu_solve(A, rhs)
        c = [sol[j] for j in range(deg + 1)]
        err = lambda x: (mp.polyval(c[::-1], x) - f(x)) * w(x)
        # new extrema: sample densely, pick local maxima of |err| between sign changes
        grid = [a + (b - a) * mp.mpf(i) / 4000 for i in range(4001)]
        vals = [err(x) for x in grid]
        ext = []
        i = 0
        while i < len(grid):
            j = i
            while j + 1 < len(grid) and (vals[j + 1] > 0) == (vals[i] > 0):
                j += 1
            k = max(range(i, j + 1), key=lambda t: abs(vals[t]))
            ext.append(grid[k])
            i = j + 1
        if len(ext) != n:
            break
        xs = ext
    emax = max(abs(v) for v in vals)
    return c, emax


def show(name, c, emax):
    print(f"// {name}: max error {mp.nstr(emax, 5)}")
    print("{" + ", ".join(mp.nstr(x, 20) for x in c) + "}")


if __name__ == "__main__":
    # exp(r) = 1 + r + r^2 * q(r), |r| <= ln2/2 ; q of degree 9 (total degree 11)
    L = mp.log(2) / 2
    q = lambda r: (mp.exp(r) - 1 - r) / r ** 2 if abs(r) > mp.mpf('1e-15') else mp.mpf(1) / 2 + r / 6 + r * r / 24
    c, e = remez(q, -L, L, 9)
    show("exp q(r) deg 9 (abs error of q; times r^2 <= 0.12)", c, e)
    c, e = remez(q, -L, L, 10)
    show("exp q(r) deg 10", c, e)
