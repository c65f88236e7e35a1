"""Hand-derived ordinary-kriging fixture: exact rational arithmetic, no pykrige restatement involved.

Test infrastructure (see oracle/__init__.py).  Every OK number elsewhere in tests/golden/ was produced by
oracle/pykrige_restated.py - the sibling of the code under test.  This one is not: sample and target points sit on a
3-4-5 lattice so that every distance is rational, the variogram is linear with rational slope and nugget, and the
kriging system of the published algorithm (pykrige ok.py: A = [[-gamma(D), 1], [1', 0]] with zero diagonal,
b = [-gamma(d); 1] with b_i = 0 where d_i = 0, x = A^-1 b, z = sum x_i Z_i, sigma^2 = -sum x_i b_i over all n + 1
entries) is solved with ``fractions.Fraction`` by Gauss-Jordan elimination.  The committed JSON holds the exact
fractions and their float values.

    python oracle/make_hand_fixture.py        # writes tests/golden/ok_hand_derived.json
"""
from __future__ import annotations

import json
import os
from fractions import Fraction as F
from math import isqrt


def rational_distance(p, q) -> F:
    dx, dy = p[0] - q[0], p[1] - q[1]
    d2 = dx * dx + dy * dy
    num, den = d2.numerator, d2.denominator
    rn, rd = isqrt(num), isqrt(den)
    if rn * rn != num or rd * rd != den:
        raise ValueError(f"distance between {p} and {q} is irrational")
    return F(rn, rd)


def solve(a, b):
    """Gauss-Jordan with exact pivots (first non-zero entry of the column)."""
    n = len(a)
    m = [row[:] + [b[i]] for i, row in enumerate(a)]
    for c in range(n):
        p = next(r for r in range(c, n) if m[r][c] != 0)
        m[c], m[p] = m[p], m[c]
        piv = m[c][c]
        m[c] = [v / piv for v in m[c]]
        for r in range(n):
            if r != c and m[r][c] != 0:
                f = m[r][c]
                m[r] = [vr - f * vc for vr, vc in zip(m[r], m[c])]
    return [m[i][n] for i in range(n)]


def krige(samples, values, target, slope: F, nugget: F):
    n = len(samples)
    gamma = lambda h: slope * h + nugget  # noqa: E731  (pykrige linear_variogram_model: m[0] * d + m[1])
    a = [[F(0)] * (n + 1) for _ in range(n + 1)]
    for i in range(n):
        for j in range(n):
            if i != j:
                a[i][j] = -gamma(rational_distance(samples[i], samples[j]))
        a[i][n] = a[n][i] = F(1)
    b = []
    for i in range(n):
        d = rational_distance(samples[i], target)
        b.append(F(0) if d == 0 else -gamma(d))  # exact_values rule
    b.append(F(1))
    x = solve(a, b)
    z = sum(x[i] * values[i] for i in range(n))
    sigmasq = sum(x[i] * -b[i] for i in range(n + 1))
    return x, z, sigmasq


def main():
    cases = []
    configs = [
        ("triangle_3_4_5", [(F(0), F(0)), (F(3), F(0)), (F(0), F(4))], [F(1), F(2), F(-1, 2)],
         [(F(3), F(4)), (F(0), F(0)), (F(3, 2), F(2))], F(1, 2), F(1, 4)),
        ("rectangle_3x4", [(F(0), F(0)), (F(3), F(0)), (F(0), F(4)), (F(3), F(4))], [F(1), F(2), F(-1, 2), F(3)],
         [(F(3, 2), F(2)), (F(3), F(4)), (F(3), F(8)) if False else (F(0), F(4))], F(2, 3), F(0)),
    ]
    for name, samples, values, targets, slope, nugget in configs:
        out = []
        for t in targets:
            x, z, s2 = krige(samples, values, t, slope, nugget)
            out.append({"target": [float(t[0]), float(t[1])], "weights_exact": [str(v) for v in x[:-1]], "mu_exact": str(x[-1]),
                        "z_exact": str(z), "sigmasq_exact": str(s2), "z": float(z), "sigmasq": float(s2),
                        "weights": [float(v) for v in x[:-1]]})
            assert sum(x[:-1]) == 1
        cases.append({"name": name, "x": [float(p[0]) for p in samples], "y": [float(p[1]) for p in samples],
                      "values": [float(v) for v in values], "variogram_model": "linear",
                      "variogram_parameters": [float(slope), float(nugget)], "slope_exact": str(slope), "nugget_exact": str(nugget),
                      "targets": out})
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ok_hand_derived.json")
    with open(path, "w") as f:
        json.dump({"source": "oracle/make_hand_fixture.py: exact rational Gauss-Jordan of the published pykrige system; no restated code",
                   "cases": cases}, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
