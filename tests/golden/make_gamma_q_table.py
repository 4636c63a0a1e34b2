"""Generates tests/golden/gamma_q_table.json: the regularized upper incomplete gamma function Q(a, x) at 40 significant
digits (mpmath) over the (a, x) range Spr_study reaches for its above-root region (core/spr_study.cpp:334-369,
core/safe_gamma_math.h:46-96): a = f m + 1 with f = 0.8 (core/subrun.cpp:511) and m = 0..80 minimum mutations, plus
a few non-0.8 annealing factors; x = lambda_X f s from 1e-2 (the power-law switch-over, core/spr_study.cpp:347) to the
point where Q underflows.  The reference takes these values from Boost.Math 1.84 (not available offline), so this table is
the absolute pin for oracle/gamma_q.h and for the device's dev_gamma_q.

    python tests/golden/make_gamma_q_table.py
"""
import json
import os

import mpmath

mpmath.mp.dps = 40


def main():
    a_values = sorted({0.8 * m + 1 for m in list(range(0, 31)) + [40, 50, 60, 80]} | {1.0, 1.5, 2.0, 3.7, 10.0, 25.0})
    rel = [1e-2, 0.03, 0.1, 0.3, 0.5, 0.8, 0.95, 1.0, 1.05, 1.2, 1.5, 2.0, 3.0, 5.0, 10.0, 30.0]   # x as a multiple of a
    rows = []
    for a in a_values:
        xs = sorted({r * a for r in rel} | {1e-2, 0.5, 1.0, 2.0, 5.0, 20.0, 100.0, 400.0, 650.0})
        for x in xs:
            q = mpmath.gammainc(mpmath.mpf(a), mpmath.mpf(x), mpmath.inf, regularized=True)
            if q < mpmath.mpf("1e-290"):
                continue
            rows.append({"a": float(a), "x": float(x), "Q": mpmath.nstr(q, 25), "Q_f64": float(q)})
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gamma_q_table.json")
    with open(out, "w") as f:
        json.dump({"generator": "mpmath %s, mp.dps = 40, gammainc(a, x, inf, regularized=True)" % mpmath.__version__,
                   "rows": rows}, f, indent=0)
    print(len(rows), "rows ->", out)


if __name__ == "__main__":
    main()
