"""Accuracy of the pair kernel's table of f(s) = erfc(alpha sqrt(s))/sqrt(s) (csrc/nbs_internal.h ERFC_TAB_*, built by
buildErfcTable in csrc/nbs_api.cu, evaluated in csrc/k_pair.cu pairStep), restated in NumPy: degree-D interpolation at
Chebyshev nodes per row, c0 in double, c1..cD in single precision, the position inside the row quantised to 23 bits and
centred, the remainder evaluated in single precision.  Prints worst relative error, mean bias and worst per-row bias.
usage: erfc_table_check.py [alpha]"""
import sys

import numpy as np
from scipy.special import erfc

alpha = float(sys.argv[1]) if len(sys.argv) > 1 else 2.628261


def f(s):
    r = np.sqrt(s)
    return erfc(alpha*r)/r


rng = np.random.default_rng(0)
for M, D in ((128, 4), (256, 4), (256, 3), (64, 5)):
    worst, bias = 0.0, []
    for e in range(-7, 1):
        for m in range(0, M, 5):
            lo, w = np.ldexp(1.0 + m/M, e), np.ldexp(1.0/M, e)
            k = np.arange(D + 1)
            dn = 0.5*np.cos(np.pi*(2*k + 1)/(2*(D + 1)))
            co = np.linalg.solve(np.vander(dn, D + 1, increasing=True), f(lo + w/2 + dn*w))
            t = rng.random(200)
            d = ((np.floor(t*2**23)/2**23 - 0.5).astype(np.float32) + np.float32(2.0**-24)).astype(np.float32)
            cf = co[1:].astype(np.float32)
            p = cf[-1]
            for j in range(D - 2, -1, -1):
                p = np.float32(p*d + cf[j])
            err = (co[0] + np.float64(np.float32(p*d)) - f(lo + t*w))/f(lo + t*w)
            worst = max(worst, np.abs(err).max())
            bias.append(err.mean())
    print(f"{M:4d} rows/octave, degree {D}: worst relative error {worst:.2e}, mean bias {np.mean(bias):.2e}, worst row bias {np.abs(bias).max():.2e}")
