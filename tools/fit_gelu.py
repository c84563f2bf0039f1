"""Fits the polynomial of the accurate erf-GELU epilogue (csrc/common.cuh: gelu_neg_log2_odds / gelu4).

Phi(x) = 1 / (1 + 2^(-x q(x^2))) holds exactly when x q(x^2) ln 2 = logit(Phi(x)); q is fitted as a degree-DEG
polynomial in x^2 on |x| <= R by (nonlinear) least squares on the error of x Phi(x), then the formula is evaluated
in float32 (exact exp2; the hardware's ex2.approx / rcp.approx add ~2^-22) on [-8, 8] with x^2 clamped at R^2.

    python tools/fit_gelu.py        # prints the coefficients and the maximum absolute error per degree
"""
import numpy as np
from scipy.optimize import least_squares
from scipy.special import log_ndtr, ndtr

R = 5.5


def horner(c, t):
    y = np.zeros_like(t)
    for a in c[::-1]:
        y = y * t + a
    return y


def main():
    x = np.linspace(1e-4, R, 20001)
    phi = ndtr(x)
    target = (log_ndtr(x) - log_ndtr(-x)) / (x * np.log(2))
    for deg in (3, 4, 5):
        V = np.vander(x * x, deg + 1, increasing=True)
        wgt = x * phi * (1 - phi) * np.log(2) * x
        c0 = np.linalg.lstsq(V * wgt[:, None], target * wgt, rcond=None)[0]
        c = least_squares(lambda c: (x / (1 + np.exp2(-x * horner(c, x * x))) - x * phi) * 1e6, c0, xtol=1e-15,
                          ftol=1e-15, gtol=1e-15).x
        xs = np.linspace(-8, 8, 400001).astype(np.float32)
        t = np.minimum(xs * xs, np.float32(R * R)).astype(np.float32)
        q = np.zeros_like(t)
        for a in c[::-1]:
            q = (q * t + np.float32(a)).astype(np.float32)
        e = np.exp2(np.minimum((-xs * q).astype(np.float32), np.float32(28)).astype(np.float64)).astype(np.float32)
        y = (xs / (np.float32(1) + e)).astype(np.float32)
        err = np.abs(y - xs.astype(np.float64) * ndtr(xs.astype(np.float64)))
        print(f"degree {deg}: max |x Phi(x) error| = {err.max():.2e} at x = {xs[err.argmax()]:.2f}; coefficients "
              f"(lowest power first): {[float(np.float32(a)) for a in c]}")


if __name__ == "__main__":
    main()
