"""Coefficients of the FMA-pipe tanh used for part of the layer-1 epilogue of csrc/rl_policy.cu.

tanh(x) ~ clamp(x R(min(x^2, c^2)), -1, 1) with R a polynomial of degree d in u = x^2 (Horner in
packed FFMA2).  Minimises the maximum ABSOLUTE error over all x (the result is rounded to bf16, whose
half-ulp is 2e-3 in [0.5, 1)); beyond c the linear continuation reaches 1 and is clamped there.
Prints the fp32 coefficients (highest degree first) and the measured error of the fp32 evaluation.

    python tools/tanh_poly_fit.py [d] [c]
"""
import sys
import numpy as np
from numpy.polynomial import chebyshev as C


def fit(c, d, iters=200):
    x = np.linspace(1e-4, c, 40001)
    s = 2 * x * x / (c * c) - 1
    A = np.stack([x * C.chebval(s, np.eye(d + 1)[k]) for k in range(d + 1)], 1)
    y = np.tanh(x)
    w = np.ones_like(x)
    best = None
    for _ in range(iters):
        coef, *_ = np.linalg.lstsq(A * w[:, None], y * w, rcond=None)
        e = np.abs(A @ coef - y)
        m = e.max()
        if best is None or m < best[0]:
            best = (m, coef.copy())
        w = w * (1 + 3 * e / m) ** 0.5
        w /= w.mean()
    # Chebyshev series in s -> monomials in u
    p_s = C.cheb2poly(best[1])                      # ascending powers of s
    # s = 2u/c^2 - 1
    p_u = np.zeros(d + 1)
    base = np.array([-1.0, 2.0 / (c * c)])
    cur = np.array([1.0])
    for k in range(d + 1):
        p_u[:len(cur)] += p_s[k] * cur
        cur = np.convolve(cur, base)
    return best[0], p_u


def evaluate(p_u, c):
    x = np.linspace(-12, 12, 2_400_001).astype(np.float32)
    u = np.minimum(x * x, np.float32(c * c))
    r = np.full_like(x, np.float32(p_u[-1]))
    for k in range(len(p_u) - 2, -1, -1):
        r = (r * u + np.float32(p_u[k])).astype(np.float32)
    t = np.clip(x * r, -1, 1)
    return np.abs(t.astype(np.float64) - np.tanh(x.astype(np.float64))).max()


if __name__ == "__main__":
    d = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    cs = [float(sys.argv[2])] if len(sys.argv) > 2 else [3.2, 3.4, 3.5, 3.6, 3.8, 4.0]
    for c in cs:
        m, p_u = fit(c, d)
        print(f"d={d} c={c}: fit max abs err {m:.2e}; fp32 evaluation incl. clamp over [-12, 12]: {evaluate(p_u, c):.2e}")
        print("   coefficients, highest degree first:", ", ".join(f"{v:.9e}f" for v in p_u[::-1]))
