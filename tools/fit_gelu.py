"""Fit and check the polynomial behind `gelu_erfc_poly` (csrc/common.cuh): erfc(t / sqrt 2) = 2^-q(t), q of degree 6 on [0, 6],
so that gelu(x) = max(x, 0) - 0.5 |x| 2^-q(min(|x|, 6)).  Weighted least squares on Chebyshev nodes, weights iterated towards
the minimax of the GELU's absolute error; the check runs the fp32 evaluation order of the kernel against scipy's erf."""
import numpy as np
from scipy.special import erf, erfc

T, DEG = 6.0, 6
q = lambda t: -np.log2(erfc(t / np.sqrt(2)))
tt = np.cos(np.pi * (np.arange(4000) + 0.5) / 4000) * T / 2 + T / 2
w = np.ones_like(tt)
for _ in range(30):
    c = np.polynomial.polynomial.polyfit(tt, q(tt), DEG, w=w)
    e = np.polynomial.polynomial.polyval(tt, c) - q(tt)
    ge = np.abs(0.5 * tt * np.exp2(-q(tt)) * np.log(2) * e)
    w = w * (1 + ge / ge.max())
c32 = c.astype(np.float32)
x = np.linspace(-8, 8, 400001)
t = np.minimum(np.abs(x), T).astype(np.float32)
p = np.zeros_like(t)
for k in range(DEG, -1, -1):
    p = (p * t + c32[k]).astype(np.float32)
g = np.maximum(x, 0).astype(np.float32) - np.float32(0.5) * np.abs(x).astype(np.float32) * np.exp2(-p).astype(np.float32)
ref = 0.5 * x * (1 + erf(x / np.sqrt(2)))
err = np.abs(g - ref)
print("coefficients c0..c6:", [float(v) for v in c32])
print(f"max abs error {err.max():.2e} at x = {x[err.argmax()]:.3f}; max error / max(|gelu|, 1e-2) = {(err / np.maximum(np.abs(ref), 1e-2)).max():.2e}")
