"""Derivation of the coefficients of gelu_fast (csrc/common.cuh): gelu(x) = relu(x) - |x| h(|x|), h(a) = 0.5 erfc(a / sqrt 2) = 2^P(a),
P a polynomial fit of log2 h on [0, 5.5] (weighted least squares on Chebyshev nodes, re-weighted towards the minimax solution).
CPU only (numpy / scipy): python profiles/fit_gelu_poly.py"""
import numpy as np
from scipy.special import erfc

A = 5.5
a = np.cos(np.pi * (np.arange(4001) + 0.5) / 4001) * A / 2 + A / 2
f = np.log2(0.5 * erfc(a / np.sqrt(2)))
xs = np.linspace(0, A, 200001)
fx = np.log2(0.5 * erfc(xs / np.sqrt(2)))
for deg in (4, 5, 6, 7):
    w = np.ones_like(a)
    for _ in range(30):
        coef = np.polyfit(a, f, deg, w=w)
        e = np.abs(np.polyval(coef, a) - f)
        w = w * (1 + 2 * e / e.max())
        w /= w.mean()
    h, hh = 0.5 * erfc(xs / np.sqrt(2)), 2.0 ** np.polyval(coef, xs)
    print(f"degree {deg}: max |dlog2| {np.abs(np.polyval(coef, xs) - fx).max():.2e}, relative error of h {(np.abs(hh - h) / h).max():.2e}, "
          f"max abs error of a*h {np.abs(xs * (hh - h)).max():.2e}")
    print("   coefficients (high -> low):", ", ".join(f"{c:.9e}f" for c in coef))
# end-to-end check of the degree-6 form in float32 against the exact GELU
x = np.linspace(-12, 12, 2000001).astype(np.float32)
c6 = [2.615383824e-05, -6.609828710e-04, 7.488321837e-03, -5.197044650e-02, -4.603294121e-01, -1.150584037e+00, -1.000036059e+00]
aa = np.minimum(np.abs(x), np.float32(5.5)).astype(np.float32)
pl = np.float32(c6[0])
for c in c6[1:]:
    pl = (pl * aa + np.float32(c)).astype(np.float32)
y = (np.maximum(x, 0) - np.abs(x) * np.exp2(pl)).astype(np.float32)
from scipy.special import erf  # noqa: E402
exact = 0.5 * x.astype(np.float64) * (1 + erf(x.astype(np.float64) / np.sqrt(2)))
print(f"float32 evaluation on [-12, 12]: max abs error vs exact GELU {np.abs(y - exact).max():.2e}")
