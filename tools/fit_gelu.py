"""Fit used by dahitra_b200/csrc/decoder_tc.cu::gelu_phi8.

Phi(-u) = 2^g(u) with g a degree-8 polynomial on [0, 6]: iteratively re-weighted least squares (Lawson) towards the
minimax of the error of GELU(h) = h * Phi(h), i.e. weight |h| * Phi(-|h|) on the error of g.  Prints the float32
coefficients (constant term first) and the measured max abs error of the float32 evaluation against the fp64 GELU,
next to the error of the usual fp32 formula 0.5 h (1 + erf(h / sqrt 2)).
"""
import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P
from scipy.special import erf, erfc, ndtr

U, N = 6.0, 8
u = np.linspace(0, U, 40001)
g = np.log2(0.5 * erfc(u / np.sqrt(2)))
w0 = np.maximum(u, 0.25) * 0.5 * erfc(u / np.sqrt(2))
w, x = w0.copy(), 2 * u / U - 1
for _ in range(200):
    c = C.chebfit(x, g, N, w=w)
    e = np.abs(C.chebval(x, c) - g) * w0
    w = w * (1 + 4 * e / e.max())
    w /= w.max()
p, lin, comp, powx = C.cheb2poly(c), np.array([-1, 2 / U]), np.array([0.0]), np.array([1.0])
for k in range(N + 1):
    comp = P.polyadd(comp, p[k] * powx)
    powx = P.polymul(powx, lin)
print("coefficients (c0..c8):", ", ".join("%.9ef" % np.float32(v) for v in comp))

hs = np.linspace(-8, 8, 800001).astype(np.float32)
uu = np.minimum(np.abs(hs), np.float32(U))
acc = np.full_like(uu, np.float32(comp[-1]))
for k in range(N - 1, -1, -1):
    acc = (acc * uu + np.float32(comp[k])).astype(np.float32)
q = np.exp2(acc.astype(np.float64)).astype(np.float32)
gelu = (hs * np.where(hs >= 0, np.float32(1) - q, q).astype(np.float32)).astype(np.float64)
ref = hs.astype(np.float64) * ndtr(hs.astype(np.float64))
f32 = (np.float32(0.5) * hs * (np.float32(1) + erf((hs * np.float32(0.70710678)).astype(np.float32)).astype(np.float32))).astype(np.float64)
for r in (0.5, 1, 2, 3, 8):
    m = np.abs(hs) <= r
    print("|h| <= %-3s  exp2-poly %.2e   fp32 erf formula %.2e" % (r, np.abs(gelu - ref)[m].max(), np.abs(f32 - ref)[m].max()))
