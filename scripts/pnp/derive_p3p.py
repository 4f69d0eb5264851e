"""Derivation of the P3P quartic used by csrc/pnp.cu (Grunert's formulation): with s2 = u s1, s3 = v s1 the three
law-of-cosines equations reduce to one quartic in v; sympy expands it and prints the coefficients as C expressions."""
import sympy as sp

v, u = sp.symbols("v u")
ca, cb, cg = sp.symbols("ca cb cg")        # cos(alpha) = f2.f3, cos(beta) = f1.f3, cos(gamma) = f1.f2
q1, q2 = sp.symbols("q1 q2")               # q1 = (a^2 - c^2) / b^2, q2 = c^2 / b^2   (a = |X2-X3|, b = |X1-X3|, c = |X1-X2|)
# (A) - (B):  u * 2 (cg - v ca) = q1 (1 + v^2 - 2 v cb) - v^2 + 1
num = q1 * (1 + v**2 - 2 * v * cb) - v**2 + 1
den = 2 * (cg - v * ca)
# (B): 1 + u^2 - 2 u cg = q2 (1 + v^2 - 2 v cb)      with u = num / den, times den^2
poly = sp.expand(den**2 + num**2 - 2 * num * den * cg - q2 * (1 + v**2 - 2 * v * cb) * den**2)
P = sp.Poly(poly, v)
assert P.degree() == 4
for k, c in enumerate(reversed(P.all_coeffs())):
    print(f"A{k} =", sp.ccode(sp.factor(c)))
