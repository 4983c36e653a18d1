"""Gauss-Legendre quadrature collapsed onto the reference tetrahedron.

API mirror of the reference's src/diffelastic/gauss.py:4-37
(`calculate_legendre_roots_weights`, `generate_gauss_points_weights`): same
rule, same fp32 rounding points, so that the constant tables derived from it
(element mass table, stiffness contraction table) agree bit for bit with what
the reference computes on the CPU.
"""
import numpy as np
from numpy.polynomial import legendre as _leg


def calculate_legendre_roots_weights(order):
    """Roots of P_order on [-1, 1] and their Gauss weights (float64 arrays)."""
    coef = np.zeros(order + 1, dtype=np.float32)
    coef[order] = 1
    roots = _leg.legroots(coef)
    slope = _leg.Legendre(coef).deriv()(roots)
    return roots, 2 / ((1 - roots ** 2) * slope ** 2)


def generate_gauss_points_weights(order):
    """order**3 points in barycentric form (x, y, z, w) and weights, float32.

    The 1-D rule is mapped to [0, 1] and collapsed (Duffy) so that the weights
    sum to the volume 1/6 of the unit tet.  Index = i*order^2 + j*order + k.
    """
    r, g = calculate_legendre_roots_weights(order)
    r = (r + 1) / 2
    f32 = np.float32
    pts = np.empty((order ** 3, 4), dtype=f32)
    wts = np.empty(order ** 3, dtype=f32)
    q = 0
    for i in range(order):
        w = f32(r[i])
        for j in range(order):
            z = f32(r[j] * (1 - w))
            for k in range(order):
                y = f32(r[k] * (1 - w - z))
                x = f32(1 - w - z - y)
                pts[q] = (x, y, z, w)
                wts[q] = g[i] * g[j] * g[k] * (1 - w) * (1 - w - z) / 8
                q += 1
    return pts, wts
