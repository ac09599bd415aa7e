"""Zonotope.polygon_vertices (the export behind examples/1.double_integrator_sim.py:170): pure host code, no GPU."""
import itertools

import numpy as np


def _hull_area(pts):
    from scipy.spatial import ConvexHull
    return ConvexHull(pts).volume


def test_polygon_is_the_convex_hull_of_the_vertex_enumeration():
    from tzddpc_b200.zonotope import Zonotope
    rng = np.random.default_rng(0)
    for g in (1, 2, 3, 6, 9):
        c = rng.normal(size=2)
        G = rng.normal(size=(2, g))
        if g >= 3:
            G[:, 1] = 0.0                    # an all-zero generator
            G[:, 2] = -2.0 * G[:, 0]         # parallel generators
        Z = Zonotope(c, G)
        P = Z.polygon_vertices()
        pts = np.array([c + G @ np.array(s) for s in itertools.product((-1.0, 1.0), repeat=g)])
        if g == 1 or np.linalg.matrix_rank(G) < 2:
            assert np.allclose(np.sort(P, axis=0)[[0, -1]], np.sort(pts, axis=0)[[0, -1]])
            continue
        # shoelace area of the ordered boundary = area of the hull of all 2^g candidate vertices, and every candidate is inside
        x, y = P[:, 0], P[:, 1]
        area = 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))
        assert np.isclose(area, _hull_area(pts), rtol=1e-10)
        d = np.roll(P, -1, axis=0) - P
        for q in pts:
            cross = d[:, 0] * (q[1] - P[:, 1]) - d[:, 1] * (q[0] - P[:, 0])
            assert (cross >= -1e-9).all()     # counter-clockwise boundary: every point on the left of every edge
