"""Known-answer tests for the tabulated elements and quadrature (inputs shared by oracle and device kernels)."""
import math

import numpy as np
import pytest

from dolfinx_mpc_b200 import elements as el


@pytest.mark.parametrize("cell,vol", [("interval", 1.0), ("triangle", 0.5), ("tetrahedron", 1 / 6),
                                      ("quadrilateral", 1.0), ("hexahedron", 1.0)])
@pytest.mark.parametrize("degree", [0, 1, 2, 3, 4, 6])
def test_quadrature_integrates_monomials(cell, vol, degree):
    pts, wts = el.make_quadrature(cell, degree)
    assert abs(wts.sum() - vol) < 1e-14
    tdim = el.CELL_TDIM[cell]
    rng = np.random.default_rng(0)
    for _ in range(5):
        if el.is_simplex(cell):
            e = rng.multinomial(degree, np.ones(tdim + 1) / (tdim + 1))[:tdim]
            exact = math.prod(math.factorial(int(k)) for k in e) / math.factorial(int(e.sum()) + tdim)
        else:
            e = rng.integers(0, degree + 1, size=tdim)
            exact = math.prod(1.0 / (k + 1) for k in e)
        num = (wts * np.prod(pts ** e[None, :], axis=1)).sum()
        assert abs(num - exact) < 1e-14


@pytest.mark.parametrize("cell,degree,nd", [("triangle", 1, 3), ("triangle", 2, 6), ("tetrahedron", 1, 4),
                                            ("tetrahedron", 2, 10), ("quadrilateral", 1, 4), ("hexahedron", 1, 8)])
def test_lagrange_basis(cell, degree, nd):
    tdim = el.CELL_TDIM[cell]
    pts = np.random.default_rng(1).random((7, tdim)) / tdim
    phi, dphi = el.tabulate(cell, degree, pts)
    assert phi.shape == (7, nd) and dphi.shape == (7, tdim, nd)
    assert np.allclose(phi.sum(axis=1), 1.0) and np.allclose(dphi.sum(axis=2), 0.0)
    h = 1e-6
    for a in range(tdim):  # derivative tables against central differences
        dp = pts.copy(); dp[:, a] += h
        dm = pts.copy(); dm[:, a] -= h
        fd = (el.tabulate(cell, degree, dp)[0] - el.tabulate(cell, degree, dm)[0]) / (2 * h)
        assert np.allclose(fd, dphi[:, a, :], atol=1e-8)


def test_nodal_property_p2_tet():
    """Kronecker property at the DOLFINx node ordering: vertices, then edges (2,3),(1,3),(1,2),(0,3),(0,2),(0,1)."""
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=float)
    nodes = [v[i] for i in range(4)] + [(v[a] + v[b]) / 2 for a, b in el._TET_EDGES]
    phi, _ = el.tabulate("tetrahedron", 2, np.array(nodes))
    assert np.allclose(phi, np.eye(10))
