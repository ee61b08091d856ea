"""Tabulated Lagrange elements and quadrature rules (host side, numpy).

The reference receives its element kernel as an opaque FFCx-generated
``tabulate_tensor`` whose basis tables are baked into the generated C source
(call site ``cpp/assemble_matrix.cpp:438-439,505-506``).  FFCx/Basix are not in
this image, so the tables are produced here and handed, as plain arrays, to the
device kernels (and to the CPU oracle): quadrature weights, basis values and
reference derivatives at the quadrature points, and the reference derivatives of
the geometry map.

Node ordering follows the Basix/DOLFINx convention: vertices first, then edge
midpoints with edges ordered by (tet) (2,3),(1,3),(1,2),(0,3),(0,2),(0,1) and
(triangle) (1,2),(0,2),(0,1); tensor-product cells number vertices with x
fastest.
"""
from __future__ import annotations

import dataclasses
import functools
from typing import Optional

import numpy as np
from scipy.special import roots_jacobi

CELL_TDIM = {"interval": 1, "triangle": 2, "tetrahedron": 3, "quadrilateral": 2, "hexahedron": 3}
_TRI_EDGES = ((1, 2), (0, 2), (0, 1))
_TET_EDGES = ((2, 3), (1, 3), (1, 2), (0, 3), (0, 2), (0, 1))


def is_simplex(cell_type: str) -> bool:
    return cell_type in ("interval", "triangle", "tetrahedron")


def _gauss_jacobi_01(m: int, alpha: int):
    """m-point Gauss-Jacobi rule for weight (1-x)^alpha on [0, 1]."""
    x, w = roots_jacobi(m, alpha, 0)
    return 0.5 * (x + 1.0), w / 2.0 ** (alpha + 1)


@functools.lru_cache(maxsize=None)
def make_quadrature(cell_type: str, degree: int):
    """Quadrature exact for polynomials of total ``degree`` on the reference cell.

    Simplices use the collapsed (Stroud conical product) Gauss-Jacobi rule,
    tensor cells a Gauss-Legendre product rule.
    """
    m = max(1, (degree + 2) // 2)
    tdim = CELL_TDIM[cell_type]
    if cell_type == "interval":
        x, w = _gauss_jacobi_01(m, 0)
        return x[:, None].copy(), w.copy()
    if cell_type == "triangle":
        if degree <= 1:
            return np.array([[1 / 3, 1 / 3]]), np.array([0.5])
        if degree == 2:  # 3-point rule (what Basix' default scheme gives for degree 2)
            return np.array([[1 / 6, 1 / 6], [1 / 6, 2 / 3], [2 / 3, 1 / 6]]), np.full(3, 1 / 6)
        r, wr = _gauss_jacobi_01(m, 1)
        s, ws = _gauss_jacobi_01(m, 0)
        pts = np.array([[ri, sj * (1 - ri)] for ri in r for sj in s])
        wts = np.array([wi * wj for wi in wr for wj in ws])
        return pts, wts
    if cell_type == "tetrahedron":
        if degree <= 1:
            return np.array([[0.25, 0.25, 0.25]]), np.array([1 / 6])
        if degree == 2:  # 4-point rule
            a, b = 0.5854101966249685, 0.1381966011250105
            return np.array([[b, b, b], [a, b, b], [b, a, b], [b, b, a]]), np.full(4, 1 / 24)
        r, wr = _gauss_jacobi_01(m, 2)
        s, ws = _gauss_jacobi_01(m, 1)
        t, wt = _gauss_jacobi_01(m, 0)
        pts = np.array([[ri, sj * (1 - ri), tk * (1 - ri) * (1 - sj)] for ri in r for sj in s for tk in t])
        wts = np.array([wi * wj * wk for wi in wr for wj in ws for wk in wt])
        return pts, wts
    x, w = _gauss_jacobi_01(m, 0)
    if tdim == 2:
        pts = np.array([[xi, yj] for yj in x for xi in x])
        wts = np.array([wi * wj for wj in w for wi in w])
    else:
        pts = np.array([[xi, yj, zk] for zk in x for yj in x for xi in x])
        wts = np.array([wi * wj * wk for wk in w for wj in w for wi in w])
    return pts, wts


def _simplex_barycentric(pts: np.ndarray):
    tdim = pts.shape[1]
    lam = np.concatenate([1.0 - pts.sum(axis=1, keepdims=True), pts], axis=1)  # (nq, tdim+1)
    dlam = np.zeros((tdim, tdim + 1))
    dlam[:, 0] = -1.0
    for a in range(tdim):
        dlam[a, a + 1] = 1.0
    return lam, dlam


def _lagrange_1d(degree: int, x: np.ndarray):
    """Equispaced 1-D Lagrange basis, Basix ordering (end points first)."""
    if degree == 1:
        return np.stack([1 - x, x], 1), np.stack([-np.ones_like(x), np.ones_like(x)], 1)
    if degree == 2:
        phi = np.stack([(1 - x) * (1 - 2 * x), x * (2 * x - 1), 4 * x * (1 - x)], 1)
        dphi = np.stack([4 * x - 3, 4 * x - 1, 4 - 8 * x], 1)
        return phi, dphi
    raise NotImplementedError(f"degree {degree}")


def tabulate(cell_type: str, degree: int, pts: np.ndarray):
    """Basis values (nq, nd) and reference derivatives (nq, tdim, nd)."""
    pts = np.atleast_2d(np.asarray(pts, dtype=np.float64))
    tdim = CELL_TDIM[cell_type]
    nq = pts.shape[0]
    if is_simplex(cell_type):
        lam, dlam = _simplex_barycentric(pts)
        nv = tdim + 1
        if degree == 1:
            return lam.copy(), np.broadcast_to(dlam[None], (nq, tdim, nv)).copy()
        if degree == 2:
            edges = {1: ((0, 1),), 2: _TRI_EDGES, 3: _TET_EDGES}[tdim]
            nd = nv + len(edges)
            phi = np.zeros((nq, nd))
            dphi = np.zeros((nq, tdim, nd))
            for v in range(nv):
                phi[:, v] = lam[:, v] * (2 * lam[:, v] - 1)
                dphi[:, :, v] = (4 * lam[:, v, None] - 1) * dlam[None, :, v]
            for e, (a, b) in enumerate(edges):
                phi[:, nv + e] = 4 * lam[:, a] * lam[:, b]
                dphi[:, :, nv + e] = 4 * (lam[:, a, None] * dlam[None, :, b] + lam[:, b, None] * dlam[None, :, a])
            return phi, dphi
        raise NotImplementedError(f"Lagrange degree {degree} on {cell_type}")
    if degree != 1:
        raise NotImplementedError(f"Lagrange degree {degree} on {cell_type}")
    p1 = [_lagrange_1d(1, pts[:, a]) for a in range(tdim)]
    nd = 2**tdim
    phi = np.ones((nq, nd))
    dphi = np.ones((nq, tdim, nd))
    for v in range(nd):
        for a in range(tdim):
            bit = (v >> a) & 1
            phi[:, v] *= p1[a][0][:, bit]
            for b in range(tdim):
                dphi[:, b, v] *= p1[a][1][:, bit] if a == b else p1[a][0][:, bit]
    return phi, dphi


def num_element_dofs(cell_type: str, degree: int) -> int:
    return tabulate(cell_type, degree, np.zeros((1, CELL_TDIM[cell_type]))).__getitem__(0).shape[1]


@dataclasses.dataclass(frozen=True)
class ElementTables:
    """Arrays a tabulated kernel consumes; layout matches ``mpcx_tables`` (include/mpcx.h)."""

    cell_type: str
    degree: int
    tdim: int
    gdim: int
    nd: int
    ng: int
    nq: int
    weights: np.ndarray  # (nq,)
    phi: np.ndarray  # (nq, nd)
    dphi: np.ndarray  # (nq, tdim, nd)
    gdphi: np.ndarray  # (nq, tdim, ng)
    # exterior-facet tables: the arrays above hold nfacets consecutive copies, one per local facet (quadrature points
    # of the reference facet mapped into the cell), and ftan the tangents of the reference facet map
    nfacets: int = 0
    ftan: Optional[np.ndarray] = None  # (nfacets, tdim - 1, tdim)
    # rectangular forms (test and trial elements differ, e.g. the div / grad blocks of Taylor-Hood): the trial
    # element's basis at the same quadrature points; nd1 == 0 means "same element on both sides"
    degree1: int = 0
    nd1: int = 0
    phi1: Optional[np.ndarray] = None  # (nq, nd1)
    dphi1: Optional[np.ndarray] = None  # (nq, tdim, nd1)


@functools.lru_cache(maxsize=None)
def element_tables(cell_type: str, degree: int, qdegree: int) -> ElementTables:
    pts, wts = make_quadrature(cell_type, qdegree)
    phi, dphi = tabulate(cell_type, degree, pts)
    _, gdphi = tabulate(cell_type, 1, pts)
    tdim = CELL_TDIM[cell_type]
    return ElementTables(
        cell_type, degree, tdim, tdim, phi.shape[1], gdphi.shape[2], len(wts),
        np.ascontiguousarray(wts), np.ascontiguousarray(phi), np.ascontiguousarray(dphi),
        np.ascontiguousarray(gdphi),
    )


@functools.lru_cache(maxsize=None)
def mixed_element_tables(cell_type: str, degree0: int, degree1: int, qdegree: int) -> ElementTables:
    """Tables of a cell integral whose test element (degree0) and trial element (degree1) differ."""
    t0 = element_tables(cell_type, degree0, qdegree)
    pts, _ = make_quadrature(cell_type, qdegree)
    phi1, dphi1 = tabulate(cell_type, degree1, pts)
    return dataclasses.replace(t0, degree1=degree1, nd1=phi1.shape[1], phi1=np.ascontiguousarray(phi1),
                               dphi1=np.ascontiguousarray(dphi1))


# Local facets as vertex tuples, DOLFINx / Basix numbering: simplex facet i is opposite vertex i; tensor-product
# cells list their facets in lexicographic order of the vertex sets.
FACETS = {
    "triangle": ((1, 2), (0, 2), (0, 1)),
    "tetrahedron": ((1, 2, 3), (0, 2, 3), (0, 1, 3), (0, 1, 2)),
    "quadrilateral": ((0, 1), (0, 2), (1, 3), (2, 3)),
    "hexahedron": ((0, 1, 2, 3), (0, 1, 4, 5), (0, 2, 4, 6), (1, 3, 5, 7), (2, 3, 6, 7), (4, 5, 6, 7)),
}
FACET_TYPE = {"triangle": "interval", "tetrahedron": "triangle", "quadrilateral": "interval", "hexahedron": "quadrilateral"}


def reference_vertices(cell_type: str) -> np.ndarray:
    tdim = CELL_TDIM[cell_type]
    if is_simplex(cell_type):
        return np.concatenate([np.zeros((1, tdim)), np.eye(tdim)])
    return np.array([[(v >> a) & 1 for a in range(tdim)] for v in range(2**tdim)], dtype=np.float64)


@functools.lru_cache(maxsize=None)
def facet_tables(cell_type: str, degree: int, qdegree: int) -> ElementTables:
    """Tables for exterior-facet integrals (the reference passes the local facet index to the generated kernel,
    ``cpp/assemble_matrix.cpp:343-362``): per local facet the basis tabulated at the facet quadrature points.
    The surface measure at a point is ``|J t|`` (2-D) or ``|J t_1 x J t_2|`` (3-D) with ``t`` the tangents of the
    reference facet map, and the weights are those of the reference facet."""
    tdim = CELL_TDIM[cell_type]
    fpts, fw = make_quadrature(FACET_TYPE[cell_type], qdegree)
    verts = reference_vertices(cell_type)
    phis, dphis, gdphis, tans = [], [], [], []
    for fv in FACETS[cell_type]:
        v0 = verts[fv[0]]
        t = np.array([verts[fv[a + 1]] - v0 for a in range(tdim - 1)])  # (tdim-1, tdim)
        pts = v0[None, :] + fpts @ t
        phi, dphi = tabulate(cell_type, degree, pts)
        _, gdphi = tabulate(cell_type, 1, pts)
        phis.append(phi); dphis.append(dphi); gdphis.append(gdphi); tans.append(t)
    phi, dphi, gdphi = np.concatenate(phis), np.concatenate(dphis), np.concatenate(gdphis)
    return ElementTables(
        cell_type, degree, tdim, tdim, phi.shape[1], gdphi.shape[2], len(fw),
        np.ascontiguousarray(fw), np.ascontiguousarray(phi), np.ascontiguousarray(dphi), np.ascontiguousarray(gdphi),
        len(FACETS[cell_type]), np.ascontiguousarray(np.array(tans, dtype=np.float64)),
    )
