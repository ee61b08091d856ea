"""Synthetic, deterministic problem generators (host side, numpy).

Structured simplicial / tensor meshes, Lagrange P1/P2 spaces and multi-point
constraint data in exactly the array layout ``MultiPointConstraint.add_constraint``
takes in the reference (``python/src/dolfinx_mpc/multipointconstraint.py:118-153``):
``slaves`` (local, int32), ``masters`` (global, int64), ``coeffs``, ``owners``,
``offsets``.  They stand in for the reference's geometric constraint builders
(``cpp/PeriodicConstraint.h``, ``cpp/SlipConstraint.h``, ``cpp/ContactConstraint.h``),
which are cold-path, need DOLFINx geometry search and are out of scope
(SURVEY.md section 2, rows 10-12); on matching structured grids their output is
known analytically and is what is produced here.
"""
from __future__ import annotations

import itertools
from typing import Optional, Sequence

import numpy as np

from . import elements as _el
from .fem import FunctionSpace, IndexMap, Mesh


# ------------------------------------------------------------------ meshes

def _lattice(n: Sequence[int], p0, p1):
    n = list(n)
    axes = [np.linspace(a, b, m + 1) for a, b, m in zip(p0, p1, n)]
    while len(axes) < 3:
        axes.append(np.zeros(1))
    Z, Y, X = np.meshgrid(axes[2], axes[1], axes[0], indexing="ij")
    return np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)


def create_rectangle(nx: int, ny: int, cell_type: str = "triangle", p0=(0.0, 0.0), p1=(1.0, 1.0)) -> Mesh:
    """Structured rectangle; triangles use the "right" diagonal (0,0)-(1,1) of every square."""
    x = _lattice((nx, ny), p0, p1)
    sx = nx + 1
    I, J = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v0 = (I + sx * J).ravel().astype(np.int64)
    if cell_type == "triangle":
        cells = np.stack([np.stack([v0, v0 + 1, v0 + 1 + sx], 1), np.stack([v0, v0 + sx, v0 + 1 + sx], 1)], 1)
        cells = cells.reshape(-1, 3)
    elif cell_type == "quadrilateral":
        cells = np.stack([v0, v0 + 1, v0 + sx, v0 + sx + 1], 1)
    else:
        raise ValueError(cell_type)
    return Mesh(x, cells.astype(np.int32), cell_type)


def create_unit_square(nx: int, ny: int, cell_type: str = "triangle") -> Mesh:
    return create_rectangle(nx, ny, cell_type)


def create_box(nx: int, ny: int, nz: int, cell_type: str = "tetrahedron", p0=(0.0, 0.0, 0.0),
               p1=(1.0, 1.0, 1.0)) -> Mesh:
    """Structured box; tetrahedra are the Kuhn split (6 per cube, all sharing the (0,0,0)-(1,1,1) diagonal).

    Cells are numbered cube-major (x fastest, z slowest), the 6 tetrahedra of a cube consecutively; int32
    throughout so that the 10^8-cell benchmark meshes are built without 64-bit temporaries."""
    x = _lattice((nx, ny, nz), p0, p1)
    sx, sy = nx + 1, (nx + 1) * (ny + 1)
    assert (nx + 1) * (ny + 1) * (nz + 1) < 2**31
    v0 = (np.arange(nx, dtype=np.int32)[None, None, :] + np.int32(sx) * np.arange(ny, dtype=np.int32)[None, :, None]
          + np.int32(sy) * np.arange(nz, dtype=np.int32)[:, None, None]).reshape(-1)
    e = (np.int32(1), np.int32(sx), np.int32(sy))
    if cell_type == "tetrahedron":
        cells = np.empty((len(v0), 6, 4), dtype=np.int32)
        for t, (a, b, c) in enumerate(itertools.permutations(range(3))):
            cells[:, t, 0] = v0
            cells[:, t, 1] = v0 + e[a]
            cells[:, t, 2] = v0 + (e[a] + e[b])
            cells[:, t, 3] = v0 + (e[a] + e[b] + e[c])
        cells = cells.reshape(-1, 4)
    elif cell_type == "hexahedron":
        cells = np.stack([v0 + (i * e[0] + j * e[1] + k * e[2]) for k in (0, 1) for j in (0, 1) for i in (0, 1)], 1)
    else:
        raise ValueError(cell_type)
    return Mesh(x, cells.astype(np.int32, copy=False), cell_type)


def create_unit_cube(nx: int, ny: int, nz: int, cell_type: str = "tetrahedron") -> Mesh:
    return create_box(nx, ny, nz, cell_type)


def rotate_mesh(mesh: Mesh, theta: float, axis) -> Mesh:
    """Rigid rotation of the geometry (``python/tests/test_cube_contact.py:28,45`` inclines its boxes this way)."""
    k = np.asarray(axis, dtype=np.float64)
    k = k / np.linalg.norm(k)
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(theta) * Kx + (1 - np.cos(theta)) * (Kx @ Kx)
    return Mesh(mesh.x @ R.T, mesh.x_dofmap, mesh.cell_type)


# ------------------------------------------------------------------ spaces

def _cell_edges(cell_type: str):
    return {"interval": ((0, 1),), "triangle": _el._TRI_EDGES, "tetrahedron": _el._TET_EDGES}[cell_type]


def functionspace(mesh: Mesh, degree: int = 1, bs: int = 1) -> FunctionSpace:
    """Lagrange space of ``degree`` with block size ``bs`` (serial: every block owned)."""
    nn = mesh.x.shape[0]
    if degree == 1:
        dofmap = mesh.x_dofmap.copy()
        coords = mesh.x.copy()
        nblocks = nn
    elif degree == 2 and _el.is_simplex(mesh.cell_type):
        edges = _cell_edges(mesh.cell_type)
        c = mesh.x_dofmap.astype(np.int64)
        keys = []
        for a, b in edges:
            lo = np.minimum(c[:, a], c[:, b])
            hi = np.maximum(c[:, a], c[:, b])
            keys.append(lo * nn + hi)
        keys = np.stack(keys, 1)
        uniq, inv = np.unique(keys.ravel(), return_inverse=True)
        edge_ids = inv.reshape(keys.shape) + nn
        dofmap = np.concatenate([c, edge_ids], axis=1).astype(np.int32)
        mid = 0.5 * (mesh.x[uniq // nn] + mesh.x[uniq % nn])
        coords = np.concatenate([mesh.x, mid], axis=0)
        nblocks = nn + len(uniq)
    else:
        raise NotImplementedError(f"Lagrange degree {degree} on {mesh.cell_type}")
    return FunctionSpace(mesh, degree, dofmap, bs, IndexMap(nblocks), coords)


# ------------------------------------------------------------------ constraints

def _coord_keys(X: np.ndarray, lo: np.ndarray, hi: np.ndarray) -> np.ndarray:
    q = float(1 << 20)
    span = np.where(hi > lo, hi - lo, 1.0)
    k = np.rint((X - lo) / span * q).astype(np.int64)
    return k[:, 0] + (k[:, 1] << 21) + (k[:, 2] << 42)


def empty_constraint():
    return (np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0, np.float64), np.zeros(0, np.int32),
            np.zeros(1, np.int32))


def periodic_constraint(V: FunctionSpace, axes: Sequence[int] = (0,), scale: float = 1.0,
                        exclude_dofs: Optional[np.ndarray] = None, tol: float = 1e-9):
    """Periodic condition u(x) = scale * u(relation(x)) on a box mesh with matching grids.

    Every block with a coordinate at the upper bound along one of ``axes`` is a slave; its master is the block at
    the point with those coordinates moved to the lower bound (combined map as in
    ``python/tests/test_stokes_channelflow.py:48-55``).  On matching grids the reference finds exactly one master
    per slave with coefficient ``scale`` (``cpp/PeriodicConstraint.h:195-200``).  Slaves on Dirichlet dofs
    (``exclude_dofs``) are dropped as ``dolfinx_mpc::is_bc`` does (``cpp/utils.h:1459-1496``).
    """
    X = V.tabulate_dof_coordinates()
    lo, hi = X.min(axis=0), X.max(axis=0)
    at_hi = np.zeros(X.shape[0], dtype=bool)
    Xm = X.copy()
    for a in axes:
        m = np.abs(X[:, a] - hi[a]) < tol * max(1.0, abs(hi[a]))
        at_hi |= m
        Xm[m, a] = lo[a]
    sblocks = np.flatnonzero(at_hi)
    keys = _coord_keys(X, lo, hi)
    order = np.argsort(keys, kind="stable")
    pos = np.searchsorted(keys[order], _coord_keys(Xm[sblocks], lo, hi))
    mblocks = order[pos]
    assert np.allclose(X[mblocks], Xm[sblocks], atol=1e-7), "non-matching periodic grid"
    bs = V.bs
    comp = np.arange(bs)
    slaves = (sblocks[:, None] * bs + comp[None, :]).reshape(-1)
    masters = (mblocks[:, None] * bs + comp[None, :]).reshape(-1)
    if exclude_dofs is not None and len(exclude_dofs):
        keep = ~np.isin(slaves, exclude_dofs)
        slaves, masters = slaves[keep], masters[keep]
    n = len(slaves)
    return (slaves.astype(np.int32), V.index_map.local_to_global(masters // bs) * bs + masters % bs,
            np.full(n, scale, dtype=np.float64), V.index_map.owner_of_local(masters // bs),
            np.arange(n + 1, dtype=np.int32))


def slip_constraint(V: FunctionSpace, blocks: np.ndarray, normals: np.ndarray,
                    exclude_dofs: Optional[np.ndarray] = None):
    """u . n = 0 on ``blocks``: slave = component with the largest |n_i|, masters = the other components of the
    same block with coefficient -n_j / n_slave (``cpp/SlipConstraint.h:123-140``)."""
    bs = V.bs
    blocks = np.asarray(blocks, dtype=np.int64)
    normals = np.broadcast_to(np.asarray(normals, dtype=np.float64), (len(blocks), bs))
    if exclude_dofs is not None and len(exclude_dofs):
        bad = np.isin(blocks, np.asarray(exclude_dofs) // bs)
        blocks, normals = blocks[~bad], normals[~bad]
    s_idx = np.argmax(np.abs(normals), axis=1)
    slaves = blocks * bs + s_idx
    others = np.array([[j for j in range(bs) if j != s] for s in s_idx], dtype=np.int64).reshape(len(blocks), bs - 1)
    masters = (blocks[:, None] * bs + others).reshape(-1)
    n_s = normals[np.arange(len(blocks)), s_idx]
    coeffs = (-np.take_along_axis(normals, others, axis=1) / n_s[:, None]).reshape(-1)
    gm = V.index_map.local_to_global(masters // bs) * bs + masters % bs
    return (slaves.astype(np.int32), gm, coeffs, np.full(len(masters), V.index_map.rank, np.int32),
            (np.arange(len(blocks) + 1) * (bs - 1)).astype(np.int32))


def general_constraint(V: FunctionSpace, slave_master_coeff: dict, comp_slave: int = 0, comp_master: int = 0):
    """Point-dictionary constraint ``{slave_point: {master_point: coeff}}`` -- the serial part of
    ``python/src/dolfinx_mpc/dictcondition.py:31-232`` (used by most reference tests)."""
    X = V.tabulate_dof_coordinates()

    def find(p):
        p = np.asarray(list(p) + [0.0] * (3 - len(p)), dtype=np.float64)
        hits = np.flatnonzero(np.all(np.isclose(X, p[None, :], atol=1e-10), axis=1))
        if len(hits) != 1:
            raise RuntimeError(f"no unique dof at {p}")
        return int(hits[0])

    slaves, masters, coeffs, offsets = [], [], [], [0]
    for sp, mm in slave_master_coeff.items():
        slaves.append(find(sp) * V.bs + comp_slave)
        for mp, cf in mm.items():
            masters.append(find(mp) * V.bs + comp_master)
            coeffs.append(cf)
        offsets.append(len(masters))
    return (np.array(slaves, np.int32), np.array(masters, np.int64), np.array(coeffs, np.float64),
            np.zeros(len(masters), np.int32), np.array(offsets, np.int32))


def tie_constraint(V: FunctionSpace, slave_blocks: np.ndarray, master_block: int, coeff: float, comp: int = 0):
    """Many slaves tied to one master (``python/tests/test_surface_integral.py:83-87``): stress test for
    contention on a single master row."""
    slave_blocks = np.asarray(slave_blocks, dtype=np.int64)
    slave_blocks = slave_blocks[slave_blocks != master_block]
    n = len(slave_blocks)
    return ((slave_blocks * V.bs + comp).astype(np.int32), np.full(n, master_block * V.bs + comp, np.int64),
            np.full(n, coeff), np.zeros(n, np.int32), np.arange(n + 1, dtype=np.int32))


def contact_constraint(V: FunctionSpace, z_interface: float, normal=(0.0, 0.0, 1.0), tol: float = 1e-6,
                       exclude_dofs: Optional[np.ndarray] = None):
    """Contact-slip between two stacked P1 boxes that meet at ``z = z_interface`` with non-matching grids.

    Slaves: blocks of the upper body on the interface; slave component = argmax |n|; masters = the other
    components of the same block (-n_j/n_s) plus, for every component b, the dofs of the lower-body facet
    containing the slave point with coefficient n_b/n_s * phi_j(x); |coeff| <= tol dropped
    (``cpp/ContactConstraint.h:71,87-152``).  The mesh must come from :func:`create_stacked_boxes`.
    """
    info = V.mesh.stack_info
    bs = V.bs
    X = V.tabulate_dof_coordinates()
    n = np.asarray(normal, dtype=np.float64)
    s_c = int(np.argmax(np.abs(n)))
    up_blocks = info["upper_interface_nodes"]
    if exclude_dofs is not None and len(exclude_dofs):
        up_blocks = up_blocks[~np.isin(up_blocks, np.asarray(exclude_dofs) // bs)]
    nxl, nyl = info["lower_n"][:2]
    lower_top = info["lower_top_node0"]
    P = X[up_blocks]
    u = np.clip(P[:, 0] * nxl, 0, nxl - 1e-12)
    v = np.clip(P[:, 1] * nyl, 0, nyl - 1e-12)
    i, j = np.floor(u).astype(np.int64), np.floor(v).astype(np.int64)
    fu, fv = u - i, v - j
    sx = nxl + 1
    v00 = lower_top + i + sx * j
    # Kuhn tets leave the "right" diagonal on every z-face: triangles (v00, v10, v11) and (v00, v01, v11)
    lower = fu >= fv
    nodes = np.where(lower[:, None], np.stack([v00, v00 + 1, v00 + 1 + sx], 1), np.stack([v00, v00 + sx, v00 + 1 + sx], 1))
    phis = np.where(lower[:, None], np.stack([1 - fu, fu - fv, fv], 1), np.stack([1 - fv, fv - fu, fu], 1))
    slaves, masters, coeffs, offsets = [], [], [], [0]
    for k, blk in enumerate(up_blocks):
        slaves.append(blk * bs + s_c)
        for b in range(bs):
            if b != s_c:
                cf = -n[b] / n[s_c]
                if abs(cf) > tol:
                    masters.append(blk * bs + b)
                    coeffs.append(cf)
        for b in range(bs):
            for node, ph in zip(nodes[k], phis[k]):
                cf = n[b] / n[s_c] * ph
                if abs(cf) > tol:
                    masters.append(node * bs + b)
                    coeffs.append(cf)
        offsets.append(len(masters))
    return (np.array(slaves, np.int32), np.array(masters, np.int64), np.array(coeffs, np.float64),
            np.zeros(len(masters), np.int32), np.array(offsets, np.int32))


def create_stacked_boxes(n_lower: Sequence[int], n_upper: Sequence[int]) -> Mesh:
    """Two unit-footprint boxes, [0,1]^2 x [0,0.5] and [0,1]^2 x [0.5,1], meshed independently (non-matching
    interface, the geometry of ``python/tests/test_cube_contact.py:31-45``)."""
    lo = create_box(*n_lower, p0=(0, 0, 0), p1=(1, 1, 0.5))
    up = create_box(*n_upper, p0=(0, 0, 0.5), p1=(1, 1, 1.0))
    nl = lo.x.shape[0]
    mesh = Mesh(np.concatenate([lo.x, up.x]), np.concatenate([lo.x_dofmap, up.x_dofmap + nl]), "tetrahedron")
    nxl, nyl, nzl = n_lower
    nxu, nyu, _ = n_upper
    mesh.stack_info = {
        "lower_n": tuple(n_lower),
        "lower_top_node0": (nxl + 1) * (nyl + 1) * nzl,
        "upper_interface_nodes": nl + np.arange((nxu + 1) * (nyu + 1), dtype=np.int64),
        "num_lower_nodes": nl,
    }
    return mesh
