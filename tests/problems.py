"""Shared problem factory for the CPU (oracle) and GPU (parity) tests.

Each case mirrors a fixture of the reference's test-suite (cited per case) on the structured meshes of
``dolfinx_mpc_b200.generators``; sizes are small enough for the oracle to finish in well under a second.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Optional

import numpy as np

from dolfinx_mpc_b200 import fem, generators as gen


@dataclasses.dataclass
class Case:
    name: str
    V: fem.FunctionSpace
    a: fem.Form
    L: Optional[fem.Form]
    data: tuple  # add_constraint arrays
    bcs: list
    a_lift: Optional[fem.Form] = None


def _source(V, expr):
    f = fem.Function(V)
    f.interpolate(expr)
    return fem.source(V, f)


def _f2d(x):
    return x[1] * np.sin(2 * np.pi * x[0]) + 0.3


def _f3d(x):
    return x[0] * np.sin(5 * np.pi * x[1]) + np.exp(-((x[0] - 0.5) ** 2 + (x[1] - 0.5) ** 2 + (x[2] - 0.5) ** 2) / 0.02)


def _vec(f, bs):
    return lambda x: np.stack([(k + 1.0) * f(x) + 0.1 * k for k in range(bs)])


def case_general_2d(cell="triangle", degree=1, nx=5, ny=3, master=(1, 1)):
    """python/tests/test_matrix_assembly.py:23-57 (5x3) and :61-102 (1x8): two slaves, one with two masters."""
    mesh = gen.create_unit_square(nx, ny, cell)
    V = gen.functionspace(mesh, degree)
    data = gen.general_constraint(V, {(1, 0): {(0, 1): 0.43, (1, 1): 0.11}, (0, 0): {master: 0.69}})
    return Case(f"general2d-{cell}-P{degree}-{nx}x{ny}-m{master}", V, fem.laplace(V), _source(V, _f2d), data, [])


def case_periodic_2d(n=8, degree=1, with_bc=True):
    """BASELINE config 1 in small: periodic in x, Dirichlet on y in {0, 1} (python/benchmarks/bench_periodic.py:63-91)."""
    mesh = gen.create_unit_square(n, n)
    V = gen.functionspace(mesh, degree)
    bcs = []
    exclude = None
    if with_bc:
        dofs = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[1], 0) | np.isclose(x[1], 1))
        bcs = [fem.DirichletBC(V, dofs, 0.7)]
        exclude = dofs
    data = gen.periodic_constraint(V, axes=(0,), exclude_dofs=exclude)
    a = fem.laplace(V)
    return Case(f"periodic2d-P{degree}-{n}-bc{int(with_bc)}", V, a, _source(V, _f2d), data, bcs, a_lift=a)


def case_periodic_3d(n=4, degree=1, axes=(0, 1), with_bc=True, bs=1):
    """BASELINE configs 2/4 in small: periodic x/y (or all faces), Dirichlet on z in {0, 1}
    (combined map of python/tests/test_stokes_channelflow.py:48-55)."""
    mesh = gen.create_unit_cube(n, n, n)
    V = gen.functionspace(mesh, degree, bs)
    bcs = []
    exclude = None
    if with_bc:
        dofs = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[2], 0) | np.isclose(x[2], 1))
        bcs = [fem.DirichletBC(V, dofs, 0.25)]
        exclude = dofs
    data = gen.periodic_constraint(V, axes=axes, exclude_dofs=exclude)
    a = fem.laplace(V, 1.3) + fem.mass(V, 0.7) if bs == 1 else fem.laplace(V, 1.3)
    L = _source(V, _f3d if bs == 1 else _vec(_f3d, bs))
    return Case(f"periodic3d-P{degree}-bs{bs}-{n}-ax{len(axes)}-bc{int(with_bc)}", V, a, L, data, bcs, a_lift=a)


def case_slip_elasticity_3d(n=3, degree=2, theta=np.pi / 5):
    """BASELINE config 3 in small: P2 (or P1) vector elasticity on a rotated cube, slip u.n = 0 on one face
    (cpp/SlipConstraint.h:123-140), Dirichlet on the opposite face (python/benchmarks/bench_elasticity_edge.py)."""
    mesh0 = gen.create_unit_cube(n, n, n)
    V0 = gen.functionspace(mesh0, degree, 3)
    X0 = V0.tabulate_dof_coordinates()
    slip_blocks = np.flatnonzero(np.isclose(X0[:, 0], 1.0))
    fixed = np.flatnonzero(np.isclose(X0[:, 0], 0.0))
    mesh = gen.rotate_mesh(mesh0, theta, (1, 1, 0))
    V = gen.functionspace(mesh, degree, 3)
    k = np.array([1.0, 1.0, 0.0]) / np.sqrt(2)
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(theta) * Kx + (1 - np.cos(theta)) * (Kx @ Kx)
    normal = R @ np.array([1.0, 0.0, 0.0])
    bc_dofs = (fixed[:, None] * 3 + np.arange(3)[None, :]).reshape(-1).astype(np.int32)
    bcs = [fem.DirichletBC(V, bc_dofs, np.array([0.01, -0.02, 0.03]))]
    data = gen.slip_constraint(V, slip_blocks, normal, exclude_dofs=bc_dofs)
    E, nu = 1.0e4, 0.1
    mu, lmbda = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
    a = fem.elasticity(V, mu, lmbda)
    L = _source(V, _vec(_f3d, 3))
    return Case(f"slip3d-P{degree}-{n}", V, a, L, data, bcs, a_lift=a)


def case_contact_3d(nl=(3, 3, 2), nu=(4, 4, 2)):
    """BASELINE config 5 in small: two stacked boxes with non-matching grids, contact-slip on the interface with
    many masters per slave (cpp/ContactConstraint.h:87-152; python/tests/test_cube_contact.py:163-302)."""
    mesh = gen.create_stacked_boxes(nl, nu)
    V = gen.functionspace(mesh, 1, 3)
    X = V.tabulate_dof_coordinates()
    bottom = np.flatnonzero(np.isclose(X[:, 2], 0.0))
    bc_dofs = (bottom[:, None] * 3 + np.arange(3)[None, :]).reshape(-1).astype(np.int32)
    bcs = [fem.DirichletBC(V, bc_dofs, 0.0)]
    n = np.array([0.2, -0.1, 1.0])
    data = gen.contact_constraint(V, 0.5, normal=n / np.linalg.norm(n))
    mu, lmbda = 1.0e3 / 2, 0.0
    a = fem.elasticity(V, mu, lmbda)
    return Case("contact3d", V, a, _source(V, _vec(_f3d, 3)), data, bcs, a_lift=a)


def case_tie_2d(n=6, bs=2):
    """python/tests/test_surface_integral.py:83-87: N-1 slaves on one edge all tied to a single master."""
    mesh = gen.create_unit_square(n, n)
    V = gen.functionspace(mesh, 1, bs)
    X = V.tabulate_dof_coordinates()
    edge = np.flatnonzero(np.isclose(X[:, 1], 1.0))
    master = edge[0]
    data = gen.tie_constraint(V, edge, master, 0.8, comp=bs - 1)
    left = np.flatnonzero(np.isclose(X[:, 0], 0.0) & ~np.isclose(X[:, 1], 1.0))
    bc_dofs = (left * bs).astype(np.int32)
    bcs = [fem.DirichletBC(V, bc_dofs, 1.5)]
    a = fem.laplace(V) + fem.mass(V, 2.0)
    return Case(f"tie2d-bs{bs}", V, a, _source(V, _vec(_f2d, bs) if bs > 1 else _f2d), data, bcs, a_lift=a)


def case_lifting_single_quad():
    """python/tests/test_lifting.py:24-76: one Q1 cell, u = 2.3 on x = 1, slave (0,0) -> master (0,1), alpha = 1."""
    mesh = gen.create_unit_square(1, 1, "quadrilateral")
    V = gen.functionspace(mesh, 1)
    dofs = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[0], 1))
    bcs = [fem.DirichletBC(V, dofs, 2.3)]
    data = gen.general_constraint(V, {(0, 0): {(0, 1): 1.0}})
    a = fem.laplace(V)
    return Case("lifting-quad", V, a, _source(V, _f2d), data, bcs, a_lift=a)


def case_varcoef_subdomains(n=6):
    """python/tests/test_integration_domains.py:23-133 (several cell integrals on sub-domains) and
    python/tests/test_mpc_pipeline.py:24-112 (a coefficient Function inside the bilinear form)."""
    mesh = gen.create_unit_square(n, n)
    V = gen.functionspace(mesh, 1)
    cells = np.arange(mesh.num_cells_local, dtype=np.int32)
    xc = mesh.x[mesh.x_dofmap].mean(axis=1)
    left, right = cells[xc[:, 0] < 0.5], cells[xc[:, 0] >= 0.5]
    w = fem.Function(V)
    w.interpolate(lambda x: 1.0 + x[0] + 2 * x[1])
    a = fem.laplace(V, 1.0, cells=left) + fem.laplace_varcoef(V, w, 2.0, cells=right) + fem.mass(V, 0.5)
    data = gen.periodic_constraint(V, axes=(1,), scale=0.5)
    f = fem.Function(V)
    f.interpolate(_f2d)
    L = fem.source(V, f, 1.0, cells=left) + fem.source(V, f, 3.0, cells=right)
    return Case("varcoef-subdomains", V, a, L, data, [], a_lift=None)


def case_empty(n=3):
    """No constraint at all: the MPC path must reduce to plain assembly."""
    mesh = gen.create_unit_cube(n, n, n)
    V = gen.functionspace(mesh, 1)
    return Case("empty-mpc", V, fem.laplace(V), _source(V, _f3d), gen.empty_constraint(), [])


def case_hex(n=3):
    mesh = gen.create_unit_cube(n, n, n, "hexahedron")
    V = gen.functionspace(mesh, 1)
    data = gen.periodic_constraint(V, axes=(0,))
    return Case("hex-periodic", V, fem.laplace(V) + fem.mass(V), _source(V, _f3d), data, [])


def case_surface_traction_2d(n=6):
    """python/tests/test_surface_integral.py:27-120: vector P1 elasticity on the unit square, Dirichlet on the left
    wall, traction ``inner(g, v) * ds`` on the top facets, N-1 slaves on x = 1 tied to the master at (1, 1) with
    coefficient 0.8 (component 1) -- plus a Robin term ``2 inner(u, v) * ds`` on the bottom facets so that the
    bilinear form has an exterior-facet integral as well (cpp/assemble_matrix.cpp:271-415)."""
    mesh = gen.create_unit_square(n, n)
    V = gen.functionspace(mesh, 1, 2)
    X = V.tabulate_dof_coordinates()
    left = np.flatnonzero(np.isclose(X[:, 0], 0.0))
    bc_dofs = (left[:, None] * 2 + np.arange(2)[None, :]).reshape(-1).astype(np.int32)
    bcs = [fem.DirichletBC(V, bc_dofs, 0.2)]
    right = np.flatnonzero(np.isclose(X[:, 0], 1.0) & ~np.isclose(X[:, 1], 1.0) & ~np.isclose(X[:, 1], 0.0))
    master = int(np.flatnonzero(np.isclose(X[:, 0], 1.0) & np.isclose(X[:, 1], 1.0))[0])
    data = gen.tie_constraint(V, right, master, 0.8, comp=1)
    top = fem.locate_exterior_facets(mesh, lambda x: np.isclose(x[1], 1.0))
    bottom = fem.locate_exterior_facets(mesh, lambda x: np.isclose(x[1], 0.0))
    g = fem.Function(V)
    g.interpolate(lambda x: np.stack([0.3 * x[0], -9.81 + 0.0 * x[0]]))
    a = fem.elasticity(V, 50.0, 0.0) + fem.mass(V, 2.0, facets=bottom)
    L = _source(V, _vec(_f2d, 2)) + fem.source(V, g, 1.0, facets=top)
    return Case("surface-traction2d", V, a, L, data, bcs, a_lift=a)


def case_surface_robin_3d(n=3, degree=1):
    """Scalar Robin / Neumann terms over all boundary facets of a tetrahedral mesh with a periodic constraint
    (exterior-facet integrals of matrix, vector and lifting: cpp/assemble_matrix.cpp:271-415,
    cpp/assemble_vector.cpp:196-240, cpp/lifting.h:316-397)."""
    mesh = gen.create_unit_cube(n, n, n)
    V = gen.functionspace(mesh, degree)
    dofs = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[2], 0.0))
    bcs = [fem.DirichletBC(V, dofs, -0.4)]
    data = gen.periodic_constraint(V, axes=(0,), exclude_dofs=dofs)
    fac = fem.locate_exterior_facets(mesh, lambda x: np.isclose(x[2], 1.0) | np.isclose(x[1], 0.0) | np.isclose(x[1], 1.0))
    allf = fem.locate_exterior_facets(mesh)
    h = fem.Function(V)
    h.interpolate(lambda x: 1.0 + x[0] * x[1] - x[2])
    a = fem.laplace(V) + fem.mass(V, 1.5, facets=allf)
    L = _source(V, _f3d) + fem.source(V, h, 0.7, facets=fac)
    return Case(f"surface-robin3d-P{degree}", V, a, L, data, bcs, a_lift=a)


@dataclasses.dataclass
class StokesCase:
    """Taylor-Hood blocks of python/tests/test_rectangular_assembly.py:25-199: P2 vector velocity x P1 pressure on a
    rotated square, slip constraint on the (rotated) face x = 1, Dirichlet velocity on the face x = 0."""

    V: fem.FunctionSpace
    Q: fem.FunctionSpace
    a: list  # [[a00, a01], [a10, None]]
    L: list  # [L0, L1]
    data_v: tuple
    bcs: list


def case_stokes_2d(n=4, theta=np.pi / 4, cell="triangle"):
    mesh0 = gen.create_unit_square(n, n, cell)
    V0 = gen.functionspace(mesh0, 2, 2)
    X0 = V0.tabulate_dof_coordinates()
    inlet = np.flatnonzero(np.isclose(X0[:, 0], 0.0))
    slip_blocks = np.flatnonzero(np.isclose(X0[:, 0], 1.0))
    mesh = gen.rotate_mesh(mesh0, theta, (0, 0, 1))
    V = gen.functionspace(mesh, 2, 2)
    Q = gen.functionspace(mesh, 1, 1)
    bc_dofs = (inlet[:, None] * 2 + np.arange(2)[None, :]).reshape(-1).astype(np.int32)
    bcs = [fem.DirichletBC(V, bc_dofs, np.array([0.3, -0.1]))]
    normal = np.array([np.cos(theta), np.sin(theta)])
    data_v = gen.slip_constraint(V, slip_blocks, normal, exclude_dofs=bc_dofs)
    a = [[fem.laplace(V), fem.div_test(V, Q, -1.0)], [fem.div_trial(Q, V, -1.0), None]]
    zero = fem.Function(Q)
    L = [_source(V, _vec(_f2d, 2)), fem.source(Q, zero)]
    return StokesCase(V, Q, a, L, data_v, bcs)


def case_vector_poisson_cross_component(n=5):
    """python/tests/test_vector_poisson.py:25-133: block size 2, the slave in sub-space 0 and its master in sub-space 1
    (two such pairs, one with two masters), Dirichlet condition on one wall in both components + lifting."""
    mesh = gen.create_unit_square(n, n)
    V = gen.functionspace(mesh, 1, 2)
    X = V.tabulate_dof_coordinates()
    wall = np.flatnonzero(np.isclose(X[:, 0], 0.0))
    bc_dofs = (wall[:, None] * 2 + np.arange(2)[None, :]).reshape(-1).astype(np.int32)
    bcs = [fem.DirichletBC(V, bc_dofs, 0.1)]
    h = 1.0 / n
    data = gen.general_constraint(V, {(1.0, 0.0): {(1.0, 1.0): 0.1, (1.0 - h, 1.0): 0.3}, (1.0, 2 * h): {(1.0, 1.0 - h): 0.7}},
                                  comp_slave=0, comp_master=1)
    a = fem.laplace(V) + fem.mass(V, 0.5)
    return Case("vector-poisson-cross-component", V, a, _source(V, _vec(_f2d, 2)), data, bcs, a_lift=a)


def case_hex_elasticity(n=3):
    """python/tests/test_cube_contact.py:163-302 in small: Q1 hexahedra, vector elasticity (bs 3), the top plane of the
    cube tied component-wise to the bottom plane (periodic in z), Dirichlet on the wall x = 0."""
    mesh = gen.create_unit_cube(n, n, n, "hexahedron")
    V = gen.functionspace(mesh, 1, 3)
    X = V.tabulate_dof_coordinates()
    wall = np.flatnonzero(np.isclose(X[:, 0], 0.0))
    bc_dofs = (wall[:, None] * 3 + np.arange(3)[None, :]).reshape(-1).astype(np.int32)
    bcs = [fem.DirichletBC(V, bc_dofs, 0.0)]
    data = gen.periodic_constraint(V, axes=(2,), exclude_dofs=bc_dofs)
    a = fem.elasticity(V, 3.0, 1.5)
    return Case("hex-elasticity-periodic", V, a, _source(V, _vec(_f3d, 3)), data, bcs, a_lift=a)


ALL_CASES: dict = {}
for _c in (
    lambda: case_general_2d("triangle", 1), lambda: case_general_2d("triangle", 2),
    lambda: case_general_2d("quadrilateral", 1), lambda: case_general_2d("triangle", 1, 1, 8, (0, 1)),
    lambda: case_general_2d("triangle", 2, 1, 8, (1, 1)),
    lambda: case_periodic_2d(8, 1, True), lambda: case_periodic_2d(6, 2, False),
    lambda: case_periodic_3d(4, 1, (0, 1), True), lambda: case_periodic_3d(3, 1, (0, 1, 2), False),
    lambda: case_periodic_3d(3, 2, (0, 1), True), lambda: case_periodic_3d(3, 1, (0,), True, bs=3),
    lambda: case_slip_elasticity_3d(2, 2), lambda: case_slip_elasticity_3d(3, 1),
    case_contact_3d, lambda: case_tie_2d(6, 2), lambda: case_tie_2d(5, 1), case_lifting_single_quad,
    case_varcoef_subdomains, case_empty, case_hex, case_surface_traction_2d,
    lambda: case_surface_robin_3d(3, 1), lambda: case_surface_robin_3d(2, 2),
    case_vector_poisson_cross_component, case_hex_elasticity,
):
    _k = _c()
    ALL_CASES[_k.name] = _c
