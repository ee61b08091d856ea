"""Array-backed stand-ins for the DOLFINx objects the reference's API takes.

dolfinx_mpc's assembly entry points receive ``dolfinx.fem.Form``,
``FunctionSpace``, ``DirichletBC`` ... objects (``python/src/dolfinx_mpc/
assemble_matrix.py:21-28``).  DOLFINx is not available in this image, so this
module carries exactly the arrays the hot path reads from those objects
("array-level ABI"; ``numba/assemble_matrix.py:58-104``):
geometry ``x`` / ``x_dofmap``, the blocked ``dofmap``, block size, the
owned/ghost split of the index map, packed coefficients / constants and the
integration domains.  Nothing here computes; everything is numpy on the host.
"""
from __future__ import annotations

import dataclasses
import enum
from typing import Optional, Sequence

import numpy as np

from . import elements as _el


class Kernel(enum.IntEnum):
    """Hand-written element kernels (ids shared with include/mpcx.h).

    They replace the opaque FFCx ``tabulate_tensor`` pointer the reference
    obtains from ``a.kernel(IntegralType::cell, i, 0)`` (``cpp/assemble_matrix.cpp:622``).
    """

    LAPLACE = 0  # c[0] * inner(grad u, grad v) dx   (block-diagonal when bs > 1)
    MASS = 1  # c[0] * inner(u, v) dx
    ELASTICITY = 2  # inner(sigma(u), grad v) dx, c = [mu, lambda], bs == gdim
    SOURCE = 3  # c[0] * inner(f, v) dx, w = f at the cell dofs
    LAPLACE_VARCOEF = 4  # c[0] * w * inner(grad u, grad v) dx, w scalar, same element
    # rectangular blocks (test and trial spaces differ: python/tests/test_rectangular_assembly.py:83-86)
    DIV_TEST = 5  # c[0] * inner(p, div(v)) dx: test = vector space (bs == gdim), trial = scalar space
    DIV_TRIAL = 6  # c[0] * inner(div(u), q) dx: test = scalar space, trial = vector space (bs == gdim)
    CUSTOM = 7  # Integral.custom: tabulate_tensor source compiled at run time (CustomKernel)


@dataclasses.dataclass
class IndexMap:
    """Owned/ghost split of a (blocked) dof numbering: owned ``[0, size_local)``, ghosts after
    (the DOLFINx convention relied on in ``cpp/MultiPointConstraint.h:52-54,112-115``)."""

    size_local: int
    ghosts: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0, np.int64))  # global ids
    owners: np.ndarray = dataclasses.field(default_factory=lambda: np.zeros(0, np.int32))
    local_range: tuple = (0, 0)
    size_global: int = 0
    rank: int = 0

    def __post_init__(self):
        if self.local_range == (0, 0):
            self.local_range = (0, self.size_local)
        if self.size_global == 0:
            self.size_global = self.size_local

    @property
    def num_ghosts(self) -> int:
        return len(self.ghosts)

    def owner_of_local(self, local: np.ndarray) -> np.ndarray:
        """Owning rank of local blocks (this rank for owned ones)."""
        local = np.asarray(local, dtype=np.int64)
        out = np.full(local.shape, self.rank, dtype=np.int32)
        g = local >= self.size_local
        out[g] = self.owners[local[g] - self.size_local]
        return out

    def local_to_global(self, local: np.ndarray) -> np.ndarray:
        local = np.asarray(local, dtype=np.int64)
        out = local + self.local_range[0]
        g = local >= self.size_local
        out[g] = self.ghosts[local[g] - self.size_local]
        return out

    def global_to_local(self, glob: np.ndarray) -> np.ndarray:
        """-1 where the global index is neither owned nor ghosted here."""
        glob = np.asarray(glob, dtype=np.int64)
        out = np.full(glob.shape, -1, dtype=np.int64)
        own = (glob >= self.local_range[0]) & (glob < self.local_range[1])
        out[own] = glob[own] - self.local_range[0]
        if self.num_ghosts:
            order = np.argsort(self.ghosts, kind="stable")
            sg = self.ghosts[order]
            pos = np.searchsorted(sg, glob[~own])
            pos = np.minimum(pos, len(sg) - 1)
            hit = sg[pos] == glob[~own]
            res = np.where(hit, order[pos] + self.size_local, -1)
            out[~own] = res
        return out


@dataclasses.dataclass
class Mesh:
    x: np.ndarray  # (num_nodes, 3) float64, always 3-padded (cpp/assemble_matrix.cpp:473)
    x_dofmap: np.ndarray  # (num_cells, ng) int32
    cell_type: str
    num_cells_local: int = -1  # owned cells (all cells here: ghost_mode none)
    rank: int = 0
    comm_size: int = 1

    def __post_init__(self):
        self.x = np.ascontiguousarray(self.x, dtype=np.float64)
        self.x_dofmap = np.ascontiguousarray(self.x_dofmap, dtype=np.int32)
        if self.num_cells_local < 0:
            self.num_cells_local = self.x_dofmap.shape[0]
        self._dev = {}

    @property
    def tdim(self) -> int:
        return _el.CELL_TDIM[self.cell_type]

    @property
    def num_cells(self) -> int:
        return self.x_dofmap.shape[0]


@dataclasses.dataclass(eq=False)
class FunctionSpace:
    mesh: Mesh
    degree: int
    dofmap: np.ndarray  # (num_cells, nd) int32, blocked indices
    bs: int
    index_map: IndexMap
    dof_coordinates: Optional[np.ndarray] = None  # (num_blocks, 3)

    def __post_init__(self):
        self.dofmap = np.ascontiguousarray(self.dofmap, dtype=np.int32)
        self._dev = {}

    @property
    def nd(self) -> int:
        return self.dofmap.shape[1]

    @property
    def num_blocks(self) -> int:
        return self.index_map.size_local + self.index_map.num_ghosts

    @property
    def num_dofs(self) -> int:
        """Unrolled local dofs including ghosts (rows of the local matrix)."""
        return self.num_blocks * self.bs

    def tabulate_dof_coordinates(self) -> np.ndarray:
        if self.dof_coordinates is None:
            raise RuntimeError("dof coordinates were not provided for this space")
        return self.dof_coordinates

    def with_index_map(self, index_map: IndexMap, dof_coordinates=None) -> "FunctionSpace":
        return FunctionSpace(self.mesh, self.degree, self.dofmap, self.bs, index_map,
                             self.dof_coordinates if dof_coordinates is None else dof_coordinates)


class Function:
    """Nodal coefficient on a space (``array`` has length ``V.num_dofs``)."""

    def __init__(self, V: FunctionSpace, array: Optional[np.ndarray] = None):
        self.function_space = V
        self.array = np.zeros(V.num_dofs) if array is None else np.ascontiguousarray(array, dtype=np.float64)
        assert self.array.shape == (V.num_dofs,)

    def interpolate(self, f):
        """f(x) -> values, x of shape (3, n); for bs > 1 return shape (bs, n)."""
        X = self.function_space.tabulate_dof_coordinates().T
        v = np.asarray(f(X), dtype=np.float64)
        bs = self.function_space.bs
        self.array[:] = v.reshape(-1) if bs == 1 else np.asarray(v).reshape(bs, -1).T.reshape(-1)


class DirichletBC:
    """Dirichlet condition on unrolled local dofs (``mark_dofs`` / ``set`` of dolfinx::fem::DirichletBC,
    as used in ``cpp/assemble_matrix.cpp:691-705`` and ``cpp/lifting.h:176-180``)."""

    _next_uid = 0

    def __init__(self, V: FunctionSpace, dofs: np.ndarray, value=0.0):
        self.function_space = V
        self.dofs = np.ascontiguousarray(dofs, dtype=np.int32)
        self._value = value
        self.version = 0  # bumped when ``value`` is reassigned (a Function / array value can also change in place)
        # monotonic identity: keys device-side caches (id() of a collected object can be handed out again)
        self.uid = DirichletBC._next_uid
        DirichletBC._next_uid += 1
        self._dev = {}

    @property
    def value(self):
        return self._value

    @value.setter
    def value(self, v):
        self._value = v
        self.version += 1

    @property
    def value_is_mutable(self) -> bool:
        """True when the value can change without ``value`` being reassigned (a Function or an array updated in
        place -- the usual DOLFINx pattern for time-dependent conditions); such values are re-read on every
        ``apply_lifting`` / ``set_bc`` call, as the reference does (``cpp/lifting.h:166-180``)."""
        return isinstance(self._value, Function) or np.ndim(self._value) != 0

    def mark_dofs(self, markers: np.ndarray):
        markers[self.dofs] = 1

    def values_at_dofs(self) -> np.ndarray:
        """The prescribed values at ``self.dofs`` (what ``set`` writes)."""
        if isinstance(self.value, Function):
            return self.value.array[self.dofs]
        if np.ndim(self.value) == 0:
            return np.full(len(self.dofs), float(self.value))
        v = np.asarray(self.value, dtype=np.float64)
        bs = self.function_space.bs
        return v[self.dofs % bs] if v.shape == (bs,) else v[self.dofs]

    def set(self, values: np.ndarray):
        values[self.dofs] = self.values_at_dofs()


def locate_dofs_geometrical(V: FunctionSpace, marker) -> np.ndarray:
    """Unrolled dofs (all block components) whose node satisfies ``marker(x)``, x of shape (3, n)."""
    X = V.tabulate_dof_coordinates().T
    blocks = np.flatnonzero(marker(X)).astype(np.int32)
    return (blocks[:, None] * V.bs + np.arange(V.bs, dtype=np.int32)[None, :]).reshape(-1)


def _qdegree(kernel: Kernel, cell_type: str, degree: int) -> int:
    simplex = _el.is_simplex(cell_type)
    if kernel in (Kernel.LAPLACE, Kernel.ELASTICITY):
        return 2 * degree - 2 if simplex else 2 * degree
    if kernel in (Kernel.MASS, Kernel.SOURCE):
        return 2 * degree
    if kernel == Kernel.LAPLACE_VARCOEF:
        return 3 * degree - 2 if simplex else 3 * degree
    if kernel == Kernel.CUSTOM:
        return 1  # the tables only describe the element sizes to the library; the kernel has its own quadrature
    raise ValueError(kernel)


def _qdegree_mixed(cell_type: str, degree0: int, degree1: int) -> int:
    """phi * d(phi'): sum of the degrees minus one on simplices (affine map), sum of the degrees on tensor cells."""
    return degree0 + degree1 - 1 if _el.is_simplex(cell_type) else degree0 + degree1


class CustomKernel:
    """An element kernel outside the registry: C / CUDA source defining
    ``void <entry>(double* A, const double* w, const double* c, const double* coordinate_dofs, const int*
    entity_local_index, const uint8_t* quadrature_permutation)`` -- the UFCx ``tabulate_tensor`` signature, i.e. what
    ``a.kernel(IntegralType::cell, i, 0)`` hands the reference (``cpp/assemble_matrix.cpp:438-439, 620-636``); FFCx
    output can be passed as generated.  Compiled on first use with NVRTC (``mpcx_custom_kernel_create``)."""

    def __init__(self, source: str, entry: str):
        self.source, self.entry = source, entry
        self._handles = {}

    def handle(self, num_entries: int, num_coordinate_dofs: int, num_coefficient_values: int):
        from . import _lib

        key = (num_entries, num_coordinate_dofs, num_coefficient_values)
        if key not in self._handles:
            import ctypes as C

            h = C.c_void_p()
            _lib.check(_lib.load().mpcx_custom_kernel_create(self.source.encode(), self.entry.encode(), num_entries,
                                                            num_coordinate_dofs, num_coefficient_values, C.byref(h)))
            self._handles[key] = h
        return self._handles[key]

    def __del__(self):
        try:
            from . import _lib

            for h in self._handles.values():
                _lib.load().mpcx_custom_kernel_destroy(h)
        except Exception:
            pass


@dataclasses.dataclass
class Integral:
    """One integral of a form: kernel id, integration domain, coefficients, constants.

    ``facets``: (n, 2) array of (cell, local facet) pairs makes it an exterior-facet integral -- the layout of the
    reference's facet lists (``cpp/assemble_matrix.cpp:343-348``); ``cells`` then holds the facets' cells."""

    kernel: Kernel
    constants: np.ndarray
    coefficients: Sequence[Function] = ()
    cells: Optional[np.ndarray] = None  # active cells (int32); None = every owned cell in order
    integral_type: str = "cell"
    facets: Optional[np.ndarray] = None
    local_facets: Optional[np.ndarray] = None
    custom: Optional[CustomKernel] = None  # kernel == Kernel.CUSTOM

    def __post_init__(self):
        self.constants = np.ascontiguousarray(self.constants, dtype=np.float64)
        if self.facets is not None:
            self.facets = np.ascontiguousarray(self.facets, dtype=np.int32).reshape(-1, 2)
            self.cells = self.facets[:, 0]
            self.local_facets = np.ascontiguousarray(self.facets[:, 1])
            self.integral_type = "exterior_facet"
        if self.cells is not None:
            self.cells = np.ascontiguousarray(self.cells, dtype=np.int32)
        self._dev = {}


class Form:
    """A compiled form: rank, argument spaces and its integrals.

    ``rank == 2`` bilinear (``function_spaces = (V0, V1)`` test/trial), ``rank == 1`` linear.
    """

    def __init__(self, rank: int, function_spaces: Sequence[FunctionSpace], integrals: Sequence[Integral]):
        self.rank = rank
        self.function_spaces = tuple(function_spaces)
        self.integrals = list(integrals)
        self.mesh = self.function_spaces[0].mesh
        for it in self.integrals:
            if it.integral_type == "interior_facet":
                # cpp/assemble_matrix.cpp:658-659, cpp/assemble_vector.cpp:242-245
                raise RuntimeError("Not implemented yet")

    def __add__(self, other: "Form") -> "Form":
        assert self.rank == other.rank and self.function_spaces == other.function_spaces
        return Form(self.rank, self.function_spaces, self.integrals + other.integrals)

    def tables(self, integral: Integral) -> _el.ElementTables:
        V = self.function_spaces[0]
        if integral.kernel in (Kernel.DIV_TEST, Kernel.DIV_TRIAL):
            if integral.integral_type != "cell":
                raise RuntimeError("the div coupling kernels are cell integrals")
            V1 = self.function_spaces[1]
            return _el.mixed_element_tables(self.mesh.cell_type, V.degree, V1.degree,
                                            _qdegree_mixed(self.mesh.cell_type, V.degree, V1.degree))
        if (integral.kernel == Kernel.CUSTOM and self.rank == 2 and integral.integral_type == "cell"
                and self.function_spaces[1].degree != V.degree):
            # a custom kernel between different elements: the tables only carry the two elements' sizes
            V1 = self.function_spaces[1]
            return _el.mixed_element_tables(self.mesh.cell_type, V.degree, V1.degree, 1)
        make = _el.facet_tables if integral.integral_type == "exterior_facet" else _el.element_tables
        return make(self.mesh.cell_type, V.degree, _qdegree(integral.kernel, self.mesh.cell_type, V.degree))

    def active_cells(self, integral: Integral) -> np.ndarray:
        if integral.cells is None:
            return np.arange(self.mesh.num_cells_local, dtype=np.int32)
        return integral.cells

    def pack_coefficients(self, integral: Integral):
        """Host mirror of ``dolfinx::fem::pack_coefficients`` (called at ``cpp/assemble_matrix.cpp:587-589``):
        one row of ``cstride`` scalars per active cell, in iteration order."""
        if not integral.coefficients:
            return None, 0
        cells = self.active_cells(integral)
        rows = []
        for f in integral.coefficients:
            Vf = f.function_space
            d = Vf.dofmap[cells].astype(np.int64)
            idx = (d[:, :, None] * Vf.bs + np.arange(Vf.bs)[None, None, :]).reshape(len(cells), -1)
            rows.append(f.array[idx])
        w = np.ascontiguousarray(np.concatenate(rows, axis=1))
        return w, w.shape[1]


# -- form constructors (what ``dolfinx.fem.form(ufl_expression)`` is to the reference) ---------------------------

def laplace(V: FunctionSpace, kappa: float = 1.0, cells=None) -> Form:
    """``kappa * inner(grad(u), grad(v)) * dx`` (e.g. ``python/tests/test_matrix_assembly.py:36-38``)."""
    return Form(2, (V, V), [Integral(Kernel.LAPLACE, [kappa], cells=cells)])


def mass(V: FunctionSpace, rho: float = 1.0, cells=None, facets=None) -> Form:
    """``rho * inner(u, v) * dx``, or ``* ds`` over the given exterior facets (a Robin term)."""
    return Form(2, (V, V), [Integral(Kernel.MASS, [rho], cells=cells, facets=facets)])


def elasticity(V: FunctionSpace, mu: float, lmbda: float, cells=None) -> Form:
    """``inner(sigma(u), grad(v)) * dx`` as in ``python/benchmarks/bench_elasticity_edge.py:125-135``."""
    assert V.bs == V.mesh.tdim
    return Form(2, (V, V), [Integral(Kernel.ELASTICITY, [mu, lmbda], cells=cells)])


def div_test(V: FunctionSpace, Q: FunctionSpace, scale: float = -1.0, cells=None) -> Form:
    """``scale * inner(p, div(v)) * dx`` with ``v`` in the vector space ``V`` (test) and ``p`` in the scalar space
    ``Q`` (trial): the ``a01`` block of ``python/tests/test_rectangular_assembly.py:83-86`` for ``scale = -1``."""
    assert V.bs == V.mesh.tdim and Q.bs == 1 and V.mesh is Q.mesh
    return Form(2, (V, Q), [Integral(Kernel.DIV_TEST, [scale], cells=cells)])


def div_trial(Q: FunctionSpace, V: FunctionSpace, scale: float = -1.0, cells=None) -> Form:
    """``scale * inner(div(u), q) * dx`` with ``q`` in the scalar space ``Q`` (test) and ``u`` in the vector space
    ``V`` (trial): the ``a10`` block of the same test."""
    assert V.bs == V.mesh.tdim and Q.bs == 1 and V.mesh is Q.mesh
    return Form(2, (Q, V), [Integral(Kernel.DIV_TRIAL, [scale], cells=cells)])


def laplace_varcoef(V: FunctionSpace, w: Function, scale: float = 1.0, cells=None) -> Form:
    assert w.function_space.bs == 1 and w.function_space.nd == V.nd
    return Form(2, (V, V), [Integral(Kernel.LAPLACE_VARCOEF, [scale], (w,), cells=cells)])


def source(V: FunctionSpace, f: Function, scale: float = 1.0, cells=None, facets=None) -> Form:
    """``scale * inner(f, v) * dx`` with ``f`` interpolated into ``V`` (``python/benchmarks/bench_periodic.py:85-91``),
    or ``* ds`` over the given exterior facets (the traction term of ``python/tests/test_surface_integral.py:52-72``)."""
    assert f.function_space.bs == V.bs and f.function_space.nd == V.nd
    return Form(1, (V,), [Integral(Kernel.SOURCE, [scale], (f,), cells=cells, facets=facets)])


def custom_form(spaces: Sequence[FunctionSpace], kernel: CustomKernel, constants=(), coefficients=(), cells=None,
                facets=None) -> Form:
    """A form whose element tensor comes from a :class:`CustomKernel` -- ``dolfinx.fem.form(any UFL expression)`` in
    the reference: ``len(spaces)`` is the rank, ``coefficients`` are packed per cell in the order given (each
    ``[nd][bs]``, as ``pack_coefficients`` does), ``constants`` fill ``c``."""
    spaces = tuple(spaces)
    return Form(len(spaces), spaces, [Integral(Kernel.CUSTOM, list(constants), tuple(coefficients), cells=cells, facets=facets,
                                               custom=kernel)])


def locate_exterior_facets(mesh: Mesh, marker=None) -> np.ndarray:
    """(cell, local facet) pairs of the boundary facets (facets of exactly one cell) whose vertices all satisfy
    ``marker(x)``, x of shape (3, n) -- ``locate_entities_boundary`` + the facet-to-cell lookup DOLFINx does when it
    builds a ``ds`` integration domain."""
    fv = np.array(_el.FACETS[mesh.cell_type])  # (nf, nvf)
    nc = mesh.num_cells_local
    nodes = mesh.x_dofmap[:nc][:, fv]  # (nc, nf, nvf) geometry vertices of every local facet (P1 geometry)
    key = np.sort(nodes.reshape(nc * fv.shape[0], -1), axis=1)
    _, inv, cnt = np.unique(key, axis=0, return_inverse=True, return_counts=True)
    boundary = cnt[inv.reshape(-1)] == 1
    if marker is not None:
        on = marker(mesh.x.T)
        boundary &= on[nodes.reshape(nc * fv.shape[0], -1)].all(axis=1)
    idx = np.flatnonzero(boundary)
    return np.stack([idx // fv.shape[0], idx % fv.shape[0]], axis=1).astype(np.int32)
