"""``assemble_matrix`` / ``create_sparsity_pattern`` / ``create_matrix`` -- the reference's Python surface
(``python/src/dolfinx_mpc/assemble_matrix.py:21-146``) in front of the device kernels."""
from __future__ import annotations

import ctypes as C
import os
from collections.abc import Sequence
from typing import Optional, Union

import numpy as np

from . import _lib
from . import device as _dev
from .fem import DirichletBC, Form
from .la import Matrix
from .multipointconstraint import MultiPointConstraint


def _pair(constraint):
    if isinstance(constraint, Sequence):
        assert len(constraint) == 2
        return tuple(constraint)
    return (constraint, constraint)


def _mpc_host(mpc: MultiPointConstraint):
    keep = (np.ascontiguousarray(mpc.masters.array), np.ascontiguousarray(mpc.masters.offsets),
            np.ascontiguousarray(mpc.cell_to_slaves.array), np.ascontiguousarray(mpc.cell_to_slaves.offsets))
    return _lib.MpcHostS(*[k.ctypes.data for k in keep]), keep


def create_sparsity_pattern(form: Form, mpc: Union[MultiPointConstraint, Sequence[MultiPointConstraint]],
                            num_threads: int = 0):
    """Sparsity pattern with the MPC additions (``assemble_matrix.py:68-88`` -> ``cpp/utils.h:381-496``).

    Returns ``(row_ptr int64, col int32)`` of the scalar CSR over the local (owned + ghost) rows of the
    constraints' function spaces.
    """
    mpc0, mpc1 = _pair(mpc)
    for m in (mpc0, mpc1):
        m._not_finalized()
    if form.rank != 2:
        raise RuntimeError("Cannot create sparsity pattern. Form is not a bilinear form")
    lib = _lib.load()
    V0, V1 = form.function_spaces
    nrows_b = mpc0.function_space.num_blocks
    h0, k0 = _mpc_host(mpc0)
    h1, k1 = _mpc_host(mpc1)
    rp = C.POINTER(C.c_int64)()
    cl = C.POINTER(C.c_int32)()
    nnz = C.c_int64(0)
    nc = form.mesh.num_cells_local
    _lib.check(lib.mpcx_create_pattern_host(V0.dofmap.ctypes.data, V0.nd, V0.bs, V1.dofmap.ctypes.data, V1.nd, V1.bs,
                                            nc, nrows_b, C.byref(h0), C.byref(h1),
                                            num_threads or (os.cpu_count() or 1), C.byref(rp), C.byref(cl),
                                            C.byref(nnz)))
    nrows = nrows_b * V0.bs
    row_ptr = np.ctypeslib.as_array(rp, shape=(nrows + 1,)).copy()
    col = np.ctypeslib.as_array(cl, shape=(max(1, nnz.value),))[: nnz.value].copy()
    lib.mpcx_free_host(rp)
    lib.mpcx_free_host(cl)
    return row_ptr, col


def create_sparsity_pattern_device(form: Form, mpc: Union[MultiPointConstraint, Sequence[MultiPointConstraint]]):
    """The same pattern built on the device (``mpcx_pattern_create`` / ``mpcx_pattern_export``): returns device
    tensors ``(row_ptr int64, col int32)``.  Raises ``MpcxError`` with status ``ERR_UNSUPPORTED`` beyond 2^31 cell
    couplings on one device."""
    import torch

    mpc0, mpc1 = _pair(mpc)
    for m in (mpc0, mpc1):
        m._not_finalized()
    if form.rank != 2:
        raise RuntimeError("Cannot create sparsity pattern. Form is not a bilinear form")
    lib = _lib.load()
    V0, V1 = form.function_spaces
    n0, n1 = mpc0.function_space.num_dofs, mpc1.function_space.num_dofs
    d0 = _dev.dofmap_struct(V0, n0)
    d1 = _dev.dofmap_struct(V1, n1)
    m0 = _dev.mpc_dev(mpc0)["struct"]
    m1 = _dev.mpc_dev(mpc1)["struct"]
    st = _dev.stream_ptr()
    handle = C.c_void_p()
    nnz = C.c_int64(0)
    _lib.check(lib.mpcx_pattern_create(C.byref(d0), C.byref(d1), form.mesh.num_cells_local, C.byref(m0), C.byref(m1), st,
                                       C.byref(handle), C.byref(nnz)))
    try:
        row_ptr = torch.empty(n0 + 1, dtype=torch.int64, device=_dev.device())
        col = torch.empty(max(1, nnz.value), dtype=torch.int32, device=_dev.device())[: nnz.value]
        _lib.check(lib.mpcx_pattern_export(handle, _dev.ptr(row_ptr), _dev.ptr(col), st))
        torch.cuda.current_stream().synchronize()
    finally:
        lib.mpcx_pattern_destroy(handle)
    return row_ptr, col


def create_matrix(form: Form, mpc0: MultiPointConstraint, mpc1: Optional[MultiPointConstraint] = None) -> Matrix:
    """``cpp.mpc.create_matrix`` (``cpp/utils.h:140-173``): pattern + zeroed device CSR.  The pattern is built on the
    device; ``MPCX_PATTERN=host`` (or more than 2^31 cell couplings) selects the threaded host builder."""
    mpc1 = mpc0 if mpc1 is None else mpc1
    shape = (mpc0.function_space.num_dofs, mpc1.function_space.num_dofs)
    bs = (form.function_spaces[0].bs, form.function_spaces[1].bs)
    if os.environ.get("MPCX_PATTERN", "device") != "host":
        try:
            row_ptr, col = create_sparsity_pattern_device(form, (mpc0, mpc1))
            return Matrix(row_ptr, col, shape, bs)
        except _lib.MpcxError as e:
            if getattr(e, "status", None) != _lib.ERR_UNSUPPORTED:
                raise
    row_ptr, col = create_sparsity_pattern(form, (mpc0, mpc1))
    return Matrix(row_ptr, col, shape, bs)


def _bc_markers(V, bcs, n: int):
    """Device int8 marker over the unrolled dofs of ``V`` or None when no bc lives on it
    (``cpp/assemble_matrix.cpp:687-705``).  Built on the host once per set of bcs and kept on the device."""
    mine = [bc for bc in bcs if bc.function_space is V or bc.function_space.dofmap is V.dofmap]
    if not mine:
        return None
    key = ("bc_markers", n) + tuple(bc.uid for bc in mine)  # monotonic uids: an id() could be reused
    if key not in V._dev:
        m = np.zeros(n, dtype=np.int8)
        for bc in mine:
            bc.mark_dofs(m)
        V._dev[key] = _dev.to_dev(m)
    return V._dev[key]


def _bc_owned_dofs(bc):
    """Owned dofs of a bc on the device (``dolfinx.fem.petsc.insert_diagonal`` touches owned rows only)."""
    if "owned_dofs" not in bc._dev:
        V = bc.function_space
        dofs = bc.dofs[bc.dofs < V.index_map.size_local * V.bs]
        bc._dev["owned_dofs"] = (_dev.to_dev(dofs) if len(dofs) else None, len(dofs))
    return bc._dev["owned_dofs"]


def assemble_matrix(form: Form, constraint: Union[MultiPointConstraint, Sequence[MultiPointConstraint]],
                    bcs: Optional[Sequence[DirichletBC]] = None, diagval: float = 1.0, A: Optional[Matrix] = None,
                    num_threads: Optional[int] = 1) -> Matrix:
    """Assemble a bilinear form into a device CSR matrix with multi point constraints and Dirichlet
    conditions; same arguments as the reference (``python/src/dolfinx_mpc/assemble_matrix.py:21-65``).
    ``num_threads`` is accepted and ignored."""
    bcs = [] if bcs is None else list(bcs)
    if not isinstance(constraint, Sequence):
        assert form.function_spaces[0] is form.function_spaces[1]
    mpc0, mpc1 = _pair(constraint)
    lib = _lib.load()
    if A is None:
        A = create_matrix(form, mpc0, mpc1)
    V0, V1 = form.function_spaces
    st = _dev.stream_ptr()
    if V0 is V1:
        # one marker array for rows and columns (a distributed matrix has extra off-process columns that no local
        # cell references): identical markers on both sides let the tile kernels use their symmetric plan
        bc0_d = bc1_d = _bc_markers(V0, bcs, max(A.shape))
    else:
        bc0_d = _bc_markers(V0, bcs, A.shape[0])
        bc1_d = _bc_markers(V1, bcs, A.shape[1])
    mesh_s = _dev.mesh_dev(form.mesh)["struct"]
    d0 = _dev.dofmap_struct(V0, A.shape[0])
    d1 = _dev.dofmap_struct(V1, A.shape[1])
    m0 = _dev.mpc_dev(mpc0)["struct"]
    m1 = _dev.mpc_dev(mpc1)["struct"]
    keep = []
    # Blocked P1 elasticity: the first integral can be assembled by the row-gather kernel, which STORES every row of A
    # (zeros included) instead of adding into a zeroed matrix -- no zero-fill then (csrc/mpcx_rowgather.cuh)
    row_plan = None
    if (A.scatter == "tile" and mpc0 is mpc1 and V0 is V1 and form.integrals and os.environ.get("MPCX_ROWGATHER", "1") != "0"):
        it0 = form.integrals[0]
        if it0.integral_type == "cell":
            s0 = _dev.integral_struct(form, it0, (mpc0, mpc1), keep)
            row_plan = A.row_plan(form, it0, s0, (id(mpc0), id(mpc1)), keepalive=(mpc0, mpc1))
    if row_plan is None:
        A.zeroEntries()
    As = A.struct()  # after zeroEntries: with async_zero the values live in the other buffer now
    for n_it, it in enumerate(form.integrals):
        if it.integral_type not in ("cell", "exterior_facet"):
            raise RuntimeError(f"{it.integral_type} integrals have no device kernel yet")
        s = _dev.integral_struct(form, it, (mpc0, mpc1), keep)
        if n_it == 0 and row_plan is not None:
            sp = (A.slave_plan(form, it, s, bc0_d, bc1_d, mpc0, mpc1)
                  if s.num_slave_cells > 0 and os.environ.get("MPCX_SLAVE_PLAN", "1") != "0" else None)
            _lib.check(lib.mpcx_assemble_matrix_rowgather_f64(C.byref(s), C.byref(mesh_s), C.byref(d0), _dev.ptr(bc0_d),
                                                              C.byref(m0), C.byref(As), row_plan[0], sp, st))
            continue
        facet = it.integral_type == "exterior_facet"  # surface-sized: generic kernel, row search per entry
        tile = (A.tile_plan(form, it, s, bc0_d, bc1_d, (id(mpc0), id(mpc1)), keepalive=(mpc0, mpc1))
                if A.scatter == "tile" and not facet else None)
        if tile is not None:
            try:
                _lib.check(lib.mpcx_assemble_matrix_tiled_f64(C.byref(s), C.byref(mesh_s), C.byref(d0), C.byref(d1),
                                                              _dev.ptr(bc0_d), _dev.ptr(bc1_d), C.byref(m0),
                                                              C.byref(m1), C.byref(As), tile[0], st))
                continue
            except _lib.MpcxError as e:
                # e.g. a coefficient not laid out like the trial element (several packed coefficients): the
                # library refuses before launching anything; the generic device kernel takes the integral
                if getattr(e, "status", None) != _lib.ERR_UNSUPPORTED:
                    raise
        plan = None if facet else A.plan(form, it)
        _lib.check(lib.mpcx_assemble_matrix_f64(C.byref(s), C.byref(mesh_s), C.byref(d0), C.byref(d1),
                                                _dev.ptr(bc0_d), _dev.ptr(bc1_d), C.byref(m0), C.byref(m1),
                                                C.byref(As), None if plan is None else C.byref(plan), st))
    _add_diagonals(A, As, form, mpc0, mpc1, bcs, diagval, st)
    A.check_device_errors(st)
    A.assemble()
    return A


def _add_diagonals(A, As, form, mpc0, mpc1, bcs, diagval, st):
    lib = _lib.load()
    V0, V1 = form.function_spaces
    # slave diagonal for owned slaves when both sides share the constraint space (cpp/assemble_matrix.cpp:711-724).
    # In the reference every MultiPointConstraint owns a freshly created (extended) function space
    # (cpp/MultiPointConstraint.h:117-120), so the shared_ptr comparison holds only for one and the same object.
    if mpc0 is mpc1 and mpc0.num_local_slaves > 0:
        sl = _dev.mpc_dev(mpc0)["slaves"]
        _lib.check(lib.mpcx_add_diagonal_f64(C.byref(As), _dev.ptr(sl), mpc0.num_local_slaves, float(diagval), st))
    # Dirichlet diagonal (assemble_matrix.py:59-62 -> dolfinx insert_diagonal: owned dofs of every bc)
    if V0 is V1:
        for bc in bcs:
            if bc.function_space is V0 or bc.function_space.dofmap is V0.dofmap:
                t, nd_ = _bc_owned_dofs(bc)
                if nd_:
                    _lib.check(lib.mpcx_add_diagonal_f64(C.byref(As), _dev.ptr(t), nd_, float(diagval), st))


def create_matrix_nest(a: Sequence[Sequence[Optional[Form]]], constraints: Sequence[MultiPointConstraint]):
    """Block list of matrices with MPC sparsity (``assemble_matrix.py:91-115``); a nested list stands in for
    the PETSc "nest" Mat."""
    assert len(constraints) == len(a)
    return [[None if a[i][j] is None else create_matrix(a[i][j], constraints[i], constraints[j])
             for j in range(len(a[0]))] for i in range(len(a))]


def assemble_matrix_nest(A, a: Sequence[Sequence[Optional[Form]]], constraints: Sequence[MultiPointConstraint],
                         bcs: Sequence[DirichletBC] = (), diagval: float = 1.0, num_threads: Optional[int] = 1):
    """``assemble_matrix.py:118-146``."""
    for i, a_row in enumerate(a):
        for j, a_block in enumerate(a_row):
            if a_block is not None:
                assemble_matrix(a_block, (constraints[i], constraints[j]), bcs=bcs, diagval=diagval, A=A[i][j],
                                num_threads=num_threads)
