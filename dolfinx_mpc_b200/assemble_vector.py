"""``assemble_vector`` / ``apply_lifting`` -- the reference's Python surface
(``python/src/dolfinx_mpc/assemble_vector.py:25-147``) in front of the device kernels."""
from __future__ import annotations

import ctypes as C
from collections.abc import Sequence
from typing import Optional

import numpy as np
import torch

from . import _lib
from . import device as _dev
from .fem import DirichletBC, Form
from .fem import Function as fem_Function
from .la import Vector
from .multipointconstraint import MultiPointConstraint


def create_vector(constraint: MultiPointConstraint) -> Vector:
    return Vector(constraint.function_space.num_dofs)


def assemble_vector(form: Form, constraint: MultiPointConstraint, b: Optional[Vector] = None,
                    num_threads: Optional[int] = 1) -> Vector:
    """Assemble a linear form into ``b`` with the multi point constraint applied (``K^T b``);
    ``b`` is zeroed first, as the reference's wrapper does (``assemble_vector.py:79-104``)."""
    lib = _lib.load()
    constraint._not_finalized()
    if b is None:
        b = create_vector(constraint)
    b.set(0.0)
    V = form.function_spaces[0]
    st = _dev.stream_ptr()
    mesh_s = _dev.mesh_dev(form.mesh)["struct"]
    dm = _dev.dofmap_struct(V, constraint.function_space.num_dofs)
    m = _dev.mpc_dev(constraint)["struct"]
    keep = []
    for it in form.integrals:
        if it.integral_type not in ("cell", "exterior_facet"):
            raise RuntimeError(f"{it.integral_type} integrals have no device kernel yet")
        s = _dev.integral_struct(form, it, (constraint,), keep)
        plan = None if it.integral_type != "cell" else _vector_tile_plan(form, it, s, constraint, mesh_s, dm)
        if plan is not None and b.tile_ok:
            try:
                _lib.check(lib.mpcx_assemble_vector_tiled_f64(C.byref(s), C.byref(mesh_s), C.byref(dm), C.byref(m),
                                                              _dev.ptr(b.data), plan[0], st))
                continue
            except _lib.MpcxError as e:  # coefficient layout without a tile kernel: generic device kernel
                if getattr(e, "status", None) != _lib.ERR_UNSUPPORTED:
                    raise
        _lib.check(lib.mpcx_assemble_vector_f64(C.byref(s), C.byref(mesh_s), C.byref(dm), C.byref(m),
                                                _dev.ptr(b.data), st))
    return b


class _PlanHandle:
    """Owns a tile plan of libmpcx (released with the object that caches it)."""

    def __init__(self, handle):
        self.handle = handle

    def __del__(self):
        try:
            _lib.load().mpcx_tile_plan_destroy(self.handle)
        except Exception:
            pass


def _vector_tile_plan(form: Form, it, s_integral, constraint, mesh_s, dm):
    """Vector tile plan (csrc/mpcx_tile.cuh) of one integral, built on the device on first use and cached on
    the integral; None when the element has no vector tile kernel (or MPCX_SCATTER=atomic)."""
    import os

    if os.environ.get("MPCX_SCATTER", "tile") != "tile":
        return None
    V = form.function_spaces[0]
    tab = form.tables(it)
    p1 = V.nd == tab.tdim + 1 and tab.ng == tab.tdim + 1 and V.bs == 1
    if not p1 or int(it.kernel) != 3:
        return None
    if len(it.coefficients) != 1 or s_integral.coeff_nd != V.nd or s_integral.coeff_bs != 1:
        return None  # the C side's w_ok condition: the coefficient must be laid out like the test element
    key = ("vector_tile_plan", id(V), id(constraint))
    if key not in it._dev:
        lib = _lib.load()
        ncells = int(s_integral.num_cells)
        skip = None
        if s_integral.num_slave_cells > 0:
            skip = torch.zeros(ncells, dtype=torch.int8, device=_dev.device())
            skip[it._dev[("slave_cells", id(constraint))][0].long()] = 1
        handle = C.c_void_p()
        try:
            _lib.check(lib.mpcx_vector_tile_plan_create(C.byref(mesh_s), C.byref(dm), s_integral.cells, ncells,
                                                        _dev.ptr(skip), _dev.stream_ptr(), C.byref(handle)))
        except _lib.MpcxError as e:  # tiles that do not fit the plan format: atomic-scatter kernel instead
            if getattr(e, "status", None) != _lib.ERR_UNSUPPORTED:
                raise
            it._dev[key] = None
            return None
        _lib.check(lib.mpcx_device_error(_dev.stream_ptr()))
        info = (C.c_int64 * 15)()
        lib.mpcx_tile_plan_info(handle, info, 15)
        it._dev[("keepalive",) + key] = (V, constraint)  # their id() keys the plan: must not be reused while cached
        it._dev[key] = (handle, _PlanHandle(handle),
                        dict(zip(("tiles", "cells_per_tile", "bulk_cells", "max_nodes", "max_dests", "tile_nodes",
                                  "dests", "bytes", "max_slots", "slots", "max_runs", "runs", "max_stage", "symmetric", "interface_tiles"), [int(v) for v in info])))
    return it._dev[key]


def apply_lifting(b: Vector, form: Sequence[Form], bcs: Sequence[Sequence[DirichletBC]],
                  constraint: MultiPointConstraint, x0: Optional[Sequence] = None, scale: float = 1.0,
                  num_threads: Optional[int] = 1):
    """``b <- b - scale * K^T A_j (g_j - x0_j)`` (``assemble_vector.py:25-76`` -> ``cpp/lifting.h:441-670``)."""
    lib = _lib.load()
    constraint._not_finalized()
    x0 = [] if x0 is None else list(x0)
    # argument checks of cpp/lifting.h:452-465
    if len(x0) and len(x0) != len(form):
        raise RuntimeError("Mismatch in size between x0 and bilinear form in assembler.")
    if len(form) != len(bcs):
        raise RuntimeError("Mismatch in size between a and bcs in assembler.")
    st = _dev.stream_ptr()
    keep = []
    for j, a in enumerate(form):
        if a is None:
            continue
        V0, V1 = a.function_spaces
        n1 = V1.num_dofs
        if not bcs[j]:
            continue
        # bc_markers1 / bc_values1 of cpp/lifting.h:166-180.  Markers, the flagged cell lists and the bc dof indices
        # are structure: cached per tuple of bcs (keyed by their monotonic uid, never by id()).  The VALUES are
        # re-read on every call when they can change in place (Function- or array-valued bcs), or when ``value``
        # was reassigned: only the bc-dof values travel (a persistent device buffer is updated in place).
        mkey = ("lift", n1) + tuple(bc.uid for bc in bcs[j])
        if mkey not in V1._dev:
            markers = np.zeros(n1, dtype=np.int8)
            for bc in bcs[j]:
                bc.mark_dofs(markers)
            if markers.any():
                V1._dev[mkey] = {"markers": _dev.to_dev(markers),
                                 "values": torch.zeros(n1, dtype=torch.float64, device=_dev.device()),
                                 "dofs": [_dev.to_dev(bc.dofs.astype(np.int64)) for bc in bcs[j]],
                                 "versions": None}
            else:
                V1._dev[mkey] = None
        entry = V1._dev[mkey]
        if entry is None:
            continue
        versions = tuple(bc.version for bc in bcs[j])
        if entry["versions"] != versions or any(bc.value_is_mutable for bc in bcs[j]):
            for bc, dofs_d in zip(bcs[j], entry["dofs"]):  # in order: a later bc overrides an earlier one
                if np.ndim(bc.value) == 0 and not isinstance(bc.value, fem_Function):
                    entry["values"].index_fill_(0, dofs_d, float(bc.value))
                else:
                    entry["values"].index_copy_(0, dofs_d, _dev.to_dev(bc.values_at_dofs()))
            entry["versions"] = versions
        mk, vl = entry["markers"], entry["values"]
        x0_d = None
        if len(x0):
            x0_d = x0[j].data if isinstance(x0[j], Vector) else _dev.to_dev(np.asarray(x0[j], dtype=np.float64))
        keep += [mk, vl, x0_d]
        mesh_s = _dev.mesh_dev(a.mesh)["struct"]
        d0 = _dev.dofmap_struct(V0, constraint.function_space.num_dofs)
        d1 = _dev.dofmap_struct(V1, n1)
        m0 = _dev.mpc_dev(constraint)["struct"]
        for it in a.integrals:
            if it.integral_type not in ("cell", "exterior_facet"):
                raise RuntimeError(f"{it.integral_type} integrals have no device kernel yet")
            s = _dev.integral_struct(a, it, (constraint,), keep)
            # cells of this integral with a Dirichlet column: found once on the device, then reused
            ckey = ("lift_cells",) + mkey[1:]
            if ckey not in it._dev:
                ncells = int(s.num_cells)
                flags = torch.zeros(ncells, dtype=torch.int8, device=_dev.device())
                _lib.check(lib.mpcx_flag_cells(C.byref(d1), s.cells, ncells, _dev.ptr(mk), _dev.ptr(flags), st))
                it._dev[ckey] = torch.nonzero(flags).to(torch.int32).reshape(-1)
            lst = it._dev[ckey]
            if lst.numel() == 0:
                continue
            _lib.check(lib.mpcx_apply_lifting_f64(C.byref(s), C.byref(mesh_s), C.byref(d0), C.byref(d1),
                                                  _dev.ptr(mk), _dev.ptr(vl), _dev.ptr(x0_d), float(scale),
                                                  C.byref(m0), _dev.ptr(lst), lst.numel(), _dev.ptr(b.data), st))


def set_bc(b: Vector, bcs: Sequence[DirichletBC], x0=None, scale: float = 1.0):
    """``dolfinx.fem.petsc.set_bc``: b[dofs] = scale * (g - x0) on owned bc dofs (host-side index write)."""
    import torch

    for bc in bcs:
        vals = np.zeros(b.data.numel())
        bc.set(vals)
        n_owned = bc.function_space.index_map.size_local * bc.function_space.bs
        dofs = bc.dofs[bc.dofs < n_owned]
        g = vals[dofs] - (0.0 if x0 is None else np.asarray(x0)[dofs])
        b.data[torch.from_numpy(dofs.astype(np.int64)).to(b.data.device)] = torch.from_numpy(scale * g).to(b.data.device)


def create_vector_nest(L: Sequence[Form], constraints: Sequence[MultiPointConstraint]):
    assert len(constraints) == len(L)
    return [create_vector(c) for c in constraints]


def assemble_vector_nest(b, L: Sequence[Form], constraints: Sequence[MultiPointConstraint],
                         num_threads: Optional[int] = 1):
    """``assemble_vector.py:130-147``."""
    assert len(constraints) == len(L)
    for i, L_row in enumerate(L):
        assemble_vector(L_row, constraints[i], b=b[i], num_threads=num_threads)
