"""The assembly block of the reference's ``LinearProblem.solve`` (``python/src/dolfinx_mpc/problem.py:539-572``):
zero + ``assemble_matrix`` + ``A.assemble()``, zero + ``assemble_vector``, ``apply_lifting``, ghost update -- as one
call, so that matrix and load vector of the bulk cells come out of ONE pass over the cells
(``csrc/mpcx_tile_fused.cuh``) when the forms allow it.  The result is the same as calling the three routines."""
from __future__ import annotations

import ctypes as C
import os
from collections.abc import Sequence
from typing import Optional

from . import _lib
from . import device as _dev
from .assemble_matrix import _add_diagonals, _bc_markers, assemble_matrix, create_matrix
from .assemble_vector import _vector_tile_plan, apply_lifting, assemble_vector, create_vector
from .fem import DirichletBC, Form
from .la import Matrix, Vector
from .multipointconstraint import MultiPointConstraint


def _fusable(a: Form, L: Form, A: Matrix, b: Vector) -> bool:
    if len(a.integrals) != 1 or len(L.integrals) != 1 or A.scatter != "tile" or not b.tile_ok:
        return False
    ia, iL = a.integrals[0], L.integrals[0]
    V0, V1 = a.function_spaces
    return (V0 is V1 and L.function_spaces[0] is V0 and ia.integral_type == "cell" and iL.integral_type == "cell"
            and ia.cells is None and iL.cells is None and int(ia.kernel) in (0, 1, 4) and int(iL.kernel) == 3)


def assemble_system(a: Form, L: Form, constraint: MultiPointConstraint, bcs: Optional[Sequence[DirichletBC]] = None,
                    diagval: float = 1.0, A: Optional[Matrix] = None, b: Optional[Vector] = None,
                    x0: Optional[Sequence] = None, scale: float = 1.0):
    """``A <- K^T a K`` (bcs applied, slave / Dirichlet diagonals set), ``b <- K^T (L - scale * a (g - x0))`` with
    ghost rows / entries reduced: exactly ``assemble_matrix(a, constraint, bcs, diagval, A)``,
    ``assemble_vector(L, constraint, b)``, ``apply_lifting(b, [a], [bcs], constraint, x0, scale)``,
    ``b.ghostUpdate()``.  Returns ``(A, b)``; ``A.last_system_fused`` tells which path ran."""
    bcs = [] if bcs is None else list(bcs)
    constraint._not_finalized()
    if A is None:
        A = create_matrix(a, constraint)
    if b is None:
        b = create_vector(constraint)
    A.last_system_fused = False
    plans = _fused_plans(a, L, constraint, bcs, A, b) if _fusable(a, L, A, b) else None
    if plans is None:
        assemble_matrix(a, constraint, bcs=bcs, diagval=diagval, A=A)
        assemble_vector(L, constraint, b=b)
    else:
        lib = _lib.load()
        st = _dev.stream_ptr()
        sa, sL, mesh_s, dm, bc_d, m, mplan, vplan, keep = plans[:9]
        A.zeroEntries()
        As = A.struct()  # after zeroEntries: with async_zero the values live in the other buffer now
        b.set(0.0)
        # With several ranks and MPCX_OVERLAP=1: the cells holding slaves and the tiles that touch ghost rows first
        # (part 1); the ghost rows then travel to their owners on a second stream WHILE the interior tiles are assembled
        # (part 2).  Off by default -- measured on two B200s (profiles/README.md, r02_h): the exchange costs 0.26 ms of a
        # 4.8 ms step, and the NCCL send/recv kernel, which holds whole SMs while it waits for its neighbour, delays the
        # persistent CTAs of part 2 by more than that (6.5 ms; 5.0 ms with dynamically claimed tiles, which in turn cost
        # the tile kernel 6 %).
        n_if, n_t = plans[-1]
        overlap = (A.ghost_exchange is not None and getattr(A.ghost_exchange, "mat", None) is not None and 0 < n_if < n_t
                   and os.environ.get("MPCX_OVERLAP", "0") == "1")

        def part(k):
            _lib.check(lib.mpcx_assemble_system_tiled_part_f64(C.byref(sa), C.byref(sL), C.byref(mesh_s), C.byref(dm),
                                                               _dev.ptr(bc_d), C.byref(m), C.byref(As), _dev.ptr(b.data),
                                                               mplan, vplan, k, st))

        try:
            part(1 if overlap else 0)
            A.last_system_fused = True
        except _lib.MpcxError as e:
            if getattr(e, "status", None) != _lib.ERR_UNSUPPORTED:
                raise
        if A.last_system_fused and overlap:
            import torch

            cur = torch.cuda.current_stream(A.val.device)
            if getattr(A, "_exchange_stream", None) is None:
                A._exchange_stream = torch.cuda.Stream(A.val.device)
            ready, done = torch.cuda.Event(), torch.cuda.Event()
            ready.record(cur)
            A._exchange_stream.wait_event(ready)
            with torch.cuda.stream(A._exchange_stream):
                A.assemble()  # ghost rows -> owners (mpcx_ghost_reduce_f64 on the exchange stream)
                done.record(A._exchange_stream)
            part(2)
            _add_diagonals(A, As, a, constraint, constraint, bcs, diagval, st)
            A.check_device_errors(st)
            cur.wait_event(done)
        elif A.last_system_fused:
            _add_diagonals(A, As, a, constraint, constraint, bcs, diagval, st)
            A.check_device_errors(st)
            A.assemble()
        else:  # refused before anything was launched (e.g. a coefficient layout without a tile kernel)
            assemble_matrix(a, constraint, bcs=bcs, diagval=diagval, A=A)
            assemble_vector(L, constraint, b=b)
    if bcs:
        apply_lifting(b, [a], [bcs], constraint, x0=x0, scale=scale)
    b.ghostUpdate()
    return A, b


def _fused_plans(a, L, constraint, bcs, A, b):
    V = a.function_spaces[0]
    ia, iL = a.integrals[0], L.integrals[0]
    keep = []
    bc_d = _bc_markers(V, bcs, max(A.shape))
    sa = _dev.integral_struct(a, ia, (constraint, constraint), keep)
    sL = _dev.integral_struct(L, iL, (constraint,), keep)
    mesh_s = _dev.mesh_dev(a.mesh)["struct"]
    dm = _dev.dofmap_struct(V, A.shape[0])
    mplan = A.tile_plan(a, ia, sa, bc_d, bc_d, (id(constraint), id(constraint)), keepalive=(constraint, constraint))
    if mplan is None:
        return None
    dmv = _dev.dofmap_struct(V, constraint.function_space.num_dofs)
    vplan = _vector_tile_plan(L, iL, sL, constraint, mesh_s, dmv)
    if vplan is None:
        return None
    return (sa, sL, mesh_s, dm, bc_d, _dev.mpc_dev(constraint)["struct"], mplan[0], vplan[0], keep,
            (mplan[1].get("interface_tiles", 0), mplan[1].get("tiles", 0)))
