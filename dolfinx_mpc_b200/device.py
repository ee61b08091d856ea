"""Device mirrors of the host descriptors and builders of the C-ABI structs.

PyTorch is used only as plumbing: device allocation, host<->device copies and
the current CUDA stream.  All arithmetic happens in libmpcx.so.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .fem import Form, Function, FunctionSpace, Integral, Mesh


def device() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.MpcxError("no CUDA device: dolfinx_mpc_b200 has no CPU fallback")
    return torch.device("cuda", int(os.environ.get("LOCAL_RANK", torch.cuda.current_device())))


def stream_ptr() -> int:
    return torch.cuda.current_stream(device()).cuda_stream


def to_dev(a: np.ndarray, pinned: bool = False) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(a))
    if pinned:
        t = t.pin_memory()
    return t.to(device(), non_blocking=pinned)


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()


# ---------------------------------------------------------------- mirrors (cached on the host objects)

def mesh_dev(mesh: Mesh) -> dict:
    d = mesh._dev
    if "x" not in d:
        x4 = np.zeros((mesh.x.shape[0], 4), dtype=np.float64)  # 32-byte rows: two 16-byte loads per vertex
        x4[:, :3] = mesh.x
        d["x"] = to_dev(x4)
        d["x_dofmap"] = to_dev(mesh.x_dofmap)
        d["struct"] = _lib.MeshS(ptr(d["x"]), ptr(d["x_dofmap"]), mesh.x.shape[0], mesh.x_dofmap.shape[1], 4)
    return d


def space_dev(V: FunctionSpace) -> dict:
    d = V._dev
    if "dofmap" not in d:
        # always a separate array, as with DOLFINx data where dof and geometry numbering differ
        d["dofmap"] = to_dev(V.dofmap)
    return d


def dofmap_struct(V: FunctionSpace, num_dofs: int) -> _lib.DofmapS:
    n_owned = V.index_map.size_local * V.bs
    return _lib.DofmapS(ptr(space_dev(V)["dofmap"]), V.nd, V.bs, num_dofs, n_owned if V.index_map.num_ghosts else 0)


def mpc_dev(mpc) -> dict:
    d = mpc._dev
    if "struct" not in d:
        mpc._not_finalized()
        d["is_slave"] = to_dev(mpc.is_slave)
        d["masters"] = to_dev(mpc.masters.array)
        d["coeffs"] = to_dev(mpc._coeff_map.array)
        d["offsets"] = to_dev(mpc.masters.offsets)
        d["c2s"] = to_dev(mpc.cell_to_slaves.array)
        d["c2s_off"] = to_dev(mpc.cell_to_slaves.offsets)
        d["slaves"] = to_dev(mpc.slaves)
        d["struct"] = _lib.MpcS(ptr(d["is_slave"]), ptr(d["masters"]), ptr(d["coeffs"]), ptr(d["offsets"]),
                                ptr(d["c2s"]), ptr(d["c2s_off"]), ptr(d["slaves"]), len(mpc.slaves),
                                mpc.num_local_slaves, len(mpc.is_slave))
    return d


_tables_cache: dict = {}


def tables_struct(tab, bs: int, bs1: int = 0):
    key = (tab.cell_type, tab.degree, tab.nq, bs, tab.nfacets, tab.degree1, tab.nd1, bs1, device().index)
    if key not in _tables_cache:
        keep = [to_dev(tab.weights), to_dev(tab.phi), to_dev(tab.dphi), to_dev(tab.gdphi)]
        ftan = to_dev(tab.ftan) if tab.nfacets else None
        phi1 = to_dev(tab.phi1) if tab.nd1 else None
        dphi1 = to_dev(tab.dphi1) if tab.nd1 else None
        keep += [ftan, phi1, dphi1]
        s = _lib.Tables(tab.tdim, tab.gdim, tab.nd, tab.ng, tab.nq, bs, *[ptr(k) for k in keep[:4]], tab.nfacets, ptr(ftan),
                        tab.nd1, bs1 if tab.nd1 else 0, ptr(phi1), ptr(dphi1))
        _tables_cache[key] = (s, keep)
    return _tables_cache[key][0]


def function_dev(f: Function) -> torch.Tensor:
    """Device copy of a coefficient; a tensor already placed in ``f.device_array`` is used as is."""
    da = getattr(f, "device_array", None)
    if da is not None:
        return da
    return to_dev(f.array)


def integral_struct(form: Form, it: Integral, mpcs, keep: list) -> _lib.IntegralS:
    """C struct for one integral; device arrays that must outlive the call are appended to ``keep``."""
    V = form.function_spaces[0]
    tab = tables_struct(form.tables(it), V.bs, form.function_spaces[-1].bs)
    s = _lib.IntegralS()
    s.kernel = int(it.kernel)
    s.tables = C.pointer(tab)
    d = it._dev
    if it.cells is not None and "cells" not in d:
        d["cells"] = to_dev(it.cells)
    s.cells = ptr(d.get("cells"))
    if it.local_facets is not None:  # exterior-facet integral: (cells[i], local_facets[i]) pairs
        if "local_facets" not in d:
            d["local_facets"] = to_dev(it.local_facets)
        s.local_facets = ptr(d["local_facets"])
    ncells = form.mesh.num_cells_local if it.cells is None else len(it.cells)
    s.num_cells = ncells
    if len(it.coefficients) == 1:
        f = it.coefficients[0]
        t = function_dev(f)
        keep.append(t)
        s.coeff_nodal = ptr(t)
        s.coeff_dofmap = ptr(space_dev(f.function_space)["dofmap"])
        s.coeff_nd = f.function_space.nd
        s.coeff_bs = f.function_space.bs
        s.cstride = f.function_space.nd * f.function_space.bs
    elif len(it.coefficients) > 1:
        w, cstride = form.pack_coefficients(it)
        t = to_dev(w)
        keep.append(t)
        s.coeffs = ptr(t)
        s.cstride = cstride
    s.num_constants = len(it.constants)
    for i, c in enumerate(it.constants):
        s.constants[i] = float(c)
    if it.custom is not None:
        n = 1
        for Vk in form.function_spaces:
            n *= Vk.nd * Vk.bs
        s.custom = it.custom.handle(n, form.mesh.x_dofmap.shape[1], int(s.cstride))
    # active cells holding a slave of either constraint, as positions in the active list
    key = ("slave_cells",) + tuple(id(m) for m in mpcs)
    if key not in d:
        active = form.active_cells(it)
        has = np.zeros(len(active), dtype=bool)
        for m in mpcs:
            has |= np.diff(m.cell_to_slaves.offsets)[active] > 0
        pos = np.flatnonzero(has).astype(np.int32)
        d[("keepalive",) + key] = tuple(mpcs)  # their id() keys the list: must not be reused while it is cached
        d[key] = (to_dev(pos) if len(pos) else None, len(pos))
    sc, nsc = d[key]
    # an empty list is passed as a non-null pointer to a dummy so the library knows the split is valid
    if sc is None:
        if "dummy" not in d:
            d["dummy"] = torch.zeros(1, dtype=torch.int32, device=device())
        sc = d["dummy"]
    s.slave_cells = ptr(sc)
    s.num_slave_cells = nsc
    return s


def backsubstitution(mpc, u, homogenize: bool):
    lib = _lib.load()
    md = mpc_dev(mpc)
    fn = lib.mpcx_homogenize_f64 if homogenize else lib.mpcx_backsubstitution_f64
    if isinstance(u, Function):
        t = to_dev(u.array)
        _lib.check(fn(C.byref(md["struct"]), ptr(t), stream_ptr()))
        u.array[:] = t.cpu().numpy()
    else:  # la.Vector or a raw device tensor
        t = getattr(u, "data", u)
        _lib.check(fn(C.byref(md["struct"]), ptr(t), stream_ptr()))
