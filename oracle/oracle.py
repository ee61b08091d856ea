"""ctypes front-end of the CPU oracle (``oracle/mpc_oracle.c``) -- TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs import this module.  The product package
``dolfinx_mpc_b200`` never does.

Besides the C restatement (oracle 1) this module holds the reference's own test
method as an independent second oracle (oracle 2): the global prolongation ``K``
and the identities ``K^H A K == A_mpc[free, free]``, ``K^H b == b_mpc[free]``,
``b_mpc[slaves] == 0`` (``python/src/dolfinx_mpc/utils/test.py:67-149,202-265``).

Parity status: element-tensor values are "parity unpinned" (see the header of
mpc_oracle.c); elimination + scatter are pinned through the identities above.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libmpc_oracle.so")


def build(force: bool = False, march: Optional[str] = None, out: Optional[str] = None) -> str:
    """Compile the oracle with gcc (``make -C oracle``)."""
    target = out or LIB_PATH
    src = os.path.join(_HERE, "mpc_oracle.c")
    if force or not os.path.exists(target) or os.path.getmtime(target) < os.path.getmtime(src):
        cmd = ["make", "-C", _HERE, "-B" if force else "-s"]
        if march:
            cmd.append(f"MARCH={march}")
        if out:
            cmd.append(f"OUT={os.path.relpath(out, _HERE)}")
        subprocess.run(cmd, check=True, capture_output=True)
    return target


class _Tables(C.Structure):
    _fields_ = [("tdim", C.c_int32), ("gdim", C.c_int32), ("nd", C.c_int32), ("ng", C.c_int32),
                ("nq", C.c_int32), ("bs", C.c_int32), ("weights", C.c_void_p), ("phi", C.c_void_p),
                ("dphi", C.c_void_p), ("gdphi", C.c_void_p), ("nfacets", C.c_int32), ("ftan", C.c_void_p),
                ("nd1", C.c_int32), ("bs1", C.c_int32), ("phi1", C.c_void_p), ("dphi1", C.c_void_p)]


class _Mpc(C.Structure):
    _fields_ = [("is_slave", C.c_void_p), ("masters", C.c_void_p), ("coeffs", C.c_void_p),
                ("offsets", C.c_void_p), ("c2s", C.c_void_p), ("c2s_offsets", C.c_void_p),
                ("slaves", C.c_void_p), ("num_slaves", C.c_int32), ("num_local_slaves", C.c_int32)]


class _Csr(C.Structure):
    _fields_ = [("row_ptr", C.c_void_p), ("col", C.c_void_p), ("val", C.c_void_p), ("num_rows", C.c_int64)]


class _Mesh(C.Structure):
    _fields_ = [("x", C.c_void_p), ("x_dofmap", C.c_void_p), ("ng", C.c_int32)]


class _Dofmap(C.Structure):
    _fields_ = [("map", C.c_void_p), ("nd", C.c_int32), ("bs", C.c_int32)]


_libs = {}


def lib(path: Optional[str] = None):
    path = path or LIB_PATH
    if path not in _libs:
        if not os.path.exists(path):
            build(out=None if path == LIB_PATH else path)
        _libs[path] = C.CDLL(path)
    return _libs[path]


def _a(a):
    """address for a struct field"""
    return None if a is None else a.ctypes.data


def _p(a):
    """pointer argument of a foreign call (never a bare int: ctypes would truncate it to 32 bits)"""
    return None if a is None else C.c_void_p(a.ctypes.data)


def _tables(tab, bs, bs1=0):
    return _Tables(tab.tdim, tab.gdim, tab.nd, tab.ng, tab.nq, bs, _a(tab.weights), _a(tab.phi), _a(tab.dphi),
                   _a(tab.gdphi), tab.nfacets, _a(tab.ftan) if tab.nfacets else None,
                   tab.nd1, bs1 if tab.nd1 else 0, _a(tab.phi1) if tab.nd1 else None, _a(tab.dphi1) if tab.nd1 else None)


def tabulate(kernel: int, tab, bs: int, X: np.ndarray, w=None, c=(1.0,), bs1: int = 0) -> np.ndarray:
    """One element tensor from the oracle kernels (``bs1``: block size of the trial element of a rectangular form)."""
    n = tab.nd * bs
    n1 = tab.nd1 * bs1 if tab.nd1 else n
    size = n if kernel == 3 else n * n1
    out = np.zeros(size)
    X = np.ascontiguousarray(X, dtype=np.float64)
    w = None if w is None else np.ascontiguousarray(w, dtype=np.float64)
    c = np.ascontiguousarray(c, dtype=np.float64)
    t = _tables(tab, bs, bs1)
    rc = lib().orc_tabulate(int(kernel), C.byref(t), _p(w), _p(c), _p(X), _p(out), size)
    assert rc == 0, rc
    return out if kernel == 3 else out.reshape(n, n1)


class OracleMPC:
    """Packed constraint data built by the oracle's restatement of the reference constructor
    (``cpp/MultiPointConstraint.h:36-126``)."""

    def __init__(self, V, slaves, masters_local, coeffs, owners, offsets):
        L = lib()
        slaves = np.ascontiguousarray(slaves, dtype=np.int32)
        masters_local = np.ascontiguousarray(masters_local, dtype=np.int32)
        coeffs = np.ascontiguousarray(coeffs, dtype=np.float64)
        owners = np.ascontiguousarray(owners, dtype=np.int32)
        offsets = np.ascontiguousarray(offsets, dtype=np.int32)
        self.V = V
        nd_ = V.num_dofs
        ns = len(slaves)
        nc = V.mesh.num_cells_local
        self.is_slave = np.zeros(nd_, np.int8)
        self.offsets = np.zeros(nd_ + 1, np.int32)
        nm = len(masters_local)
        self.masters = np.zeros(nm, np.int32)
        self.coeffs = np.zeros(nm, np.float64)
        self.owners = np.zeros(nm, np.int32)
        self.slaves = np.zeros(ns, np.int32)
        self.c2s_offsets = np.zeros(nc + 1, np.int32)
        nls = C.c_int32(0)
        c2s = C.POINTER(C.c_int32)()
        dm = np.ascontiguousarray(V.dofmap[:nc])
        rc = L.orc_mpc_build(C.c_int32(nd_), C.c_int32(V.index_map.size_local * V.bs), C.c_int32(ns), _p(slaves),
                             _p(masters_local), _p(coeffs), _p(owners), _p(offsets), _p(dm), C.c_int32(nc),
                             C.c_int32(V.nd), C.c_int32(V.bs), _p(self.is_slave), _p(self.offsets),
                             _p(self.masters), _p(self.coeffs), _p(self.owners), _p(self.slaves), C.byref(nls),
                             _p(self.c2s_offsets), C.byref(c2s))
        assert rc == 0
        n = int(self.c2s_offsets[-1])
        self.c2s = np.ctypeslib.as_array(c2s, shape=(max(n, 1),))[:n].copy()
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_free(c2s)
        self.num_local_slaves = nls.value

    @classmethod
    def empty(cls, V):
        return cls(V, [], [], [], [], [0])

    def struct(self):
        return _Mpc(_a(self.is_slave), _a(self.masters), _a(self.coeffs), _a(self.offsets), _a(self.c2s),
                    _a(self.c2s_offsets), _a(self.slaves), len(self.slaves), self.num_local_slaves)


class WrappedMPC:
    """Oracle view of already packed constraint arrays (e.g. a rank-local constraint whose masters were mapped
    to the extended local numbering by the caller, cpp/MultiPointConstraint.h:117-125)."""

    def __init__(self, V, is_slave, masters, coeffs, offsets, c2s, c2s_offsets, slaves, num_local_slaves):
        self.V = V
        self.is_slave = np.ascontiguousarray(is_slave, np.int8)
        self.masters = np.ascontiguousarray(masters, np.int32)
        self.coeffs = np.ascontiguousarray(coeffs, np.float64)
        self.offsets = np.ascontiguousarray(offsets, np.int32)
        self.c2s = np.ascontiguousarray(c2s, np.int32)
        self.c2s_offsets = np.ascontiguousarray(c2s_offsets, np.int32)
        self.slaves = np.ascontiguousarray(slaves, np.int32)
        self.num_local_slaves = int(num_local_slaves)

    struct = OracleMPC.struct


def mpc_from_arrays(V, data) -> OracleMPC:
    """From ``add_constraint``-style arrays with GLOBAL masters on a serial space (global == local)."""
    slaves, masters, coeffs, owners, offsets = data
    return OracleMPC(V, slaves, np.asarray(masters, dtype=np.int64).astype(np.int32), coeffs, owners, offsets)


def create_pattern(form, m0: OracleMPC, m1: OracleMPC):
    L = lib()
    V0, V1 = form.function_spaces
    d0 = _Dofmap(_a(V0.dofmap), V0.nd, V0.bs)
    d1 = _Dofmap(_a(V1.dofmap), V1.nd, V1.bs)
    nrows = m0.V.num_dofs
    rp = C.POINTER(C.c_int64)()
    cl = C.POINTER(C.c_int32)()
    nc = form.mesh.num_cells_local
    s0, s1 = m0.struct(), m1.struct()
    rc = L.orc_create_pattern(C.byref(d0), C.byref(d1), None, C.c_int64(nc), C.c_int32(nc), C.byref(s0),
                              C.byref(s1), C.c_int64(nrows), C.byref(rp), C.byref(cl))
    assert rc == 0
    row_ptr = np.ctypeslib.as_array(rp, shape=(nrows + 1,)).copy()
    nnz = int(row_ptr[-1])
    col = np.ctypeslib.as_array(cl, shape=(max(nnz, 1),))[:nnz].copy()
    L.orc_free.argtypes = [C.c_void_p]
    L.orc_free(rp)
    L.orc_free(cl)
    return row_ptr, col


def _pack_coefficients(form, it, cells, n, libpath=None):
    """``pack_coefficients`` in C (the reference calls DOLFINx's C++ routine inside the timed path)."""
    if not it.coefficients:
        return None, 0
    L = lib(libpath)
    cstride = sum(f.function_space.nd * f.function_space.bs for f in it.coefficients)
    w = np.empty((n, cstride))
    off = 0
    for f in it.coefficients:
        Vf = f.function_space
        d = _Dofmap(_a(Vf.dofmap), Vf.nd, Vf.bs)
        L.orc_pack_coefficient(_p(f.array), C.byref(d), _p(cells), C.c_int64(n), _p(w), C.c_int(cstride), C.c_int(off))
        off += Vf.nd * Vf.bs
    return w, cstride


def _integral_args(form, it, libpath=None):
    tab = form.tables(it)
    t = _tables(tab, form.function_spaces[0].bs, form.function_spaces[-1].bs)
    cells = it.cells
    n = form.mesh.num_cells_local if cells is None else len(cells)
    w, cstride = _pack_coefficients(form, it, cells, n, libpath)
    return t, tab, w, cstride, cells, n


def _bc_markers(V, bcs, n):
    mine = [bc for bc in bcs if bc.function_space is V or bc.function_space.dofmap is V.dofmap]
    if not mine:
        return None
    m = np.zeros(n, np.int8)
    for bc in mine:
        bc.mark_dofs(m)
    return m


def assemble_matrix(form, m0: OracleMPC, m1: Optional[OracleMPC] = None, bcs=(), diagval=1.0, pattern=None,
                    same_space: Optional[bool] = None, libpath: Optional[str] = None):
    """Oracle for ``dolfinx_mpc.assemble_matrix`` (C++ driver ``cpp/assemble_matrix.cpp:662-726`` plus the
    Python post-steps ``python/src/dolfinx_mpc/assemble_matrix.py:51-62``).  Returns ``(row_ptr, col, val)``."""
    L = lib(libpath)
    if same_space is None:
        same_space = m1 is None or m1 is m0
    m1 = m0 if m1 is None else m1
    row_ptr, col = pattern if pattern is not None else create_pattern(form, m0, m1)
    val = np.zeros(int(row_ptr[-1]))
    A = _Csr(_a(row_ptr), _a(col), _a(val), len(row_ptr) - 1)
    V0, V1 = form.function_spaces
    mesh = form.mesh
    ms = _Mesh(_a(mesh.x), _a(mesh.x_dofmap), mesh.x_dofmap.shape[1])
    d0 = _Dofmap(_a(V0.dofmap), V0.nd, V0.bs)
    d1 = _Dofmap(_a(V1.dofmap), V1.nd, V1.bs)
    bc0 = _bc_markers(V0, bcs, len(row_ptr) - 1)
    bc1 = _bc_markers(V1, bcs, m1.V.num_dofs)
    s0, s1 = m0.struct(), m1.struct()
    for it in form.integrals:
        t, tab, w, cstride, cells, n = _integral_args(form, it)
        rc = L.orc_assemble_cells_matrix(int(it.kernel), C.byref(t), C.byref(ms), _p(cells), C.c_int64(n), _p(w),
                                         C.c_int(cstride), _p(it.constants), C.byref(d0), C.byref(d1), _p(bc0),
                                         _p(bc1), C.byref(s0), C.byref(s1), C.byref(A), _p(it.local_facets))
        assert rc == 0, f"oracle matrix assembly failed with {rc}"
    L.orc_add_diagonal.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double]
    if same_space and m0.num_local_slaves > 0:
        rc = L.orc_add_diagonal(C.byref(A), _p(m0.slaves), m0.num_local_slaves, float(diagval))
        assert rc == 0
    if V0 is V1:
        n_owned = V0.index_map.size_local * V0.bs
        for bc in bcs:
            if bc.function_space is V0 or bc.function_space.dofmap is V0.dofmap:
                dofs = np.ascontiguousarray(bc.dofs[bc.dofs < n_owned])
                if len(dofs):
                    rc = L.orc_add_diagonal(C.byref(A), _p(dofs), len(dofs), float(diagval))
                    assert rc == 0
    return row_ptr, col, val


def assemble_vector(form, m: OracleMPC, b: Optional[np.ndarray] = None, libpath: Optional[str] = None):
    """Oracle for ``dolfinx_mpc.assemble_vector`` (zeroing of ``assemble_vector.py:101`` included)."""
    L = lib(libpath)
    V = form.function_spaces[0]
    b = np.zeros(m.V.num_dofs) if b is None else b
    b[:] = 0.0
    mesh = form.mesh
    ms = _Mesh(_a(mesh.x), _a(mesh.x_dofmap), mesh.x_dofmap.shape[1])
    d = _Dofmap(_a(V.dofmap), V.nd, V.bs)
    s = m.struct()
    for it in form.integrals:
        t, tab, w, cstride, cells, n = _integral_args(form, it)
        rc = L.orc_assemble_cells_vector(int(it.kernel), C.byref(t), C.byref(ms), _p(cells), C.c_int64(n), _p(w),
                                         C.c_int(cstride), _p(it.constants), C.byref(d), C.byref(s), _p(b),
                                         _p(it.local_facets))
        assert rc == 0
    return b


def apply_lifting(b: np.ndarray, forms: Sequence, bcs: Sequence[Sequence], m: OracleMPC, x0=None, scale=1.0):
    L = lib()
    L.orc_apply_lifting_cells.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    s = m.struct()
    for j, a in enumerate(forms):
        if a is None:
            continue
        V0, V1 = a.function_spaces
        markers = np.zeros(V1.num_dofs, np.int8)
        values = np.zeros(V1.num_dofs)
        for bc in bcs[j]:
            bc.mark_dofs(markers)
            bc.set(values)
        mesh = a.mesh
        ms = _Mesh(_a(mesh.x), _a(mesh.x_dofmap), mesh.x_dofmap.shape[1])
        d0 = _Dofmap(_a(V0.dofmap), V0.nd, V0.bs)
        d1 = _Dofmap(_a(V1.dofmap), V1.nd, V1.bs)
        x0j = None if not x0 else np.ascontiguousarray(x0[j], dtype=np.float64)
        for it in a.integrals:
            t, tab, w, cstride, cells, n = _integral_args(a, it)
            rc = L.orc_apply_lifting_cells(int(it.kernel), C.addressof(t), C.addressof(ms), _p(cells), n, _p(w),
                                           cstride, _p(it.constants), C.addressof(d0), C.addressof(d1), _p(markers),
                                           _p(values), _p(x0j), float(scale), C.addressof(s), _p(b), _p(it.local_facets))
            assert rc == 0
    return b


def backsubstitution(m: OracleMPC, u: np.ndarray):
    s = m.struct()
    lib().orc_backsubstitution(C.byref(s), _p(u))
    return u


def homogenize(m: OracleMPC, u: np.ndarray):
    s = m.struct()
    lib().orc_homogenize(C.byref(s), _p(u))
    return u


# ------------------------------------------------------------------ oracle 2: the reference's own test method

def transformation_matrix(num_dofs: int, slaves, masters, coeffs, offsets) -> sp.csr_matrix:
    """Global prolongation K (num_dofs x num_free) as in ``utils/test.py:67-149``: identity on free dofs, row
    ``slave -> coeffs`` at the renumbered master columns; a slave without masters keeps an identity column."""
    slaves = np.asarray(slaves, dtype=np.int64)
    offsets = np.asarray(offsets)
    all_removed = np.sort(slaves)
    rows, cols, vals = [], [], []
    for i, s in enumerate(slaves):
        ms = np.asarray(masters[offsets[i]:offsets[i + 1]], dtype=np.int64)
        if len(ms):
            for mdof, cf in zip(ms, coeffs[offsets[i]:offsets[i + 1]]):
                rows.append(s)
                cols.append(mdof - np.searchsorted(all_removed, mdof))
                vals.append(cf)
        else:
            rows.append(s)
            cols.append(s - np.searchsorted(all_removed, s))
            vals.append(1.0)
    free = np.setdiff1d(np.arange(num_dofs), all_removed)
    rows += list(free)
    cols += list(free - np.searchsorted(all_removed, free))
    vals += [1.0] * len(free)
    ncols = num_dofs - len(all_removed)
    return sp.coo_matrix((vals, (rows, cols)), shape=(num_dofs, max(ncols, int(max(cols)) + 1 if cols else 0))).tocsr()


def compare_mpc_lhs(A_org: sp.spmatrix, A_mpc: sp.spmatrix, K: sp.spmatrix, slaves, atol=5e-12):
    """``utils/test.py:202-242``: max |K^H A K - A_mpc[free, free]| < atol * scale."""
    KTAK = (K.T.conj() @ A_org @ K).tocsr()
    n = A_mpc.shape[0]
    free = np.setdiff1d(np.arange(n), np.asarray(slaves))
    red = A_mpc.tocsr()[free, :][:, free]
    if KTAK.shape != red.shape:
        KTAK = KTAK[: red.shape[0], : red.shape[1]]
    diff = abs(KTAK - red)
    scale = max(1.0, abs(A_org).max())
    assert diff.max() < atol * scale, f"K^T A K mismatch: {diff.max()} (scale {scale})"


def compare_mpc_rhs(b_org: np.ndarray, b_mpc: np.ndarray, K: sp.spmatrix, slaves):
    """``utils/test.py:245-265``."""
    n = len(b_mpc)
    free = np.setdiff1d(np.arange(n), np.asarray(slaves))
    red = K.T.conj() @ b_org
    assert np.allclose(b_mpc[np.asarray(slaves, dtype=np.int64)], 0)
    assert np.allclose(b_mpc[free], red[: len(free)])


# ------------------------------------------------------------------ full-size runs on all host threads

def _run_colored(chunks, fn, nthreads):
    """Run ``fn(chunk)`` for every chunk on a thread pool, even-numbered chunks first, then the odd ones.  The caller
    guarantees that chunks i and i + 2 touch disjoint rows (z-slabs of whole cube layers of a structured box), so
    the threads of one colour never write the same matrix row / vector entry."""
    import concurrent.futures as cf

    with cf.ThreadPoolExecutor(max_workers=nthreads) as ex:
        for colour in (0, 1):
            list(ex.map(fn, chunks[colour::2]))


def assemble_system_slabs(a, Lf, m, bcs, pattern, cells_per_layer: int, nthreads: int):
    """Matrix, load vector and lifting of a structured box problem by the SAME oracle routines as
    ``assemble_matrix`` / ``assemble_vector`` / ``apply_lifting``, with the cells cut into z-slabs of whole cube
    layers that are assembled concurrently (two colours, see ``_run_colored``) into ONE shared CSR / vector: the
    full-size (10^8 cells) comparison of tests/test_gpu_parity.py finishes in seconds instead of a minute.
    Single cell integral per form.  Returns ``(val, b)``."""
    L_ = lib()
    (ita,), (itL,) = a.integrals, Lf.integrals
    assert ita.cells is None and itL.cells is None and ita.integral_type == "cell" and itL.integral_type == "cell"
    row_ptr, col = pattern
    val = np.zeros(int(row_ptr[-1]))
    A = _Csr(_a(row_ptr), _a(col), _a(val), len(row_ptr) - 1)
    V = a.function_spaces[0]
    mesh = a.mesh
    ms = _Mesh(_a(mesh.x), _a(mesh.x_dofmap), mesh.x_dofmap.shape[1])
    d = _Dofmap(_a(V.dofmap), V.nd, V.bs)
    bc = _bc_markers(V, bcs, len(row_ptr) - 1)
    s = m.struct()
    nc = mesh.num_cells_local
    nlayers = nc // cells_per_layer
    assert nlayers * cells_per_layer == nc
    nchunks = max(1, min(nlayers, 2 * nthreads))
    bounds = (np.linspace(0, nlayers, nchunks + 1).astype(np.int64)) * cells_per_layer
    chunks = [np.arange(bounds[i], bounds[i + 1], dtype=np.int32) for i in range(nchunks) if bounds[i + 1] > bounds[i]]
    ta = _tables(a.tables(ita), V.bs)
    assert not ita.coefficients, "assemble_system_slabs: bilinear form without coefficients"

    def mat(cells):
        rc = L_.orc_assemble_cells_matrix(int(ita.kernel), C.byref(ta), C.byref(ms), _p(cells), C.c_int64(len(cells)),
                                          None, C.c_int(0), _p(ita.constants), C.byref(d), C.byref(d), _p(bc), _p(bc),
                                          C.byref(s), C.byref(s), C.byref(A), None)
        assert rc == 0, f"oracle matrix assembly failed with {rc}"

    _run_colored(chunks, mat, nthreads)
    L_.orc_add_diagonal.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_double]
    if m.num_local_slaves > 0:
        assert L_.orc_add_diagonal(C.byref(A), _p(m.slaves), m.num_local_slaves, 1.0) == 0
    n_owned = V.index_map.size_local * V.bs
    for bc_ in bcs:
        dofs = np.ascontiguousarray(bc_.dofs[bc_.dofs < n_owned])
        if len(dofs):
            assert L_.orc_add_diagonal(C.byref(A), _p(dofs), len(dofs), 1.0) == 0

    b = np.zeros(m.V.num_dofs)
    tl = _tables(Lf.tables(itL), V.bs)

    def vec(cells):
        w, cstride = _pack_coefficients(Lf, itL, cells, len(cells))
        rc = L_.orc_assemble_cells_vector(int(itL.kernel), C.byref(tl), C.byref(ms), _p(cells), C.c_int64(len(cells)),
                                          _p(w), C.c_int(cstride), _p(itL.constants), C.byref(d), C.byref(s), _p(b), None)
        assert rc == 0

    _run_colored(chunks, vec, nthreads)
    if bcs:
        L_.orc_apply_lifting_cells.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
        markers = np.zeros(V.num_dofs, np.int8)
        values = np.zeros(V.num_dofs)
        for bc_ in bcs:
            bc_.mark_dofs(markers)
            bc_.set(values)

        def lift(cells):
            rc = L_.orc_apply_lifting_cells(int(ita.kernel), C.addressof(ta), C.addressof(ms), _p(cells), len(cells), None, 0,
                                            _p(ita.constants), C.addressof(d), C.addressof(d), _p(markers), _p(values),
                                            None, 1.0, C.addressof(s), _p(b), None)
            assert rc == 0

        _run_colored(chunks, lift, nthreads)
    return val, b
