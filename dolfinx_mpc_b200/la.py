"""Device-resident linear-algebra containers the assembly routines fill.

``Matrix`` is a scalar CSR (int64 ``row_ptr``, int32 ``col`` ascending per row,
float64 ``val``) in HBM -- what replaces the PETSc ``Mat`` behind
``mat_add_block_values`` / ``mat_add_values`` (``python/src/dolfinx_mpc/mpc.cpp:284-287``).
``Vector`` is the local (owned + ghost) array of a PETSc ``Vec``'s local form
(``python/src/dolfinx_mpc/assemble_vector.py:100-102``).  Both are handed back to
SciPy / numpy on the host only for checking.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np
import torch

from . import _lib
from . import device as _dev


class Matrix:
    def __init__(self, row_ptr, col, shape, bs=(1, 1), max_block_row: Optional[int] = None):
        """``row_ptr`` / ``col``: numpy arrays (uploaded) or device tensors (used as they are; the host copies that
        ``getValuesCSR`` / ``to_scipy`` hand out are downloaded on first use)."""
        self.shape = tuple(int(s) for s in shape)
        self.bs = tuple(bs)
        if isinstance(row_ptr, torch.Tensor):
            self.row_ptr, self.col = row_ptr, col
            self._row_ptr_host = self._col_host = None
            self.nnz = int(row_ptr[-1])
            if max_block_row is None:
                max_block_row = int((row_ptr[1:] - row_ptr[:-1]).max()) // max(1, self.bs[1]) if len(row_ptr) > 1 else 0
        else:
            self._row_ptr_host = np.ascontiguousarray(row_ptr, dtype=np.int64)
            self._col_host = np.ascontiguousarray(col, dtype=np.int32)
            self.nnz = int(self._row_ptr_host[-1])
            self.row_ptr = _dev.to_dev(self._row_ptr_host)
            self.col = _dev.to_dev(self._col_host)
            if max_block_row is None:
                max_block_row = int(np.diff(self._row_ptr_host).max(initial=0)) // max(1, self.bs[1])
        # capacity rounded up to an even count: the tile kernels add whole 16-byte runs (include/mpcx.h)
        self._val_storage = torch.zeros(self.nnz + (self.nnz & 1), dtype=torch.float64, device=_dev.device())
        self.val = self._val_storage[: self.nnz]
        self.max_block_row = max_block_row
        self._plans = {}
        self._tile_plans = {}
        self._row_plans = []
        self._slave_plans = {}
        self._keepalive = {}  # objects whose id() keys a cached plan: kept alive so that the id cannot be reused
        # "tile": entries combined per 512-cell tile in shared memory, one reduction per (tile, entry), where a
        # tile kernel exists; "atomic": one red.global.add per element entry
        self.scatter = os.environ.get("MPCX_SCATTER", "tile")
        self.ghost_exchange = None  # set by distributed.attach_ghost_exchange
        # zero-fill overlapped with the previous assembly on a side stream (second value buffer); off by default
        self.async_zero = os.environ.get("MPCX_ASYNC_ZERO", "0") == "1"
        self._spare = None
        self.deferred_errors = False
        self._err_flag = None

    @property
    def row_ptr_host(self) -> np.ndarray:
        if self._row_ptr_host is None:
            self._row_ptr_host = self.row_ptr.cpu().numpy()
        return self._row_ptr_host

    @property
    def col_host(self) -> np.ndarray:
        if self._col_host is None:
            self._col_host = self.col.cpu().numpy()
        return self._col_host

    def struct(self) -> _lib.CsrS:
        return _lib.CsrS(_dev.ptr(self.row_ptr), _dev.ptr(self.col), _dev.ptr(self.val), self.shape[0], self.nnz)

    def check_device_errors(self, stream_ptr: int):
        """Raise if a kernel reported an insertion outside the pattern.  Default: read the device flag now (one stream
        synchronisation per assembly, as immediate as the reference's exception).  With ``deferred_errors`` the flag is
        copied to pinned host memory in stream order and examined at the NEXT assembly into this matrix (and by
        ``synchronize()``): a time loop then never drains the device queue -- the mode for steady-state loops over
        cached plans, where every position was validated when the plans were built."""
        lib = _lib.load()
        if not self.deferred_errors:
            _lib.check(lib.mpcx_device_error(stream_ptr))
            return
        if self._err_flag is None:
            self._err_flag = torch.zeros(1, dtype=torch.int32).pin_memory()
        elif int(self._err_flag[0]) != 0:
            self.synchronize()
        _lib.check(lib.mpcx_device_error_async(self._err_flag.data_ptr(), stream_ptr))

    def synchronize(self):
        """Wait for the assembly stream and raise any deferred device error."""
        torch.cuda.current_stream(self.val.device).synchronize()
        if self._err_flag is not None and int(self._err_flag[0]) != 0:
            self._err_flag[0] = 0
            _lib.check(_lib.load().mpcx_device_error(_dev.stream_ptr()))  # reads and clears the flag, raises with the message

    def zeroEntries(self):
        if self.async_zero:
            self._swap_to_zeroed_buffer()
        else:  # cudaMemsetAsync through the library on the assembly stream (no framework fill kernel on the hot path)
            _lib.check(_lib.load().mpcx_zero_f64(_dev.ptr(self.val), self.val.numel(), _dev.stream_ptr()))

    def _swap_to_zeroed_buffer(self):
        """``async_zero``: two value buffers.  The assembly that starts now writes into the spare one, which was
        cleared on a side stream WHILE the previous assembly ran (the tile kernels leave most of the DRAM bandwidth
        unused); the buffer holding the previous values is cleared the same way now.  As with plain zeroing the
        previous values are gone once this returns -- but ``val`` is a different tensor after every call, so views
        taken earlier (``to_torch_sparse_csr``, ``dlpack``) must be taken again."""
        cur = torch.cuda.current_stream(self.val.device)
        if self._spare is None:
            self._spare = torch.zeros_like(self._val_storage)
            self._spare_ready = torch.cuda.Event()
            self._spare_ready.record(cur)
            self._zero_stream = torch.cuda.Stream(self.val.device)
        cur.wait_event(self._spare_ready)  # the spare buffer is clear
        old = self._val_storage
        self._val_storage, self.val = self._spare, self._spare[: self.nnz]
        done = torch.cuda.Event()
        done.record(cur)  # everything enqueued so far that reads the old values
        self._zero_stream.wait_event(done)
        with torch.cuda.stream(self._zero_stream):
            old.zero_()
            self._spare_ready = torch.cuda.Event()
            self._spare_ready.record(self._zero_stream)
        self._spare = old

    def values_storage(self) -> torch.Tensor:
        """The device buffer behind ``val`` (capacity rounded up to an even number of entries)."""
        return self._val_storage

    def bind_values(self, storage: torch.Tensor):
        """Assemble into another value buffer from now on (double-buffered outputs, a solver's own array ...):
        a 16-byte aligned contiguous float64 device tensor with room for ``nnz`` rounded up to an even count."""
        need = self.nnz + (self.nnz & 1)
        if (storage.dtype != torch.float64 or storage.dim() != 1 or not storage.is_contiguous() or storage.numel() < need
                or storage.data_ptr() % 16):
            raise ValueError(f"value storage must be a 16-byte aligned contiguous float64 tensor of >= {need} entries")
        self._val_storage = storage
        self.val = storage[: self.nnz]
        self.async_zero, self._spare = False, None  # the caller manages the buffers from here on

    def plan(self, form, integral) -> Optional[_lib.PlanS]:
        """Scatter plan for one integral of ``form`` into this pattern (built on first use, then cached)."""
        width = 1 if self.max_block_row <= 256 else (2 if self.max_block_row <= 65536 else 0)
        if width == 0:
            return None
        key = (id(form.function_spaces[0]), id(form.function_spaces[1]), id(integral))
        if key not in self._plans:
            self._keepalive[key] = (form.function_spaces, integral)
            V0, V1 = form.function_spaces
            ncells = form.mesh.num_cells_local if integral.cells is None else len(integral.cells)
            lpos = torch.empty(ncells * V0.nd * V1.nd, dtype=torch.uint8 if width == 1 else torch.int16,
                               device=_dev.device())
            d0 = _dev.dofmap_struct(V0, self.shape[0])
            d1 = _dev.dofmap_struct(V1, self.shape[1])
            if integral.cells is not None and "cells" not in integral._dev:
                integral._dev["cells"] = _dev.to_dev(integral.cells)
            lib = _lib.load()
            A = self.struct()
            _lib.check(lib.mpcx_build_plan(C.byref(d0), C.byref(d1), _dev.ptr(integral._dev.get("cells")), ncells,
                                           C.byref(A), _dev.ptr(lpos), width, _dev.stream_ptr()))
            _lib.check(lib.mpcx_device_error(_dev.stream_ptr()))
            self._plans[key] = (lpos, _lib.PlanS(_dev.ptr(lpos), width))
        return self._plans[key][1]

    def tile_plan(self, form, integral, s_integral, bc0_d, bc1_d, key_extra=(), keepalive=()):
        """Tile plan (csrc/mpcx_tile.cuh) for one integral into this pattern; built on the device on first use.
        Returns None when the element has no tile kernel."""
        V0, V1 = form.function_spaces
        tab = form.tables(integral)
        p1 = V0.nd == tab.tdim + 1 and tab.ng == tab.tdim + 1 and V0.bs == 1 and V1.bs == 1 and V1.nd == V0.nd
        if not p1 or int(integral.kernel) not in (0, 1, 4):
            return None
        key = (id(V0), id(V1), id(integral), _dev.ptr(bc0_d), _dev.ptr(bc1_d)) + tuple(key_extra)
        if key not in self._tile_plans:
            self._keepalive[key] = (V0, V1, integral, bc0_d, bc1_d, keepalive)
            lib = _lib.load()
            ncells = int(s_integral.num_cells)
            skip = None
            if s_integral.num_slave_cells > 0:
                skip = torch.zeros(ncells, dtype=torch.int8, device=_dev.device())
                skip[integral._dev[[k for k in integral._dev if isinstance(k, tuple) and k[0] == "slave_cells"
                                    and k[1:] == tuple(key_extra)][0]][0].long()] = 1
            d0 = _dev.dofmap_struct(V0, self.shape[0])
            d1 = _dev.dofmap_struct(V1, self.shape[1])
            mesh_s = _dev.mesh_dev(form.mesh)["struct"]
            A = self.struct()
            handle = C.c_void_p()
            try:
                _lib.check(lib.mpcx_tile_plan_create(C.byref(mesh_s), C.byref(d0), C.byref(d1), s_integral.cells, ncells,
                                                     _dev.ptr(skip), _dev.ptr(bc0_d), _dev.ptr(bc1_d), C.byref(A),
                                                     _dev.stream_ptr(), C.byref(handle)))
            except _lib.MpcxError as e:
                # a mesh whose tiles do not fit the plan format (e.g. a vertex shared by thousands of cells) is
                # assembled by the atomic-scatter kernels instead -- still on the device
                if getattr(e, "status", None) != _lib.ERR_UNSUPPORTED:
                    raise
                self._tile_plans[key] = None
                return None
            _lib.check(lib.mpcx_device_error(_dev.stream_ptr()))
            if s_integral.num_slave_cells > 0 and len(keepalive) == 2 and os.environ.get("MPCX_SLAVE_PLAN", "1") != "0":
                # scatter plan of the cells holding slaves: no constraint lookups / row searches at assembly time
                m0 = _dev.mpc_dev(keepalive[0])["struct"]
                m1 = _dev.mpc_dev(keepalive[1])["struct"]
                _lib.check(lib.mpcx_tile_plan_add_slave_cells(handle, C.byref(s_integral), C.byref(d0), C.byref(d1),
                                                              _dev.ptr(bc0_d), _dev.ptr(bc1_d), C.byref(m0), C.byref(m1),
                                                              C.byref(A), _dev.stream_ptr()))
                _lib.check(lib.mpcx_device_error(_dev.stream_ptr()))
            info = (C.c_int64 * 16)()
            lib.mpcx_tile_plan_info(handle, info, 16)
            self._tile_plans[key] = (handle, dict(zip(("tiles", "cells_per_tile", "bulk_cells", "max_nodes",
                                                       "max_dests", "tile_nodes", "dests", "bytes", "max_slots", "slots", "max_runs", "runs", "max_stage", "symmetric", "interface_tiles", "stage_slots"),
                                                      [int(v) for v in info])))
        return self._tile_plans[key]

    def row_plan(self, form, integral, s_integral, key_extra=(), keepalive=()):
        """Row plan (csrc/mpcx_rowgather.cuh) of one integral into this pattern, built on the device on first use;
        None when the element has no row-gather kernel (elasticity with bs == gdim on P1 simplices and P2 tetrahedra has)."""
        V0, V1 = form.function_spaces
        tab = form.tables(integral)
        p1 = V0.nd == tab.tdim + 1
        p2tet = tab.tdim == 3 and V0.nd == 10
        if not (V0 is V1 and int(integral.kernel) == 2 and integral.integral_type == "cell" and (p1 or p2tet)
                and tab.ng == tab.tdim + 1 and V0.bs == tab.tdim):
            return None
        key = ("row", id(V0), id(integral)) + tuple(key_extra)
        if key not in self._tile_plans:
            self._keepalive[key] = (V0, integral, keepalive)
            lib = _lib.load()
            ncells = int(s_integral.num_cells)
            skip = None
            if s_integral.num_slave_cells > 0:
                skip = torch.zeros(ncells, dtype=torch.int8, device=_dev.device())
                skip[integral._dev[[k for k in integral._dev if isinstance(k, tuple) and k[0] == "slave_cells"
                                    and k[1:] == tuple(key_extra)][0]][0].long()] = 1
            d0 = _dev.dofmap_struct(V0, self.shape[0])
            A = self.struct()
            handle = C.c_void_p()
            try:
                _lib.check(lib.mpcx_row_plan_create(C.byref(d0), s_integral.cells, ncells, _dev.ptr(skip), C.byref(A),
                                                    _dev.stream_ptr(), C.byref(handle)))
                _lib.check(lib.mpcx_device_error(_dev.stream_ptr()))
            except _lib.MpcxError as e:  # e.g. a node shared by more than 255 cells: the atomic-scatter kernels take over
                if getattr(e, "status", None) != _lib.ERR_UNSUPPORTED:
                    raise
                if handle:
                    lib.mpcx_row_plan_destroy(handle)
                self._tile_plans[key] = None
                return None
            self._row_plans.append(handle)
            self._tile_plans[key] = (handle, {"row_plan": 1})
        return self._tile_plans[key]

    def slave_plan(self, form, integral, s_integral, bc0_d, bc1_d, mpc0, mpc1):
        """Scatter plan of the cells of ``integral`` that hold slaves (any element), built on the device on first use."""
        key = ("slave", id(integral), id(mpc0), id(mpc1), _dev.ptr(bc0_d), _dev.ptr(bc1_d))
        if key not in self._slave_plans:
            self._keepalive[key] = (integral, mpc0, mpc1, bc0_d, bc1_d)
            lib = _lib.load()
            V0, V1 = form.function_spaces
            d0 = _dev.dofmap_struct(V0, self.shape[0])
            d1 = _dev.dofmap_struct(V1, self.shape[1])
            m0 = _dev.mpc_dev(mpc0)["struct"]
            m1 = _dev.mpc_dev(mpc1)["struct"]
            A = self.struct()
            handle = C.c_void_p()
            _lib.check(lib.mpcx_slave_plan_create(C.byref(s_integral), C.byref(d0), C.byref(d1), _dev.ptr(bc0_d),
                                                  _dev.ptr(bc1_d), C.byref(m0), C.byref(m1), C.byref(A),
                                                  _dev.stream_ptr(), C.byref(handle)))
            _lib.check(lib.mpcx_device_error(_dev.stream_ptr()))
            self._slave_plans[key] = handle
        return self._slave_plans[key]

    def __del__(self):
        try:
            lib = _lib.load()
            for h in self._slave_plans.values():
                lib.mpcx_slave_plan_destroy(h)
            rows = set(h.value for h in self._row_plans)
            for entry in self._tile_plans.values():
                if entry is not None and entry[0].value not in rows:
                    lib.mpcx_tile_plan_destroy(entry[0])
            for h in self._row_plans:
                lib.mpcx_row_plan_destroy(h)
        except Exception:
            pass

    def assemble(self):
        """Finish assembly: with several ranks, send ghost-row values to their owners and add
        (PETSc ``MatAssemblyBegin/End`` in ``python/src/dolfinx_mpc/assemble_matrix.py:64``)."""
        if self.ghost_exchange is not None:
            self.ghost_exchange.reduce_matrix(self)

    # -- hand-off to the host, for checks only ---------------------------------------------------------------
    def getValuesCSR(self):
        return self.row_ptr_host, self.col_host, self.val.cpu().numpy()

    # -- zero-copy hand-off on the device --------------------------------------------------------------------
    def to_torch_sparse_csr(self):
        """The matrix as a ``torch.sparse_csr_tensor`` sharing ``col`` and ``val`` with this object (cuSPARSE SpMV /
        SpMM through ``A_t @ x``); only the row pointer is narrowed to int32 (n + 1 entries) because torch wants
        both index arrays of one dtype.  Counterpart of handing the assembled ``Mat`` to a solver
        (``python/src/dolfinx_mpc/problem.py:539-582``)."""
        if self.nnz >= 2**31:
            raise RuntimeError("more than 2^31 entries: use dlpack() and 64-bit row pointers")
        import warnings

        crow = self.row_ptr.to(torch.int32)
        with warnings.catch_warnings():
            warnings.filterwarnings("ignore", message="Sparse CSR tensor support is in beta state")
            return torch.sparse_csr_tensor(crow, self.col, self.val, size=self.shape, device=self.val.device)

    def dlpack(self):
        """DLPack capsules ``(row_ptr int64, col int32, val float64)`` of the device arrays -- zero-copy import into
        CuPy (``cupyx.scipy.sparse.csr_matrix``), PETSc (``MatCreateSeqAIJCUSPARSE`` after narrowing ``row_ptr``)
        or any other DLPack consumer."""
        from torch.utils.dlpack import to_dlpack

        return to_dlpack(self.row_ptr), to_dlpack(self.col), to_dlpack(self.val)

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csr_matrix((self.val.cpu().numpy(), self.col_host, self.row_ptr_host), shape=self.shape)

    def norm(self) -> float:
        return float(torch.linalg.vector_norm(self.val))


class Vector:
    def __init__(self, n: int, data: Optional[torch.Tensor] = None):
        """``data``: an external device tensor of length ``n`` to assemble into.  The tile kernels need a 16-byte
        aligned buffer; an external tensor that is not (an odd-offset view) is still valid -- ``tile_ok`` is then
        False and ``assemble_vector`` uses the atomic-scatter kernel."""
        if data is None:  # capacity rounded up to an even count (include/mpcx.h)
            self._storage = torch.zeros(n + (n & 1), dtype=torch.float64, device=_dev.device())
            data = self._storage[:n]
        else:
            if data.dtype != torch.float64 or data.dim() != 1 or data.numel() != n or not data.is_contiguous():
                raise ValueError("Vector data must be a contiguous float64 tensor of length n")
        self.data = data
        self.ghost_exchange = None

    def bind(self, data: torch.Tensor):
        """Assemble into ``data`` from now on (same length, float64, contiguous)."""
        if data.dtype != torch.float64 or data.dim() != 1 or data.numel() != self.data.numel() or not data.is_contiguous():
            raise ValueError("Vector.bind needs a contiguous float64 tensor of the same length")
        self.data = data

    @property
    def tile_ok(self) -> bool:
        return self.data.data_ptr() % 16 == 0

    def set(self, v: float):
        if v == 0.0 and self.data.is_cuda:
            _lib.check(_lib.load().mpcx_zero_f64(_dev.ptr(self.data), self.data.numel(), _dev.stream_ptr()))
        else:
            self.data.fill_(v)

    @property
    def array(self) -> np.ndarray:
        return self.data.cpu().numpy()

    def ghostUpdate(self):
        """``VecGhostUpdate(ADD_VALUES, SCATTER_REVERSE)`` (e.g. ``python/tests/test_vector_assembly.py:51``)."""
        if self.ghost_exchange is not None:
            self.ghost_exchange.reduce_vector(self)

    def norm(self) -> float:
        return float(torch.linalg.vector_norm(self.data))
