"""Host-facing I/O of an assembly step: pinned host copies of the step's input VALUES and of its outputs.

The reference keeps its data on the host, so a caller that switches to this package moves, per step, exactly
what changes between steps -- vertex coordinates, coefficient arrays, constraint coefficients in; CSR values and
the right-hand side out -- while the topology (dofmaps, constraint structure), the sparsity pattern and the tile
plans derived from them stay on the device, like the reference's ``Form`` / ``FunctionSpace`` / cached ``Mat``
(``python/src/dolfinx_mpc/assemble_matrix.py:49-51``).  ``StepIO`` owns the pinned buffers and the two copy
streams; it uses only public methods of ``Matrix`` / ``Vector`` (``bind_values`` / ``bind``) and the device
mirrors of ``device.py``.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import device as _dev


def _pinned_like(t: torch.Tensor, n: Optional[int] = None) -> torch.Tensor:
    return torch.empty(t.numel() if n is None else n, dtype=t.dtype).pin_memory()


class StepIO:
    """``download``: "full" = CSR values + RHS; "upper" = values of the upper triangle (row <= col; the constrained
    matrix of a symmetric form is symmetric, ``K^T A K``) + RHS, compacted on the device; "norms" = only two
    norms travel (the system is consumed on the device through ``Matrix.dlpack`` / ``to_torch_sparse_csr``)."""

    def __init__(self, mesh, coefficients: Sequence, mpc, A, b, download: str = "full"):
        assert download in ("full", "upper", "norms")
        self.A, self.b, self.mode = A, b, download
        mdev, cdev = _dev.mesh_dev(mesh), _dev.mpc_dev(mpc)
        dst = [mdev["x"]] + [_dev.function_dev(f) for f in coefficients] + [cdev["coeffs"]]
        for f, t in zip(coefficients, dst[1:]):
            f.device_array = t  # the assembly routines read the coefficient from this tensor
        self.dst = [t for t in dst if t is not None and t.numel() > 0]
        self.src = [t.cpu().pin_memory() for t in self.dst]
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.src)
        nnz, nb = A.nnz, b.data.numel()
        self.nnz, self.nb = nnz, nb
        self.upper = None
        if download == "upper":
            rows = torch.repeat_interleave(torch.arange(A.shape[0], device=A.val.device), A.row_ptr[1:] - A.row_ptr[:-1])
            self.upper = torch.nonzero(rows <= A.col.long()).reshape(-1)
            del rows
        n_out = {"full": nnz, "upper": 0 if self.upper is None else self.upper.numel(), "norms": 1}[download]
        self.d2h_bytes = (n_out + (nb if download != "norms" else 1)) * 8
        # outputs are double-buffered when that stays small next to the device / host memory
        self.nbuf = 2 if nnz * 8 < (8 << 30) else 1
        self._orig = (A.values_storage(), b.data)
        self.val_bufs = [self._orig[0]] + [torch.zeros_like(self._orig[0]) for _ in range(self.nbuf - 1)]
        self.b_bufs = [b.data] + [torch.zeros(nb + (nb & 1), dtype=torch.float64, device=b.data.device)[:nb]
                                  for _ in range(self.nbuf - 1)]  # even capacity (include/mpcx.h)
        self.out_val = [torch.empty(n_out, dtype=torch.float64).pin_memory() for _ in range(self.nbuf)]
        self.out_b = [torch.empty(nb if download != "norms" else 1, dtype=torch.float64).pin_memory()
                      for _ in range(self.nbuf)]
        self.pack = (torch.empty(n_out, dtype=torch.float64, device=A.val.device) if download == "upper" else None)
        self.s_in, self.s_out = torch.cuda.Stream(), torch.cuda.Stream()
        self.download_desc = {"full": "CSR values + RHS",
                              "upper": "upper-triangle CSR values (symmetric system) + RHS",
                              "norms": "Frobenius norm of the matrix and norm of the RHS (system consumed on the device)"}[download]

    def upload(self, after: Optional[torch.cuda.Event] = None) -> torch.cuda.Event:
        """Copy the input values host -> device on the upload stream (after ``after``); returns the event to wait on."""
        ev = torch.cuda.Event()
        with torch.cuda.stream(self.s_in):
            if after is not None:
                self.s_in.wait_event(after)
            for s_, d_ in zip(self.src, self.dst):
                d_.copy_(s_, non_blocking=True)
            ev.record(self.s_in)
        return ev

    def bind(self, i: int):
        """Direct the next assembly into output buffer set ``i``."""
        self.A.bind_values(self.val_bufs[i])
        self.b.bind(self.b_bufs[i])

    def download(self, i: int, after: torch.cuda.Event) -> torch.cuda.Event:
        ev = torch.cuda.Event()
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(after)
            v, bb = self.val_bufs[i][: self.nnz], self.b_bufs[i]
            if self.mode == "full":
                self.out_val[i].copy_(v, non_blocking=True)
                self.out_b[i].copy_(bb, non_blocking=True)
            elif self.mode == "upper":
                torch.index_select(v, 0, self.upper, out=self.pack)
                self.out_val[i].copy_(self.pack, non_blocking=True)
                self.out_b[i].copy_(bb, non_blocking=True)
            else:
                self.out_val[i].copy_(torch.linalg.vector_norm(v).reshape(1), non_blocking=True)
                self.out_b[i].copy_(torch.linalg.vector_norm(bb).reshape(1), non_blocking=True)
            ev.record(self.s_out)
        return ev

    def release(self):
        """Re-attach the matrix / vector to their original storage."""
        self.A.bind_values(self._orig[0])
        self.b.bind(self._orig[1])
