"""``MultiPointConstraint`` -- same surface as the reference class
(``python/src/dolfinx_mpc/multipointconstraint.py:87-631``), array-backed.

``finalize`` produces exactly the packed data the assembly kernels read
(``cpp/MultiPointConstraint.h:36-126``, ``cpp/mpc_helpers.h:19-94,139-235``):
``is_slave`` (int8 per local dof), dof-indexed master / coefficient / owner
adjacency lists, the ascending slave list with the owned ones first,
``cell_to_slaves`` over owned cells, and the function space whose index map is
extended with non-local masters as ghosts.  It is a once-per-constraint host
computation (numpy); its outputs are mirrored to the device on first use.
"""
from __future__ import annotations

from typing import Optional, Union

import numpy as np

from .fem import Function, FunctionSpace, IndexMap


class AdjacencyList:
    """``dolfinx.graph.AdjacencyList`` look-alike (``array`` + ``offsets``)."""

    def __init__(self, array: np.ndarray, offsets: np.ndarray):
        self.array = array
        self.offsets = offsets

    @property
    def num_nodes(self) -> int:
        return len(self.offsets) - 1

    def links(self, i: int) -> np.ndarray:
        return self.array[self.offsets[i]:self.offsets[i + 1]]


class MultiPointConstraint:
    """Hold data for multi point constraint relationships (see module docstring).

    Args:
        V: the function space
        dtype: scalar type of the coefficients (float64 is the only type with device kernels)
    """

    def __init__(self, V: FunctionSpace, dtype=np.float64):
        if np.dtype(dtype) != np.float64:
            raise ValueError(f"Unsupported dtype {np.dtype(dtype)} for coefficients (float64 only)")
        self._slaves = np.array([], dtype=np.int32)
        self._masters = np.array([], dtype=np.int64)
        self._coeffs = np.array([], dtype=dtype)
        self._owners = np.array([], dtype=np.int32)
        self._offsets = np.array([0], dtype=np.int32)
        self.V = V
        self.finalized = False
        self._dtype = np.dtype(dtype)
        self._dev = {}

    # -- construction (multipointconstraint.py:118-167) ---------------------------------------------------------
    def add_constraint(self, V: FunctionSpace, slaves, masters, coeffs, owners, offsets):
        """Add a constraint given by arrays: ``slaves`` local int32, ``masters`` global int64, ``coeffs``,
        ``owners`` int32, ``offsets`` int32 with ``masters_of_slave[i] = masters[offsets[i]:offsets[i+1]]``."""
        assert V is self.V
        self._already_finalized()
        slaves = np.asarray(slaves, dtype=np.int32)
        if len(slaves) > 0:
            offsets = np.asarray(offsets, dtype=np.int32)
            self._offsets = np.append(self._offsets, offsets[1:] + len(self._masters)).astype(np.int32)
            self._slaves = np.append(self._slaves, slaves).astype(np.int32)
            self._masters = np.append(self._masters, np.asarray(masters, dtype=np.int64))
            self._coeffs = np.append(self._coeffs, np.asarray(coeffs, dtype=self._dtype))
            self._owners = np.append(self._owners, np.asarray(owners, dtype=np.int32))

    def add_constraint_from_mpc_data(self, V: FunctionSpace, mpc_data):
        self._already_finalized()
        self.add_constraint(V, *mpc_data)

    def finalize(self) -> None:
        """Build the packed constraint data; no constraints can be added afterwards."""
        self._already_finalized()
        V = self.V
        bs = V.bs
        imap = V.index_map
        slaves, offsets = self._slaves, self._offsets
        n_in = len(slaves)
        counts = np.diff(offsets)

        # extended index map: masters not present locally become ghosts (mpc_helpers.h:139-235)
        mblocks = self._masters // bs
        loc = imap.global_to_local(mblocks)
        missing = loc < 0
        if missing.any():
            new_g, first = np.unique(mblocks[missing], return_index=True)
            new_o = self._owners[missing][first]
            imap = IndexMap(imap.size_local, np.concatenate([imap.ghosts, new_g]),
                            np.concatenate([imap.owners, new_o]).astype(np.int32), imap.local_range,
                            imap.size_global, imap.rank)
            coords = None
            if V.dof_coordinates is not None:  # coordinates of new ghosts are unknown locally
                coords = np.concatenate([V.dof_coordinates, np.full((len(new_g), 3), np.nan)])
            self.V = V.with_index_map(imap, coords)
            loc = imap.global_to_local(mblocks)
        masters_local = (loc * bs + self._masters % bs).astype(np.int32)
        num_dofs = bs * (imap.size_local + imap.num_ghosts)

        # is_slave marker and dof-indexed adjacency lists (MultiPointConstraint.h:55-100)
        is_slave = np.zeros(num_dofs, dtype=np.int8)
        is_slave[slaves] = 1
        num_masters = np.zeros(num_dofs, dtype=np.int64)
        num_masters[slaves] = counts
        m_offsets = np.zeros(num_dofs + 1, dtype=np.int64)
        np.cumsum(num_masters, out=m_offsets[1:])
        assert m_offsets[-1] < 2**31
        # destination of every input master entry: offsets[slave] + position within the slave
        within = np.arange(len(masters_local), dtype=np.int64) - np.repeat(offsets[:-1].astype(np.int64), counts)
        dest = np.repeat(m_offsets[slaves], counts) + within
        m_masters = np.zeros(len(masters_local), dtype=np.int32)
        m_coeffs = np.zeros(len(masters_local), dtype=self._dtype)
        m_owners = np.zeros(len(masters_local), dtype=np.int32)
        m_masters[dest] = masters_local
        m_coeffs[dest] = self._coeffs
        m_owners[dest] = self._owners
        m_offsets = m_offsets.astype(np.int32)

        # sorted slaves, owned first (MultiPointConstraint.h:105-115)
        sorted_slaves = np.flatnonzero(is_slave).astype(np.int32)
        num_local = int(np.searchsorted(sorted_slaves, bs * imap.size_local))

        # cell -> slaves over owned cells, ascending within a cell (mpc_helpers.h:19-94)
        nc = V.mesh.num_cells_local
        c2s_off = np.zeros(nc + 1, dtype=np.int32)
        c2s = np.zeros(0, dtype=np.int32)
        if n_in > 0:
            dm = V.dofmap[:nc]
            blk_has_slave = np.zeros(imap.size_local + imap.num_ghosts, dtype=bool)
            blk_has_slave[slaves // bs] = True
            cand_cells = np.flatnonzero(blk_has_slave[dm].any(axis=1))
            if len(cand_cells):
                un = (dm[cand_cells].astype(np.int64)[:, :, None] * bs + np.arange(bs)[None, None, :]).reshape(
                    len(cand_cells), -1)
                mask = is_slave[un].astype(bool)
                rows, _ = np.nonzero(mask)
                vals = un[mask]
                order = np.lexsort((vals, rows))
                cells_of = cand_cells[rows[order]]
                c2s = vals[order].astype(np.int32)
                np.cumsum(np.bincount(cells_of, minlength=nc), out=c2s_off[1:])

        self._is_slave = is_slave
        self._sorted_slaves = sorted_slaves
        self._num_local_slaves = num_local
        self._master_map = AdjacencyList(m_masters, m_offsets)
        self._coeff_map = AdjacencyList(m_coeffs, m_offsets)
        self._owner_map = AdjacencyList(m_owners, m_offsets)
        self._cell_to_slaves = AdjacencyList(c2s, c2s_off)
        self._slave_cells = np.flatnonzero(np.diff(c2s_off) > 0).astype(np.int32)
        self.finalized = True
        del (self._slaves, self._masters, self._coeffs, self._owners, self._offsets)

    def create_general_constraint(self, slave_master_dict, subspace_slave: Optional[int] = None,
                                  subspace_master: Optional[int] = None):
        """Point-dictionary constraint (``multipointconstraint.py:401-433``, serial part)."""
        from .generators import general_constraint

        self.add_constraint(self.V, *general_constraint(self.V, slave_master_dict, subspace_slave or 0,
                                                        subspace_master or 0))

    # -- accessors (multipointconstraint.py:503-584) ------------------------------------------------------------
    @property
    def is_slave(self) -> np.ndarray:
        self._not_finalized()
        return self._is_slave

    @property
    def slaves(self) -> np.ndarray:
        self._not_finalized()
        return self._sorted_slaves

    @property
    def masters(self) -> AdjacencyList:
        self._not_finalized()
        return self._master_map

    def coefficients(self):
        self._not_finalized()
        return self._coeff_map.array, self._coeff_map.offsets

    @property
    def owners(self) -> AdjacencyList:
        self._not_finalized()
        return self._owner_map

    @property
    def num_local_slaves(self) -> int:
        self._not_finalized()
        return self._num_local_slaves

    @property
    def cell_to_slaves(self) -> AdjacencyList:
        self._not_finalized()
        return self._cell_to_slaves

    @property
    def slave_cells(self) -> np.ndarray:
        """Owned cells holding at least one slave dof (``extract_slave_cells``, ``numba/helpers.py:24-34``)."""
        self._not_finalized()
        return self._slave_cells

    @property
    def function_space(self) -> FunctionSpace:
        self._not_finalized()
        return self.V

    # -- post-solve helpers (multipointconstraint.py:586-617; cpp/MultiPointConstraint.h:129-152) ----------------
    def backsubstitution(self, u: Union[Function, "object"]) -> None:
        """``u[slave] = sum_k coeff_k * u[master_k]`` on the device; ``u`` is a Function (host array is round
        tripped) or a device Vector."""
        from . import device as _dev

        _dev.backsubstitution(self, u, homogenize=False)

    def homogenize(self, u) -> None:
        from . import device as _dev

        _dev.backsubstitution(self, u, homogenize=True)

    def _already_finalized(self):
        if self.finalized:
            raise RuntimeError("MultiPointConstraint has already been finalized")

    def _not_finalized(self):
        if not self.finalized:
            raise RuntimeError("MultiPointConstraint has not been finalized")
