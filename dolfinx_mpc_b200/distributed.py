"""Multi-GPU assembly: cells sharded by ownership, one ghost-row exchange at the end of every assembly.

The reference assembles process-local cells with no communication (``README.md:30``) into rows for owned and
ghost dofs; the cross-rank reduction happens afterwards in its callers -- PETSc ``MatAssemblyBegin/End``
(``python/src/dolfinx_mpc/assemble_matrix.py:64``) and ``VecGhostUpdate(ADD_VALUES, SCATTER_REVERSE)``
(``python/tests/test_vector_assembly.py:51``).  Here every GPU owns a local CSR over its owned + ghost rows
(ghost rows last, so their values are one contiguous segment of ``val``); after the assembly kernels the
ghost-row segments travel to their owners in ONE collective (``all_to_all_single`` over NCCL / NVLink) and are
added in place by a scatter-add kernel.  As in a PETSc MPIAIJ matrix, the owner's rows are extended at setup
with the columns that only off-process cells couple to ("pattern ghosts": extra local column ids after the
index-map ghosts), so the exchange is a pure value transfer through a precomputed position table.

Everything in this module except ``GhostExchange.reduce_*`` is once-per-pattern host code (numpy +
``torch.distributed`` object collectives).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Callable, Optional

import numpy as np
import torch
import torch.distributed as dist

from . import fem
from . import generators as gen
from .fem import IndexMap
from .multipointconstraint import MultiPointConstraint


# ------------------------------------------------------------------ pre-partitioned structured problem

def slab_index_map(n: int, nzc: int, rank: int, world: int) -> IndexMap:
    """Node ownership of the z-slab partition: rank r owns the node planes [r*nzc, (r+1)*nzc) -- the last rank
    also the top plane -- and ghosts the plane (r+1)*nzc owned by rank r+1."""
    plane = n * n
    lo = rank * nzc * plane
    n_own_planes = nzc + (1 if rank == world - 1 else 0)
    size_local = n_own_planes * plane
    total = (world * nzc + 1) * plane
    if rank < world - 1:
        ghosts = np.arange(plane, dtype=np.int64) + (rank + 1) * nzc * plane
        owners = np.full(plane, rank + 1, dtype=np.int32)
    else:
        ghosts = np.zeros(0, np.int64)
        owners = np.zeros(0, np.int32)
    return IndexMap(size_local, ghosts, owners, (lo, lo + size_local), total, rank)


def build_slab_problem(n: int, rank: int, world: int, f_expr: Callable, periodic_z: bool = False,
                       nzc: Optional[int] = None, dirichlet: Optional[bool] = None, bc_value: float = 0.25):
    """P1 Poisson on the box [0,1]^2 x [0, world*nzc/(n-1)] with n x n x (world*nzc + 1) nodes, cut into
    ``world`` z-slabs of ``nzc`` cube layers.  Periodic in x and y; in z either Dirichlet on both end planes
    (BASELINE.json configs[1] stacked) or periodic as well (configs[3]: the slaves of the top plane have
    their masters on rank 0, which become extra ghosts of the last rank)."""
    nzc = (n - 1) if nzc is None else nzc
    dirichlet = (not periodic_z) if dirichlet is None else dirichlet
    h = 1.0 / (n - 1)
    z0 = rank * nzc * h
    mesh = gen.create_box(n - 1, n - 1, nzc, p0=(0.0, 0.0, z0), p1=(1.0, 1.0, z0 + nzc * h))
    mesh.rank, mesh.comm_size = rank, world
    imap = slab_index_map(n, nzc, rank, world)
    # local numbering == lattice numbering of the slab: owned planes first, the ghost (top) plane last
    V = fem.FunctionSpace(mesh, 1, mesh.x_dofmap.copy(), 1, imap, mesh.x.copy())
    X = V.tabulate_dof_coordinates()
    ztop = world * nzc * h
    bcs, exclude = [], None
    if dirichlet:
        on = np.zeros(X.shape[0], dtype=bool)
        if rank == 0:
            on |= np.isclose(X[:, 2], 0.0)
        if rank == world - 1:
            on |= np.isclose(X[:, 2], ztop)
        exclude = np.flatnonzero(on).astype(np.int32)
        bcs = [fem.DirichletBC(V, exclude, bc_value)]
    data = gen.periodic_constraint(V, axes=(0, 1), exclude_dofs=exclude)
    if periodic_z and rank == world - 1:
        # top plane -> plane z = 0 (owner rank 0) composed with the x/y wrap; replaces the x/y entries there
        plane = n * n
        top_local = np.arange(plane, dtype=np.int64) + nzc * plane
        i, j = top_local % n, (top_local // n) % n
        m_glob = (i % (n - 1)) + n * (j % (n - 1))  # k = 0 plane, wrapped in x and y
        keep = ~np.isin(data[0], top_local)
        slaves = np.concatenate([data[0][keep], top_local.astype(np.int32)])
        masters = np.concatenate([data[1][keep], m_glob])
        owners = np.concatenate([data[3][keep], np.zeros(plane, np.int32)])
        coeffs = np.ones(len(slaves))
        if exclude is not None and len(exclude):
            ok = ~np.isin(slaves, exclude)
            slaves, masters, owners, coeffs = slaves[ok], masters[ok], owners[ok], coeffs[ok]
        data = (slaves.astype(np.int32), masters, coeffs, owners.astype(np.int32),
                np.arange(len(slaves) + 1, dtype=np.int32))
    mpc = MultiPointConstraint(V)
    mpc.add_constraint(V, *data)
    mpc.finalize()
    a = fem.laplace(V)
    f = fem.Function(V)
    f.interpolate(f_expr)
    L = fem.source(V, f)
    return dict(mesh=mesh, V=V, bcs=bcs, data=data, mpc=mpc, a=a, L=L, f=f, n=n, nzc=nzc, rank=rank, world=world)


# ------------------------------------------------------------------ pattern extension + exchange plan (setup)

def _all_to_all_objects(objs, group):
    """objs[r] goes to rank r; returns the list received (one entry per source rank).  Each entry is None or a tuple
    of int64 numpy arrays.  Over NCCL the arrays travel point to point as device tensors (two
    ``all_to_all_single`` calls: sizes, then payload) -- every rank moves only its own interface data; without NCCL
    (the gloo tests on the CPU) the exchange falls back to ``all_gather_object``."""
    world = dist.get_world_size(group)
    me = dist.get_rank(group)
    if dist.get_backend(group) != "nccl":
        gathered = [None] * world
        dist.all_gather_object(gathered, objs, group=group)
        return [gathered[src][me] for src in range(world)]
    from . import device as _dev

    dev = _dev.device()
    narr = max([len(o) for o in objs if o is not None], default=0)
    t = torch.tensor([narr], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    narr = int(t.item())
    # sizes[dst, k] = length of array k sent to dst (-1: nothing for that rank)
    sizes = np.full((world, max(1, narr)), -1, dtype=np.int64)
    for r, o in enumerate(objs):
        if o is not None:
            sizes[r, : len(o)] = [len(a) for a in o]
    s_send = torch.from_numpy(sizes).to(dev)
    s_recv = torch.empty_like(s_send)
    dist.all_to_all_single(s_recv, s_send, group=group)
    rsz = s_recv.cpu().numpy()
    payload = [np.concatenate([np.asarray(a, dtype=np.int64).reshape(-1) for a in o]) if o is not None and len(o) else
               np.zeros(0, np.int64) for o in objs]
    send = torch.from_numpy(np.concatenate(payload) if payload else np.zeros(0, np.int64)).to(dev)
    in_splits = [int(np.clip(rsz[r], 0, None).sum()) for r in range(world)]
    out_splits = [len(p_) for p_ in payload]
    recv = torch.empty(sum(in_splits), dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv, send, in_splits, out_splits, group=group)
    flat = recv.cpu().numpy()
    out, pos = [], 0
    for r in range(world):
        if rsz[r, 0] < 0:
            out.append(None)
            continue
        arrs = []
        for k in range(narr):
            n = int(rsz[r, k])
            if n < 0:
                break
            arrs.append(flat[pos:pos + n])
            pos += n
        out.append(tuple(arrs))
    return out


def extend_pattern(row_ptr: np.ndarray, col: np.ndarray, imap_rows: IndexMap, imap_cols: IndexMap, bs0: int, bs1: int,
                   group=None):
    """Merge the ghost-row patterns of all ranks into their owners' rows.

    Returns ``(row_ptr, col, col_global, plan)``: the extended local CSR (ghost rows unchanged, still last),
    the global id of every local column (owned, index-map ghosts, then pattern ghosts) and the exchange plan
    ``{"send_idx": int64[], "send_counts": [...], "recv_pos": int64[], "recv_counts": [...]}`` -- values
    ``val[send_idx]`` ordered by destination rank, and the positions in the owner's ``val`` where the received
    values are added.
    """
    world = dist.get_world_size(group)
    me = dist.get_rank(group)
    n_owned_r = imap_rows.size_local * bs0
    n_rows = len(row_ptr) - 1
    ncols_local = (imap_cols.size_local + imap_cols.num_ghosts) * bs1

    def col_l2g(c):
        c = np.asarray(c, dtype=np.int64)
        return imap_cols.local_to_global(c // bs1) * bs1 + c % bs1

    # ghost rows -> (global row, global cols) per owner
    ghost_rows = np.arange(n_owned_r, n_rows, dtype=np.int64)
    g_owner = np.repeat(imap_rows.owners.astype(np.int64), bs0)
    g_glob = np.repeat(imap_rows.ghosts.astype(np.int64), bs0) * bs0 + np.tile(np.arange(bs0), imap_rows.num_ghosts)
    send, send_idx, send_counts = [], [], []
    for dst in range(world):
        rows = ghost_rows[g_owner == dst]
        if len(rows) == 0:
            send.append(None)
            send_counts.append(0)
            continue
        lens = (row_ptr[rows + 1] - row_ptr[rows]).astype(np.int64)
        if np.all(np.diff(rows) == 1):  # ghosts of one owner are usually one contiguous block of rows
            idx = np.arange(row_ptr[rows[0]], row_ptr[rows[-1] + 1])
        else:
            idx = np.concatenate([np.arange(row_ptr[r], row_ptr[r + 1]) for r in rows])
        send.append((g_glob[rows - n_owned_r], lens, col_l2g(col[idx])))
        send_idx.append(idx.astype(np.int64))
        send_counts.append(int(len(idx)))
    recv = _all_to_all_objects(send, group)

    # owner side: add the received (row, col) pairs to the owned rows
    lo_r = imap_rows.local_range[0] * bs0
    col_global = col_l2g(np.arange(ncols_local))
    extra_r, extra_c = [], []
    for src in range(world):
        if recv[src] is None:
            continue
        grow, lens, gcols = recv[src]
        extra_r.append(np.repeat(grow - lo_r, lens))
        extra_c.append(gcols)
    recv_pos, recv_counts = np.zeros(0, np.int64), [0] * world
    if extra_r:
        er = np.concatenate(extra_r)
        ec_g = np.concatenate(extra_c)
        assert er.min() >= 0 and er.max() < n_owned_r, "received a ghost row this rank does not own"
        # global -> local column, allocating pattern ghosts for unknown columns
        blk = ec_g // bs1
        loc = imap_cols.global_to_local(blk)
        unknown = loc < 0
        new_blocks = np.unique(blk[unknown])
        if len(new_blocks):
            loc[unknown] = np.searchsorted(new_blocks, blk[unknown]) + imap_cols.size_local + imap_cols.num_ghosts
            newg = (new_blocks[:, None] * bs1 + np.arange(bs1)[None, :]).reshape(-1)
            col_global = np.concatenate([col_global, newg])
        ec = loc * bs1 + ec_g % bs1
        # merge, touching only the owned rows that receive something: union of their old entries and the
        # received ones, sorted by column; every other row is copied as a block
        ncol_tot = len(col_global)
        rows_a = np.unique(er)
        old_len = np.diff(row_ptr)
        old_idx = np.concatenate([np.arange(row_ptr[r], row_ptr[r + 1]) for r in rows_a])
        old_keys = np.repeat(rows_a, old_len[rows_a]) * ncol_tot + col[old_idx].astype(np.int64)
        keys = np.unique(np.concatenate([old_keys, er * ncol_tot + ec]))  # sorted by (row, col)
        k_rows = keys // ncol_tot
        cstart = np.searchsorted(k_rows, rows_a)  # start of every affected row inside `keys`
        new_len = old_len.copy()
        new_len[rows_a] = np.diff(np.append(cstart, len(keys)))
        new_row_ptr = np.zeros(n_rows + 1, dtype=np.int64)
        np.cumsum(new_len, out=new_row_ptr[1:])
        new_col = np.empty(int(new_row_ptr[-1]), dtype=np.int32)
        prev = 0  # copy the untouched row blocks between affected rows
        for j, r in enumerate(rows_a):
            if r > prev:
                new_col[new_row_ptr[prev]:new_row_ptr[r]] = col[row_ptr[prev]:row_ptr[r]]
            e = cstart[j + 1] if j + 1 < len(rows_a) else len(keys)
            new_col[new_row_ptr[r]:new_row_ptr[r + 1]] = (keys[cstart[j]:e] % ncol_tot).astype(np.int32)
            prev = r + 1
        new_col[new_row_ptr[prev]:] = col[row_ptr[prev]:]
        # positions of the received entries, per source rank in arrival order
        kpos = np.searchsorted(keys, er * ncol_tot + ec)
        ridx = np.searchsorted(rows_a, er)
        recv_pos = (new_row_ptr[er] + (kpos - cstart[ridx])).astype(np.int64)
        recv_counts = [0 if recv[s] is None else int(recv[s][1].sum()) for s in range(world)]
        # send indices refer to ghost rows, which moved: shift by the growth of the owned part
        shift = new_row_ptr[n_owned_r] - row_ptr[n_owned_r]
        send_idx = [i + shift for i in send_idx]
        row_ptr, col = new_row_ptr, new_col
    plan = {"send_idx": np.concatenate(send_idx) if send_idx else np.zeros(0, np.int64),
            "send_counts": send_counts, "recv_pos": recv_pos, "recv_counts": recv_counts}
    return row_ptr, col, col_global, plan


def _all_to_all_tensors(objs, group, dev):
    """Tensor version of :func:`_all_to_all_objects`: ``objs[r]`` is None or a tuple of int64 tensors on ``dev``; the
    received tuples are views of one receive buffer on ``dev``.  NCCL: two ``all_to_all_single`` calls (sizes, payload),
    the payload never leaves the device; gloo (CPU tests): ``all_gather_object``."""
    world = dist.get_world_size(group)
    me = dist.get_rank(group)
    if dist.get_backend(group) != "nccl":
        gathered = [None] * world
        dist.all_gather_object(gathered, [None if o is None else tuple(t.cpu() for t in o) for o in objs], group=group)
        return [None if gathered[src][me] is None else tuple(t.to(dev) for t in gathered[src][me]) for src in range(world)]
    narr = max([len(o) for o in objs if o is not None], default=0)
    t = torch.tensor([narr], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    narr = int(t.item())
    sizes = np.full((world, max(1, narr)), -1, dtype=np.int64)
    for r, o in enumerate(objs):
        if o is not None:
            sizes[r, : len(o)] = [a.numel() for a in o]
    s_send = torch.from_numpy(sizes).to(dev)
    s_recv = torch.empty_like(s_send)
    dist.all_to_all_single(s_recv, s_send, group=group)
    rsz = s_recv.cpu().numpy()
    parts = [a.reshape(-1).to(torch.int64) for o in objs if o is not None for a in o]
    send = torch.cat(parts) if parts else torch.zeros(0, dtype=torch.int64, device=dev)
    out_splits = [0 if o is None else int(sum(a.numel() for a in o)) for o in objs]
    in_splits = [int(np.clip(rsz[r], 0, None).sum()) for r in range(world)]
    recv = torch.empty(sum(in_splits), dtype=torch.int64, device=dev)
    dist.all_to_all_single(recv, send, in_splits, out_splits, group=group)
    out, pos = [], 0
    for r in range(world):
        if rsz[r, 0] < 0:
            out.append(None)
            continue
        arrs = []
        for k in range(narr):
            n = int(rsz[r, k])
            if n < 0:
                break
            arrs.append(recv[pos:pos + n])
            pos += n
        out.append(tuple(arrs))
    return out


def extend_pattern_device(row_ptr: torch.Tensor, col: torch.Tensor, imap_rows: IndexMap, imap_cols: IndexMap, bs0: int,
                          bs1: int, group=None):
    """:func:`extend_pattern` on tensors that stay where they are (the device under NCCL): the ghost rows' (global row,
    global column) pairs travel point to point, the owner maps the columns to local ones (new pattern ghosts for
    unknown blocks), and the owned rows between the first and the last row that receive something are rebuilt in ONE
    sort of (row, column) keys -- no per-row host loop, no host copy of the pattern.  Same return values as
    :func:`extend_pattern`, as tensors (``col_global`` too)."""
    dev = row_ptr.device
    world = dist.get_world_size(group)
    i64 = torch.int64
    n_owned_r = imap_rows.size_local * bs0
    n_rows = row_ptr.numel() - 1
    nloc_blocks = imap_cols.size_local + imap_cols.num_ghosts
    ncols_local = nloc_blocks * bs1
    lo_c, hi_c = imap_cols.local_range
    ghosts_c = torch.from_numpy(np.asarray(imap_cols.ghosts, dtype=np.int64)).to(dev)

    def col_l2g(c):
        blk = torch.div(c, bs1, rounding_mode="floor")
        g = blk + lo_c
        if ghosts_c.numel():
            gi = (blk - imap_cols.size_local).clamp_(min=0)
            g = torch.where(blk < imap_cols.size_local, g, ghosts_c[gi])
        return g * bs1 + c % bs1

    # ghost rows -> (global row, row lengths, global cols) per owner; the per-row bookkeeping (O(ghost rows)) on the host
    ghost_rows = np.arange(n_owned_r, n_rows, dtype=np.int64)
    g_owner = np.repeat(np.asarray(imap_rows.owners, dtype=np.int64), bs0)
    g_glob = np.repeat(np.asarray(imap_rows.ghosts, dtype=np.int64), bs0) * bs0 + np.tile(np.arange(bs0), imap_rows.num_ghosts)
    send, send_idx, send_counts = [], [], []
    for dst in range(world):
        rows = ghost_rows[g_owner == dst]
        if len(rows) == 0:
            send.append(None)
            send_counts.append(0)
            continue
        rows_t = torch.from_numpy(rows).to(dev)
        starts = row_ptr[rows_t]
        lens = row_ptr[rows_t + 1] - starts
        if np.all(np.diff(rows) == 1):  # ghosts of one owner are usually one contiguous block of rows
            idx = torch.arange(int(starts[0]), int(row_ptr[int(rows[-1]) + 1]), device=dev, dtype=i64)
        else:
            excl = torch.cumsum(lens, 0) - lens
            idx = torch.arange(int(lens.sum()), device=dev, dtype=i64) + torch.repeat_interleave(starts - excl, lens)
        send.append((torch.from_numpy(g_glob[rows - n_owned_r]).to(dev), lens, col_l2g(col[idx].to(i64))))
        send_idx.append(idx)
        send_counts.append(int(idx.numel()))
    recv = _all_to_all_tensors(send, group, dev)

    lo_r = imap_rows.local_range[0] * bs0
    col_global = col_l2g(torch.arange(ncols_local, device=dev, dtype=i64))
    extra_r, extra_c = [], []
    for src in range(world):
        if recv[src] is None:
            continue
        grow, lens, gcols = recv[src]
        extra_r.append(torch.repeat_interleave(grow - lo_r, lens))
        extra_c.append(gcols)
    recv_pos, recv_counts = torch.zeros(0, dtype=i64, device=dev), [0] * world
    if extra_r:
        er = torch.cat(extra_r)
        ec_g = torch.cat(extra_c)
        if er.numel():
            assert int(er.min()) >= 0 and int(er.max()) < n_owned_r, "received a ghost row this rank does not own"
        # global -> local column, allocating pattern ghosts for unknown blocks
        blk = torch.div(ec_g, bs1, rounding_mode="floor")
        own = (blk >= lo_c) & (blk < hi_c)
        loc = torch.where(own, blk - lo_c, torch.full_like(blk, -1))
        if ghosts_c.numel():
            sg, order = torch.sort(ghosts_c, stable=True)
            pos = torch.searchsorted(sg, blk).clamp_(max=sg.numel() - 1)
            hit = (sg[pos] == blk) & ~own
            loc = torch.where(hit, order[pos] + imap_cols.size_local, loc)
        unknown = loc < 0
        new_blocks = torch.unique(blk[unknown])
        if new_blocks.numel():
            loc = torch.where(unknown, torch.searchsorted(new_blocks, blk) + nloc_blocks, loc)
            newg = (new_blocks[:, None] * bs1 + torch.arange(bs1, device=dev, dtype=i64)[None, :]).reshape(-1)
            col_global = torch.cat([col_global, newg])
        ec = loc * bs1 + ec_g % bs1
        ncol_tot = int(col_global.numel())
        new_keys = er * ncol_tot + ec
        # rebuild the owned rows [r_lo, r_hi): old entries and received ones, unique, sorted by (row, column)
        r_lo, r_hi = int(er.min()), int(er.max()) + 1
        e_lo, e_hi = int(row_ptr[r_lo]), int(row_ptr[r_hi])
        old_len = row_ptr[1:] - row_ptr[:-1]
        row_of = torch.repeat_interleave(torch.arange(r_lo, r_hi, device=dev, dtype=i64), old_len[r_lo:r_hi])
        keys = torch.unique(torch.cat([row_of * ncol_tot + col[e_lo:e_hi].to(i64), new_keys]))  # sorted
        del row_of
        k_rows = torch.div(keys, ncol_tot, rounding_mode="floor")
        new_len = old_len.clone()
        new_len[r_lo:r_hi] = torch.bincount(k_rows - r_lo, minlength=r_hi - r_lo)
        new_row_ptr = torch.zeros(n_rows + 1, dtype=i64, device=dev)
        torch.cumsum(new_len, 0, out=new_row_ptr[1:])
        new_col = torch.cat([col[:e_lo], (keys - k_rows * ncol_tot).to(torch.int32), col[e_hi:]])
        recv_pos = e_lo + torch.searchsorted(keys, new_keys)
        recv_counts = [0 if recv[s_] is None else int(recv[s_][1].sum()) for s_ in range(world)]
        shift = int(new_row_ptr[n_owned_r] - row_ptr[n_owned_r])  # ghost rows moved by the growth of the owned part
        send_idx = [i + shift for i in send_idx]
        row_ptr, col = new_row_ptr, new_col
    plan = {"send_idx": torch.cat(send_idx) if send_idx else torch.zeros(0, dtype=i64, device=dev),
            "send_counts": send_counts, "recv_pos": recv_pos, "recv_counts": recv_counts}
    return row_ptr, col, col_global, plan


def vector_plan(imap: IndexMap, bs: int, group=None):
    """Exchange plan of ``VecGhostUpdate(ADD, REVERSE)``: ghost entries grouped by owner, and on the owner
    the local positions they are added to."""
    world = dist.get_world_size(group)
    n_owned = imap.size_local * bs
    g_owner = np.repeat(imap.owners.astype(np.int64), bs)
    g_glob = np.repeat(imap.ghosts.astype(np.int64), bs) * bs + np.tile(np.arange(bs), imap.num_ghosts)
    send, send_idx, send_counts = [], [], []
    for dst in range(world):
        sel = np.flatnonzero(g_owner == dst)
        send.append((g_glob[sel],) if len(sel) else None)
        send_idx.append(sel + n_owned)
        send_counts.append(len(sel))
    recv = _all_to_all_objects(send, group)
    lo = imap.local_range[0] * bs
    pos = [r[0] - lo for r in recv if r is not None]
    return {"send_idx": np.concatenate(send_idx).astype(np.int64), "send_counts": send_counts,
            "recv_pos": np.concatenate(pos).astype(np.int64) if pos else np.zeros(0, np.int64),
            "recv_counts": [0 if r is None else len(r[0]) for r in recv]}


# ------------------------------------------------------------------ per-step exchange (device)

_LIB_COMM = {}


def lib_comm(group=None):
    """The NCCL communicator of libmpcx for ``group`` (``mpcx_comm_create``): the unique id is made on rank 0 and
    broadcast through ``torch.distributed``; NCCL itself is the library PyTorch already loaded.  None when the
    process group is not NCCL (the gloo tests on the CPU)."""
    if dist.get_backend(group) != "nccl":
        return None
    key = id(group)
    if key not in _LIB_COMM:
        from . import _lib

        lib = _lib.load()
        path = None
        try:  # the NCCL build PyTorch ships with (same ABI as the one it talks to)
            import glob
            import os

            import nvidia.nccl

            hits = glob.glob(os.path.join(os.path.dirname(nvidia.nccl.__file__), "lib", "libnccl.so*"))
            path = hits[0] if hits else None
        except Exception:
            path = None
        _lib.check(lib.mpcx_nccl_load(path.encode() if path else None))
        uid = (C.c_char * 128)()
        if dist.get_rank(group) == 0:
            _lib.check(lib.mpcx_comm_unique_id(uid))
        box = [bytes(uid)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        uid = (C.c_char * 128).from_buffer_copy(box[0])
        comm = C.c_void_p()
        _lib.check(lib.mpcx_comm_create(uid, dist.get_rank(group), dist.get_world_size(group), C.byref(comm)))
        _LIB_COMM[key] = comm
    return _LIB_COMM[key]


class GhostExchange:
    """Ghost values to their owners and added there, in ONE library call over NCCL (``mpcx_ghost_reduce_f64``: pack
    kernel, grouped ncclSend / ncclRecv, scatter-add kernel on the caller's stream).  Without an NCCL process group
    (gloo tests on the CPU) the same plan runs through ``all_to_all_single`` and two index primitives.

    ``send_idx`` / ``recv_pos`` live on the device.  When the ghost values to send are one contiguous slice of
    the source array (single owner, the slab case) the pack kernel is skipped and the slice is sent in place.
    """

    def __init__(self, plan: dict, device, group=None):
        self.group = group
        self.send_counts = list(plan["send_counts"])
        self.recv_counts = list(plan["recv_counts"])
        si, rp = plan["send_idx"], plan["recv_pos"]
        si = si.to(torch.int64) if isinstance(si, torch.Tensor) else torch.from_numpy(np.asarray(si, dtype=np.int64))
        rp = rp.to(torch.int64) if isinstance(rp, torch.Tensor) else torch.from_numpy(np.asarray(rp, dtype=np.int64))
        self.n_send, self.n_recv = int(si.numel()), int(rp.numel())
        self.contiguous = self.n_send > 0 and bool(torch.all(si[1:] - si[:-1] == 1))
        self.send_start = int(si[0]) if self.n_send else 0
        self.send_idx = si.to(device)
        self.recv_pos = rp.to(device)
        self.send_buf = torch.empty(self.n_send, dtype=torch.float64, device=device)
        self.recv_buf = torch.empty(self.n_recv, dtype=torch.float64, device=device)
        self.comm = None
        if torch.device(device).type == "cuda" and dist.is_initialized():
            self.comm = lib_comm(group)
            world = dist.get_world_size(group)
            self._sc = (C.c_int64 * world)(*self.send_counts)
            self._rc = (C.c_int64 * world)(*self.recv_counts)

    # the two device primitives; tests on CPU substitute torch index ops for them
    def _gather(self, src: torch.Tensor, idx: torch.Tensor, out: torch.Tensor):
        from . import _lib, device as _dev

        _lib.check(_lib.load().mpcx_gather_f64(src.data_ptr(), idx.data_ptr(), idx.numel(), out.data_ptr(),
                                               _dev.stream_ptr()))

    def _scatter_add(self, dst: torch.Tensor, idx: torch.Tensor, vals: torch.Tensor):
        from . import _lib, device as _dev

        _lib.check(_lib.load().mpcx_scatter_add_f64(dst.data_ptr(), idx.data_ptr(), idx.numel(), vals.data_ptr(),
                                                    _dev.stream_ptr()))

    def reduce(self, values: torch.Tensor):
        if self.comm is not None:
            from . import _lib, device as _dev

            _lib.check(_lib.load().mpcx_ghost_reduce_f64(
                self.comm, values.data_ptr(), None if self.contiguous else self.send_idx.data_ptr(),
                self.send_start if self.contiguous else 0, self._sc, self.recv_pos.data_ptr() if self.n_recv else None,
                self._rc, self.send_buf.data_ptr() if self.n_send else None,
                self.recv_buf.data_ptr() if self.n_recv else None, _dev.stream_ptr()))
            return
        if self.contiguous:
            send = values[self.send_start:self.send_start + self.n_send]
        else:
            send = self.send_buf
            if self.n_send:
                self._gather(values, self.send_idx, send)
        dist.all_to_all_single(self.recv_buf, send, self.recv_counts, self.send_counts, group=self.group)
        if self.n_recv:
            self._scatter_add(values, self.recv_pos, self.recv_buf)


class MatVecExchange:
    """What ``la.Matrix.assemble`` / ``la.Vector.ghostUpdate`` call."""

    def __init__(self, mat_plan, vec_plan, device, group=None, cls=GhostExchange):
        self.mat = cls(mat_plan, device, group) if mat_plan is not None else None
        self.vec = cls(vec_plan, device, group) if vec_plan is not None else None

    def reduce_matrix(self, A):
        self.mat.reduce(A.val)

    def reduce_vector(self, b):
        self.vec.reduce(b.data)


def create_matrix(a: fem.Form, mpc: MultiPointConstraint, group=None):
    """Distributed counterpart of ``create_matrix``: local pattern, extended with the off-process couplings of
    the owned rows, plus the attached ghost-row exchange."""
    from . import device as _dev
    from .assemble_matrix import create_sparsity_pattern
    from .la import Matrix

    V = mpc.function_space
    if torch.cuda.is_available() and os.environ.get("MPCX_PATTERN", "device") != "host":
        # pattern built, extended and kept on the device; col_global comes back for the callers
        from .assemble_matrix import create_sparsity_pattern_device

        row_ptr, col = create_sparsity_pattern_device(a, mpc)
        row_ptr, col, col_global, plan = extend_pattern_device(row_ptr, col, V.index_map, V.index_map, V.bs, V.bs, group)
        A = Matrix(row_ptr, col.contiguous(), (V.num_dofs, int(col_global.numel())), (V.bs, V.bs))
        A.col_global = col_global.cpu().numpy()
    else:
        row_ptr, col = create_sparsity_pattern(a, mpc)
        row_ptr, col, col_global, plan = extend_pattern(row_ptr, col, V.index_map, V.index_map, V.bs, V.bs, group)
        A = Matrix(row_ptr, col, (V.num_dofs, len(col_global)), (V.bs, V.bs))
        A.col_global = col_global
    A.ghost_exchange = MatVecExchange(plan, None, _dev.device(), group)
    return A


def attach_ghost_exchange(A, b, P, group=None):
    """bench.py helper: ``A`` must come from :func:`create_matrix`; attaches the vector exchange to ``b``."""
    from . import device as _dev

    V = P["mpc"].function_space
    b.ghost_exchange = MatVecExchange(None, vector_plan(V.index_map, V.bs, group), _dev.device(), group)
