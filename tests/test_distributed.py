"""World-size-2 and -3 tests (gloo, CPU) of the multi-GPU host logic: slab partition, extended index map with
non-local masters as ghosts, pattern extension on the owner, and the ghost-row / ghost-entry exchange plan.

The per-rank local assembly is done by the CPU oracle (the CUDA kernels need a GPU); the exchange runs through
``GhostExchange`` with its two device primitives replaced by torch CPU index ops.  The owned rows of all ranks
together must reproduce the serial assembly of the global problem (what PETSc MatAssembly / VecGhostUpdate give
the reference, python/src/dolfinx_mpc/assemble_matrix.py:64, python/tests/test_vector_assembly.py:51).
"""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _f(x):
    return x[0] * np.sin(5 * np.pi * x[1]) + x[2] ** 2 + 0.3


def _worker(rank, world, port, periodic_z, n, nzc, out_dir, cuda=False):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if cuda:
        os.environ["LOCAL_RANK"] = str(rank)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        _worker_cuda(rank, world, periodic_z, n, nzc, out_dir)
        dist.barrier()
        dist.destroy_process_group()
        return
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dolfinx_mpc_b200 import create_sparsity_pattern, distributed as D
    from oracle import oracle as orc

    class CpuExchange(D.GhostExchange):
        def _gather(self, src, idx, out):
            out.copy_(src[idx])

        def _scatter_add(self, dst, idx, vals):
            dst.index_add_(0, idx, vals)

    P = D.build_slab_problem(n, rank, world, _f, periodic_z=periodic_z, nzc=nzc)
    mpc, V = P["mpc"], P["mpc"].function_space
    rp0, col0 = create_sparsity_pattern(P["a"], mpc, num_threads=2)
    rp, col, col_global, plan = D.extend_pattern(rp0, col0, V.index_map, V.index_map, 1, 1)
    # the tensor version (what runs on the device under NCCL) gives the same pattern and the same exchange plan
    rp_t, col_t, cg_t, plan_t = D.extend_pattern_device(torch.from_numpy(rp0), torch.from_numpy(col0), V.index_map,
                                                        V.index_map, 1, 1)
    assert np.array_equal(rp_t.numpy(), rp) and np.array_equal(col_t.numpy(), col)
    assert np.array_equal(cg_t.numpy(), col_global)
    assert np.array_equal(plan_t["send_idx"].numpy(), plan["send_idx"]) and np.array_equal(plan_t["recv_pos"].numpy(), plan["recv_pos"])
    assert list(plan_t["send_counts"]) == list(plan["send_counts"]) and list(plan_t["recv_counts"]) == list(plan["recv_counts"])
    m = orc.WrappedMPC(V, mpc.is_slave, mpc.masters.array, mpc.coefficients()[0], mpc.masters.offsets,
                       mpc.cell_to_slaves.array, mpc.cell_to_slaves.offsets, mpc.slaves, mpc.num_local_slaves)
    _, _, val = orc.assemble_matrix(P["a"], m, bcs=P["bcs"], pattern=(rp, col))
    b = orc.assemble_vector(P["L"], m)
    if P["bcs"]:
        orc.apply_lifting(b, [P["a"]], [P["bcs"]], m)
    ex = D.MatVecExchange(plan, D.vector_plan(V.index_map, 1), torch.device("cpu"), cls=CpuExchange)
    tv, tb = torch.from_numpy(val), torch.from_numpy(b)
    ex.mat.reduce(tv)
    ex.vec.reduce(tb)
    n_owned = V.index_map.size_local
    rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    own = rows < n_owned
    lo = V.index_map.local_range[0]
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), r=rows[own] + lo, c=col_global[col[own]], v=val[own],
             b=b[:n_owned], lo=lo, slaves=V.index_map.local_to_global(mpc.slaves[:mpc.num_local_slaves]),
             nghost=V.index_map.num_ghosts, ncol=len(col_global))
    dist.barrier()
    dist.destroy_process_group()


def _worker_cuda(rank, world, periodic_z, n, nzc, out_dir):
    """Same slab problem assembled by the CUDA kernels, ghost rows / entries reduced over NCCL."""
    import dolfinx_mpc_b200 as mpcx
    from dolfinx_mpc_b200 import distributed as D

    P = D.build_slab_problem(n, rank, world, _f, periodic_z=periodic_z, nzc=nzc)
    mpc, V = P["mpc"], P["mpc"].function_space
    A = D.create_matrix(P["a"], mpc)
    b = mpcx.create_vector(mpc)
    D.attach_ghost_exchange(A, b, P)
    if os.environ.get("MPCX_TEST_FUSED") == "1":
        # one call: fused matrix + vector kernel in two parts, the ghost rows travelling while the interior is assembled
        for _ in range(2):  # twice into the same objects: cached plans, streams and events reused
            mpcx.assemble_system(P["a"], P["L"], mpc, bcs=P["bcs"], A=A, b=b)
        assert A.last_system_fused
        info = [e[1] for e in A._tile_plans.values() if e is not None and "interface_tiles" in e[1]]
        # rank 0's top plane is owned by rank 1: its cells below that plane are the interface tiles (the last rank
        # reaches its ghosts -- masters on rank 0 -- only through cells holding slaves: no interface tile there)
        assert info and 0 <= info[0]["interface_tiles"] < info[0]["tiles"], info
        if rank == 0:
            assert info[0]["interface_tiles"] > 0, info
    else:
        mpcx.assemble_matrix(P["a"], mpc, bcs=P["bcs"], A=A)
        mpcx.assemble_vector(P["L"], mpc, b=b)
        if P["bcs"]:
            mpcx.apply_lifting(b, [P["a"]], [P["bcs"]], mpc)
        b.ghostUpdate()
    torch.cuda.synchronize()
    rp, col, val = A.getValuesCSR()
    n_owned = V.index_map.size_local
    rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    own = rows < n_owned
    lo = V.index_map.local_range[0]
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), r=rows[own] + lo, c=A.col_global[col[own]], v=val[own],
             b=b.array[:n_owned], lo=lo, slaves=V.index_map.local_to_global(mpc.slaves[:mpc.num_local_slaves]),
             nghost=V.index_map.num_ghosts, ncol=len(A.col_global))


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["routines", "system", "system-overlap"])
@pytest.mark.parametrize("periodic_z", [False, True])
def test_two_gpu_assembly_matches_serial(oracle, tmp_path, periodic_z, mode):
    """The distributed path end to end on two real GPUs (CUDA kernels + NCCL ghost-row exchange through
    mpcx_ghost_reduce_f64) against the serial oracle assembly of the global problem; skipped on a single-GPU box.
    ``system``: through assemble_system (fused matrix + vector kernel, exchange afterwards); ``system-overlap``: its
    interface-first schedule (MPCX_OVERLAP=1), the exchange on a second stream beside the interior tiles."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    fused = mode != "routines"
    os.environ["MPCX_TEST_FUSED"] = "1" if fused else "0"
    os.environ["MPCX_OVERLAP"] = "1" if mode == "system-overlap" else "0"
    try:
        _run_and_compare(oracle, tmp_path, 2, periodic_z, cuda=True, n=9 if fused else 7, nzc=5 if fused else 4)
    finally:
        del os.environ["MPCX_TEST_FUSED"], os.environ["MPCX_OVERLAP"]


@pytest.mark.parametrize("world,periodic_z", [(2, False), (2, True), (3, False), (3, True)])
def test_multi_rank_assembly_matches_serial(oracle, tmp_path, world, periodic_z):
    """world = 3 adds a middle rank that owns one interface and ghosts another (the 4- and 8-GPU layouts)."""
    _run_and_compare(oracle, tmp_path, world, periodic_z)


def _run_and_compare(oracle, tmp_path, world, periodic_z, cuda=False, n=5, nzc=3):
    from dolfinx_mpc_b200 import fem, generators as gen

    port = 29600 + (os.getpid() % 200) + (1 if periodic_z else 0) + 2 * world + (50 if cuda else 0)
    mp.spawn(_worker, args=(world, port, periodic_z, n, nzc, str(tmp_path), cuda), nprocs=world, join=True)

    # serial global problem with the same lattice numbering
    h = 1.0 / (n - 1)
    mesh = gen.create_box(n - 1, n - 1, world * nzc, p1=(1.0, 1.0, world * nzc * h))
    V = gen.functionspace(mesh, 1)
    bcs, exclude = [], None
    if not periodic_z:
        exclude = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[2], 0) | np.isclose(x[2], world * nzc * h))
        bcs = [fem.DirichletBC(V, exclude, 0.25)]
    data = gen.periodic_constraint(V, axes=(0, 1, 2) if periodic_z else (0, 1), exclude_dofs=exclude)
    m = oracle.mpc_from_arrays(V, data)
    a = fem.laplace(V)
    f = fem.Function(V)
    f.interpolate(_f)
    L = fem.source(V, f)
    rp, col, val = oracle.assemble_matrix(a, m, bcs=bcs)
    b = oracle.assemble_vector(L, m)
    if bcs:
        oracle.apply_lifting(b, [a], [bcs], m)
    N = V.num_dofs
    S = sp.csr_matrix((val, col, rp), shape=(N, N))

    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    R = np.concatenate([p["r"] for p in parts])
    Cc = np.concatenate([p["c"] for p in parts])
    Vv = np.concatenate([p["v"] for p in parts])
    D_ = sp.csr_matrix((Vv, (R, Cc)), shape=(N, N))
    # same pattern (the union of the owners' rows is the serial pattern) ...
    Dp = sp.csr_matrix((np.ones_like(Vv), (R, Cc)), shape=(N, N))
    Sp = sp.csr_matrix((np.ones_like(val), col, rp), shape=(N, N))
    assert (Dp != Sp).nnz == 0
    assert len(R) == len(val), "owned rows hold duplicate or missing entries"
    # ... and the same values / right-hand side
    scale = np.abs(val).max()
    assert np.abs((D_ - S)).max() <= (1e-10 if cuda else 1e-12) * scale
    bd = np.concatenate([p["b"] for p in parts])
    assert np.allclose(bd, b, rtol=1e-10 if cuda else 1e-12, atol=1e-13 if cuda else 1e-14)
    assert sorted(np.concatenate([p["slaves"] for p in parts])) == sorted(data[0])
    if periodic_z:  # masters on rank 0 became extra ghosts of the last rank; rank 0 got pattern ghosts
        assert parts[-1]["nghost"] > 0 and parts[0]["ncol"] > (nzc + 1) * n * n


def test_slab_index_map_roundtrip():
    from dolfinx_mpc_b200 import distributed as D

    n, nzc, world = 4, 2, 3
    seen = np.zeros((world * nzc + 1) * n * n, dtype=int)
    for r in range(world):
        im = D.slab_index_map(n, nzc, r, world)
        loc = np.arange(im.size_local + im.num_ghosts)
        g = im.local_to_global(loc)
        assert np.array_equal(im.global_to_local(g), loc)
        seen[g[: im.size_local]] += 1
        assert np.all(im.owner_of_local(loc[im.size_local:]) == r + 1)
    assert np.all(seen == 1)
