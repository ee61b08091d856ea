"""GPU parity tests: the CUDA path (through the C ABI, libmpcx.so) against the CPU oracle on the same inputs.

Bar (BASELINE.json north_star): CSR ``row_ptr`` / ``col`` and all integer maps bit-exact; CSR values and RHS
within 1e-10 relative, with the floor ``1e-10 * ||row||_inf`` for entries that cancel to ~0 (SURVEY.md section 7,
"Relative tolerance near zero"); plus the reference's own identities ``K^H A K == A_mpc[free, free]``,
``K^H b == b_mpc[free]`` at its tolerance 5e-12 (python/src/dolfinx_mpc/utils/test.py:202-265).
"""
import numpy as np
import pytest
import scipy.sparse as sp

import problems

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _mpc(c):
    from dolfinx_mpc_b200 import MultiPointConstraint

    mpc = MultiPointConstraint(c.V)
    mpc.add_constraint(c.V, *c.data)
    mpc.finalize()
    return mpc


def assert_csr_close(rp, col, val, rp_o, col_o, val_o):
    assert np.array_equal(rp, rp_o), "row_ptr differs from the oracle"
    assert np.array_equal(col, col_o), "col differs from the oracle"
    n = len(rp) - 1
    rows = np.repeat(np.arange(n), np.diff(rp))
    rownorm = np.zeros(n)
    np.maximum.at(rownorm, rows, np.abs(val_o))
    # relative to the entry, with the floors "1e-10 of the row's largest entry" (entries that cancel inside a row) and
    # "1e-14 of the matrix' largest entry" (rows in which EVERYTHING cancels to round-off, e.g. rows of a div block on a
    # structured mesh; the reference's own comparison is absolute, atol = 5e-12: utils/test.py:196-199)
    tol = np.maximum(RTOL * np.maximum(np.abs(val_o), rownorm[rows]), 1e-14 * np.abs(val_o).max(initial=0.0))
    bad = np.abs(val - val_o) > tol
    assert not bad.any(), f"{bad.sum()} CSR values differ; worst {np.abs(val - val_o).max():.3e}"


def assert_vec_close(b, b_o):
    scale = np.abs(b_o).max() if len(b_o) else 1.0
    assert np.all(np.abs(b - b_o) <= RTOL * np.maximum(np.abs(b_o), scale)), np.abs(b - b_o).max()


@pytest.mark.parametrize("name", list(problems.ALL_CASES))
def test_matrix_vector_lifting_match_oracle(oracle, name):
    import dolfinx_mpc_b200 as mpcx

    c = problems.ALL_CASES[name]()
    mpc = _mpc(c)
    m = oracle.mpc_from_arrays(c.V, c.data)
    n = c.V.num_dofs

    # matrix
    A = mpcx.assemble_matrix(c.a, mpc, bcs=c.bcs, diagval=1.0)
    rp, col, val = A.getValuesCSR()
    rp_o, col_o, val_o = oracle.assemble_matrix(c.a, m, bcs=c.bcs, diagval=1.0)
    assert_csr_close(rp, col, val, rp_o, col_o, val_o)

    # re-assembly into the same matrix (cached pattern + plan) gives the same result
    mpcx.assemble_matrix(c.a, mpc, bcs=c.bcs, diagval=1.0, A=A)
    assert_csr_close(*A.getValuesCSR(), rp_o, col_o, val_o)

    # both scatter strategies: atomic-free tile gather (default where a tile kernel exists) and atomic scatter
    A2 = mpcx.create_matrix(c.a, mpc)
    A2.scatter = "atomic"
    mpcx.assemble_matrix(c.a, mpc, bcs=c.bcs, diagval=1.0, A=A2)
    assert_csr_close(*A2.getValuesCSR(), rp_o, col_o, val_o)

    # reference identity with an unconstrained oracle assembly
    e = oracle.OracleMPC.empty(c.V)
    A_org = sp.csr_matrix(oracle.assemble_matrix(c.a, e, bcs=c.bcs)[::-1], shape=(n, n))
    slaves, masters, coeffs, _, offsets = c.data
    K = oracle.transformation_matrix(n, slaves, masters, coeffs, offsets)
    oracle.compare_mpc_lhs(A_org, A.to_scipy(), K, slaves)

    if c.L is None:
        return
    b = mpcx.assemble_vector(c.L, mpc)
    b_o = oracle.assemble_vector(c.L, m)
    assert_vec_close(b.array, b_o)
    b0 = oracle.assemble_vector(c.L, e)
    if c.a_lift is not None and c.bcs:
        mpcx.apply_lifting(b, [c.a_lift], [c.bcs], mpc)
        oracle.apply_lifting(b_o, [c.a_lift], [c.bcs], m)
        assert_vec_close(b.array, b_o)
        oracle.apply_lifting(b0, [c.a_lift], [c.bcs], e)
    oracle.compare_mpc_rhs(b0, b.array, K, slaves)


@pytest.mark.parametrize("variant", [{"MPCX_ROWGATHER_V": "1"}, {"MPCX_ROWGATHER_V": "3"}, {"MPCX_ROWGATHER_GEO": "1"},
                                     {"MPCX_ROWGATHER_V": "3", "MPCX_ROWGATHER_GEO": "1"}, {"MPCX_ROWGATHER": "0"}],
                         ids=["warp-per-row", "flat-walk", "geometry-once", "flat-walk+geometry-once", "atomic"])
@pytest.mark.parametrize("name", ["contact3d", "slip3d-P2-2"])
def test_row_gather_variants_match_oracle(oracle, monkeypatch, name, variant):
    """The blocked-elasticity kernels kept behind switches (profiles/README.md r02_k) produce the same matrix as the
    default one (half-warp per row, lane per column) and the oracle: warp per row, the flat walk over a row's
    contributions, the geometry evaluated once per cell, and the atomic-scatter fallback."""
    import dolfinx_mpc_b200 as mpcx

    if name not in problems.ALL_CASES:
        pytest.skip("case not in this build of the problem set")
    for k, v in variant.items():
        monkeypatch.setenv(k, v)
    c = problems.ALL_CASES[name]()
    mpc = _mpc(c)
    m = oracle.mpc_from_arrays(c.V, c.data)
    A = mpcx.assemble_matrix(c.a, mpc, bcs=c.bcs, diagval=1.0)
    rp_o, col_o, val_o = oracle.assemble_matrix(c.a, m, bcs=c.bcs, diagval=1.0)
    assert_csr_close(*A.getValuesCSR(), rp_o, col_o, val_o)
    mpcx.assemble_matrix(c.a, mpc, bcs=c.bcs, diagval=1.0, A=A)  # cached plans
    assert_csr_close(*A.getValuesCSR(), rp_o, col_o, val_o)


@pytest.mark.parametrize("cfg,n", [(3, 10), (5, 21)])
def test_bench_configs_3_and_5_match_oracle(oracle, cfg, n):
    """The problems `bench.py --config 3 / 5` builds (P2 elasticity with slip on the inclined face of the rotated cube;
    two stacked boxes of P1 elasticity with a contact constraint), at a size the oracle finishes in seconds: matrix
    (row-gather kernel + slave-cell plan), vector and lifting entry for entry, through the same three calls the bench
    step makes."""
    import bench
    import dolfinx_mpc_b200 as mpcx

    P = bench.build_config(cfg, n)
    a, L, mpc, bcs = P["a"], P["L"], P["mpc"], P["bcs"]
    V = mpc.function_space
    m = oracle.WrappedMPC(V, mpc.is_slave, mpc.masters.array, mpc.coefficients()[0], mpc.masters.offsets,
                          mpc.cell_to_slaves.array, mpc.cell_to_slaves.offsets, mpc.slaves, mpc.num_local_slaves)
    A = mpcx.create_matrix(a, mpc)
    b = mpcx.create_vector(mpc)
    for _ in range(2):  # second pass through the cached plans
        mpcx.assemble_matrix(a, mpc, bcs=bcs, A=A)
        mpcx.assemble_vector(L, mpc, b=b)
        if bcs:
            mpcx.apply_lifting(b, [a], [bcs], mpc)
    assert any(k[0] == "row" for k in A._tile_plans if isinstance(k, tuple)), "the row-gather path did not run"
    rp_o, col_o, val_o = oracle.assemble_matrix(a, m, bcs=bcs)
    assert_csr_close(*A.getValuesCSR(), rp_o, col_o, val_o)
    b_o = oracle.assemble_vector(L, m)
    if bcs:
        oracle.apply_lifting(b_o, [a], [bcs], m)
    assert_vec_close(b.array, b_o)


@pytest.mark.parametrize("name", ["periodic2d-P1-8-bc1", "slip3d-P1-3", "contact3d"])
def test_lifting_with_x0_and_scale(oracle, name):
    import dolfinx_mpc_b200 as mpcx

    c = problems.ALL_CASES[name]()
    mpc = _mpc(c)
    m = oracle.mpc_from_arrays(c.V, c.data)
    rng = np.random.default_rng(5)
    x0 = rng.random(c.V.num_dofs)
    b = mpcx.assemble_vector(c.L, mpc)
    b_o = oracle.assemble_vector(c.L, m)
    mpcx.apply_lifting(b, [c.a_lift], [c.bcs], mpc, x0=[x0], scale=-0.7)
    oracle.apply_lifting(b_o, [c.a_lift], [c.bcs], m, x0=[x0], scale=-0.7)
    assert_vec_close(b.array, b_o)


def test_no_plan_path_matches(oracle):
    """Row-search scatter (plan == NULL) and planned scatter give the same matrix."""
    import dolfinx_mpc_b200 as mpcx

    c = problems.case_periodic_3d(4, 1, (0, 1), True)
    mpc = _mpc(c)
    m = oracle.mpc_from_arrays(c.V, c.data)
    A = mpcx.create_matrix(c.a, mpc)
    A.max_block_row = 1 << 20  # disables the plan
    mpcx.assemble_matrix(c.a, mpc, bcs=c.bcs, A=A)
    assert_csr_close(*A.getValuesCSR(), *oracle.assemble_matrix(c.a, m, bcs=c.bcs))


def test_backsubstitution_homogenize(oracle):
    from dolfinx_mpc_b200 import fem

    c = problems.case_contact_3d()
    mpc = _mpc(c)
    m = oracle.mpc_from_arrays(c.V, c.data)
    rng = np.random.default_rng(2)
    u = fem.Function(c.V, rng.random(c.V.num_dofs))
    u_o = u.array.copy()
    mpc.backsubstitution(u)
    oracle.backsubstitution(m, u_o)
    assert np.allclose(u.array, u_o, rtol=1e-14, atol=0)
    mpc.homogenize(u)
    oracle.homogenize(m, u_o)
    assert np.array_equal(u.array, u_o)


def test_pattern_error_is_loud():
    """An insertion outside the pattern must raise, never be dropped silently."""
    import dolfinx_mpc_b200 as mpcx
    from dolfinx_mpc_b200 import _lib

    c = problems.case_periodic_2d(6, 1, False)
    mpc = _mpc(c)
    empty = mpcx.MultiPointConstraint(c.V)
    empty.finalize()
    A = mpcx.create_matrix(c.a, empty)  # pattern without the master columns
    with pytest.raises(_lib.MpcxError):
        mpcx.assemble_matrix(c.a, mpc, A=A)


def test_rectangular_constraint_pair(oracle):
    """mpc0 != mpc1 (python/tests/test_rectangular_assembly.py): rows eliminated with one constraint, columns
    with another; no slave diagonal (cpp/assemble_matrix.cpp:711-724)."""
    import dolfinx_mpc_b200 as mpcx
    from dolfinx_mpc_b200 import generators as gen

    c = problems.case_periodic_2d(6, 1, False)
    mpc0 = _mpc(c)
    data1 = gen.periodic_constraint(c.V, axes=(1,), scale=0.5)
    mpc1 = mpcx.MultiPointConstraint(c.V)
    mpc1.add_constraint(c.V, *data1)
    mpc1.finalize()
    A = mpcx.assemble_matrix(c.a, (mpc0, mpc1))
    m0 = oracle.mpc_from_arrays(c.V, c.data)
    m1 = oracle.mpc_from_arrays(c.V, data1)
    assert_csr_close(*A.getValuesCSR(), *oracle.assemble_matrix(c.a, m0, m1, same_space=False))


def test_mid_size_properties(oracle):
    """A size the oracle still finishes in seconds (32^3 P1, periodic x/y): entry-wise parity, and the
    size-independent properties used at full BASELINE sizes: constant vectors are in the null space of the
    Laplace part on free rows, slave rows hold only the diagonal, sum(b) is preserved by K^T."""
    import dolfinx_mpc_b200 as mpcx
    from dolfinx_mpc_b200 import fem, generators as gen

    mesh = gen.create_unit_cube(32, 32, 32)
    V = gen.functionspace(mesh, 1)
    data = gen.periodic_constraint(V, axes=(0, 1))
    mpc = mpcx.MultiPointConstraint(V)
    mpc.add_constraint(V, *data)
    mpc.finalize()
    a = fem.laplace(V)
    f = fem.Function(V)
    f.interpolate(problems._f3d)
    L = fem.source(V, f)
    A = mpcx.assemble_matrix(a, mpc)
    b = mpcx.assemble_vector(L, mpc)
    m = oracle.mpc_from_arrays(V, data)
    assert_csr_close(*A.getValuesCSR(), *oracle.assemble_matrix(a, m))
    assert_vec_close(b.array, oracle.assemble_vector(L, m))
    S = A.to_scipy()
    free = np.setdiff1d(np.arange(V.num_dofs), mpc.slaves)
    ones = np.zeros(V.num_dofs)
    ones[free] = 1.0
    assert np.abs((S @ ones)[free]).max() < 1e-12
    assert np.array_equal(S.diagonal()[mpc.slaves], np.ones(len(mpc.slaves)))
    e = oracle.OracleMPC.empty(V)
    assert abs(b.array.sum() - oracle.assemble_vector(L, e).sum()) < 1e-12


def test_tile_plan_properties(oracle):
    """The tile path on a mesh whose node and cell numbering is scrambled (tiles must not rely on lexicographic
    locality): every bulk cell sits in exactly one tile, the number of reductions (dests) is well below the
    number of element entries, and the result matches the oracle for three integrals sharing one matrix."""
    import dolfinx_mpc_b200 as mpcx
    from dolfinx_mpc_b200 import fem, generators as gen

    mesh0 = gen.create_unit_cube(12, 10, 9)
    rng = np.random.default_rng(11)
    perm = rng.permutation(mesh0.x.shape[0]).astype(np.int32)  # old node -> new node
    x = np.empty_like(mesh0.x)
    x[perm] = mesh0.x
    cells = perm[mesh0.x_dofmap][rng.permutation(mesh0.num_cells)]
    mesh = fem.Mesh(x, cells, "tetrahedron")
    V = gen.functionspace(mesh, 1)
    bc_dofs = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[2], 0))
    bcs = [fem.DirichletBC(V, bc_dofs, 0.3)]
    data = gen.periodic_constraint(V, axes=(0,), exclude_dofs=bc_dofs)
    mpc = mpcx.MultiPointConstraint(V)
    mpc.add_constraint(V, *data)
    mpc.finalize()
    w = fem.Function(V)
    w.interpolate(lambda x: 1.0 + x[0] * x[1])
    a = fem.laplace(V, 1.5) + fem.mass(V, 0.25) + fem.laplace_varcoef(V, w, 2.0)
    A = mpcx.create_matrix(a, mpc)
    mpcx.assemble_matrix(a, mpc, bcs=bcs, A=A)
    infos = [info for _, info in A._tile_plans.values()]
    assert len(infos) == 3
    nbulk = mesh.num_cells_local - len(mpc.slave_cells)
    for info in infos:
        assert info["bulk_cells"] == nbulk and info["tiles"] == -(-nbulk // info["cells_per_tile"])
        assert info["symmetric"] == 1 and info["dests"] < 4 * nbulk  # 16 entries per cell before the per-tile combination
        assert info["runs"] < info["dests"]  # bulk reductions cover several CSR entries each
    m = oracle.mpc_from_arrays(V, data)
    ref = oracle.assemble_matrix(a, m, bcs=bcs)
    assert_csr_close(*A.getValuesCSR(), *ref)
    # the general (non-symmetric) tile plan gives the same matrix
    import os
    os.environ["MPCX_TILE_SYM"] = "0"
    try:
        A_g = mpcx.create_matrix(a, mpc)
        mpcx.assemble_matrix(a, mpc, bcs=bcs, A=A_g)
    finally:
        del os.environ["MPCX_TILE_SYM"]
    assert all(info["symmetric"] == 0 for _, info in A_g._tile_plans.values())
    assert_csr_close(*A_g.getValuesCSR(), *ref)
    # load vector through the vector tile plan (one reduction per (tile, row)), twice into the same vector
    L = fem.source(V, w, 0.7)
    b = mpcx.assemble_vector(L, mpc)
    plans = [v for k, v in L.integrals[0]._dev.items() if isinstance(k, tuple) and k[0] == "vector_tile_plan"]
    assert len(plans) == 1 and plans[0][2]["bulk_cells"] == nbulk and plans[0][2]["dests"] < 2 * nbulk
    b_o = oracle.assemble_vector(L, m)
    assert_vec_close(b.array, b_o)
    mpcx.assemble_vector(L, mpc, b=b)
    assert_vec_close(b.array, b_o)


def test_full_size_entry_for_entry(oracle):
    """BASELINE.json configs[1] at full size -- the problem bench.py times (256^3 P1 Poisson, periodic x/y,
    Dirichlet z; 99.5 M cells, 253 M nnz) -- assembled by the benchmarked call (the fused assemble_system) and
    compared ENTRY FOR ENTRY with the oracle: the same C routines as everywhere else, run slab-wise on all host
    threads (oracle.assemble_system_slabs; seconds at this size).  Bar: pattern bit-exact against the threaded
    host builder (itself bit-exact against the oracle on every fixture), values and right-hand side within 1e-10
    relative.  Also the separate assemble_matrix / assemble_vector / apply_lifting calls against the fused call, and
    size-independent properties (slave / Dirichlet rows hold exactly the diagonal; symmetry)."""
    import os

    import torch

    import bench
    import dolfinx_mpc_b200 as mpcx

    n = int(os.environ.get("MPCX_TEST_FULL_N", "256"))
    P = bench.build_problem(n)
    mesh, V, mpc, a, L, bcs = (P[k] for k in ("mesh", "V", "mpc", "a", "L", "bcs"))
    A = mpcx.create_matrix(a, mpc)  # pattern built on the device
    b = mpcx.create_vector(mpc)
    mpcx.assemble_system(a, L, mpc, bcs=bcs, A=A, b=b)
    assert A.last_system_fused and all(info["symmetric"] == 1 for _, info in A._tile_plans.values())
    rp, col, val = A.getValuesCSR()
    rp_h, col_h = mpcx.create_sparsity_pattern(a, mpc)
    assert np.array_equal(rp, rp_h) and np.array_equal(col, col_h), "device pattern differs from the host builder"
    del rp_h, col_h
    m = oracle.mpc_from_arrays(V, P["data"])
    val_o, b_o = oracle.assemble_system_slabs(a, L, m, bcs, (rp, col), 6 * (n - 1) ** 2, nthreads=os.cpu_count() or 1)
    assert_csr_close(rp, col, val, rp, col, val_o)
    assert_vec_close(b.array, b_o)
    # slave / Dirichlet rows: exactly the diagonal
    N = V.num_dofs
    rows = np.repeat(np.arange(N, dtype=np.int32), np.diff(rp))
    special = np.zeros(N, dtype=bool)
    special[mpc.slaves] = True
    special[bcs[0].dofs] = True
    sp_rows = special[rows]
    diag = rows == col
    assert np.all(val[sp_rows & ~diag] == 0.0) and np.all(val[sp_rows & diag] == 1.0)
    del rows, sp_rows, diag, val_o
    # the three separate calls give the same system (atomics / reductions may reorder sums)
    v_fused, b_fused = A.val.clone(), b.data.clone()
    mpcx.assemble_matrix(a, mpc, bcs=bcs, A=A)
    mpcx.assemble_vector(L, mpc, b=b)
    mpcx.apply_lifting(b, [a], [bcs], mpc)
    scale = float(v_fused.abs().max())
    assert float((A.val - v_fused).abs().max()) <= 1e-12 * scale
    assert float((b.data - b_fused).abs().max()) <= 1e-12 * float(b_fused.abs().max())
    # symmetry, on the device
    rows_d = torch.repeat_interleave(torch.arange(N, device=A.val.device), A.row_ptr[1:] - A.row_ptr[:-1])
    cols_d = A.col.long()

    def spmv(x):
        return torch.zeros(N, dtype=torch.float64, device=x.device).index_add_(0, rows_d, v_fused * x[cols_d])

    g = torch.Generator(device=A.val.device).manual_seed(3)
    x = torch.rand(N, dtype=torch.float64, device=A.val.device, generator=g)
    y = torch.rand(N, dtype=torch.float64, device=A.val.device, generator=g)
    xay, yax = float(x @ spmv(y)), float(y @ spmv(x))
    assert abs(xay - yax) <= 1e-10 * abs(xay)


@pytest.mark.parametrize("name", list(problems.ALL_CASES))
def test_device_pattern_matches_host_pattern(name):
    """create_sparsity_pattern on the device (sort + unique of 64-bit coupling keys) against the threaded host
    builder, itself bit-exact against the oracle's restatement of cpp/utils.h:381-496 (tests/test_cabi.py)."""
    import dolfinx_mpc_b200 as mpcx

    c = problems.ALL_CASES[name]()
    mpc = _mpc(c)
    rp_h, col_h = mpcx.create_sparsity_pattern(c.a, mpc)
    rp_d, col_d = mpcx.create_sparsity_pattern_device(c.a, mpc)
    assert np.array_equal(rp_d.cpu().numpy(), rp_h) and np.array_equal(col_d.cpu().numpy(), col_h)


def test_device_pattern_rectangular_pair():
    import dolfinx_mpc_b200 as mpcx
    from dolfinx_mpc_b200 import generators as gen

    c = problems.case_periodic_2d(6, 1, False)
    mpc0 = _mpc(c)
    mpc1 = mpcx.MultiPointConstraint(c.V)
    mpc1.add_constraint(c.V, *gen.periodic_constraint(c.V, axes=(1,), scale=0.5))
    mpc1.finalize()
    rp_h, col_h = mpcx.create_sparsity_pattern(c.a, (mpc0, mpc1))
    rp_d, col_d = mpcx.create_sparsity_pattern_device(c.a, (mpc0, mpc1))
    assert np.array_equal(rp_d.cpu().numpy(), rp_h) and np.array_equal(col_d.cpu().numpy(), col_h)


def test_device_handoff_spmv(oracle):
    """The assembled CSR consumed on the device without a copy of col / val: cuSPARSE SpMV through a
    torch.sparse_csr_tensor view equals the host SciPy product; DLPack round trip aliases the same memory."""
    import torch
    from torch.utils.dlpack import from_dlpack

    import dolfinx_mpc_b200 as mpcx

    c = problems.case_periodic_3d(4, 1, (0, 1), True)
    mpc = _mpc(c)
    A = mpcx.assemble_matrix(c.a, mpc, bcs=c.bcs)
    At = A.to_torch_sparse_csr()
    assert At.values().data_ptr() == A.val.data_ptr() and At.col_indices().data_ptr() == A.col.data_ptr()
    x = np.random.default_rng(4).random(A.shape[1])
    y = (At @ torch.from_numpy(x).to(A.val.device)).cpu().numpy()
    assert np.allclose(y, A.to_scipy() @ x, rtol=1e-13, atol=1e-13)
    rp, col, val = (from_dlpack(t) for t in A.dlpack())
    assert val.data_ptr() == A.val.data_ptr() and rp.dtype == torch.int64 and col.dtype == torch.int32


def test_tile_plan_fallback_on_a_fan_mesh(oracle):
    """A vertex shared by 700 triangles: its diagonal entry collects 448 sources inside one tile, more than a tile
    record holds (8-bit counts) -> the tile plan reports MPCX_ERR_UNSUPPORTED and the assembly must fall back to the
    atomic-scatter kernels on the device, with the same result."""
    import dolfinx_mpc_b200 as mpcx
    from dolfinx_mpc_b200 import fem, generators as gen

    n = 700
    ang = 2 * np.pi * np.arange(n) / n
    x = np.zeros((n + 1, 3))
    x[1:, 0], x[1:, 1] = np.cos(ang), np.sin(ang)
    cells = np.stack([np.zeros(n, np.int32), 1 + np.arange(n, dtype=np.int32), 1 + (np.arange(n, dtype=np.int32) + 1) % n], 1)
    mesh = fem.Mesh(x, cells.astype(np.int32), "triangle")
    V = gen.functionspace(mesh, 1)
    mpc = mpcx.MultiPointConstraint(V)
    mpc.add_constraint(V, *gen.empty_constraint())
    mpc.finalize()
    a = fem.laplace(V) + fem.mass(V, 0.5)
    f = fem.Function(V)
    f.interpolate(lambda x: 1.0 + x[0])
    L = fem.source(V, f)
    A = mpcx.assemble_matrix(a, mpc)
    assert list(A._tile_plans.values()) == [None, None], "the tile plan should not fit this mesh"
    m = oracle.mpc_from_arrays(V, gen.empty_constraint())
    assert_csr_close(*A.getValuesCSR(), *oracle.assemble_matrix(a, m))
    assert_vec_close(mpcx.assemble_vector(L, mpc).array, oracle.assemble_vector(L, m))


@pytest.mark.parametrize("get_assemblers", ["cuda"], indirect=True)
@pytest.mark.parametrize("master_point", [[1, 1], [0, 1]])
def test_mpc_assembly_reference_style(oracle, get_assemblers, master_point):
    """python/tests/test_matrix_assembly.py:20-57 transcribed: same fixture switch, same dictionary constraint, same
    comparison K^T A K == A_mpc through the transformation matrix -- with "cuda" as the assembler."""
    from dolfinx_mpc_b200 import MultiPointConstraint, fem, generators as gen

    assemble_matrix, _ = get_assemblers
    mesh = gen.create_unit_square(5, 3)
    V = gen.functionspace(mesh, 1)
    a = fem.laplace(V)
    s_m_c = {(1, 0): {(0, 1): 0.43, (1, 1): 0.11}, (0, 0): {tuple(master_point): 0.69}}
    mpc = MultiPointConstraint(V)
    mpc.create_general_constraint(s_m_c)
    mpc.finalize()
    A_mpc = assemble_matrix(a, mpc)
    e = oracle.OracleMPC.empty(V)
    n = V.num_dofs
    A_org = sp.csr_matrix(oracle.assemble_matrix(a, e)[::-1], shape=(n, n))
    data = gen.general_constraint(V, s_m_c)
    K = oracle.transformation_matrix(n, data[0], data[1], data[2], data[4])
    oracle.compare_mpc_lhs(A_org, A_mpc.to_scipy(), K, data[0])


@pytest.mark.parametrize("name", [k for k in problems.ALL_CASES])
def test_assemble_system_matches_oracle(oracle, name):
    """assemble_system (the assembly block of LinearProblem.solve, python/src/dolfinx_mpc/problem.py:539-572) against the
    oracle's assemble_matrix + assemble_vector + apply_lifting: the fused matrix + vector tile kernel where the
    forms allow it (one scalar P1 cell integral each), the three separate routines otherwise."""
    import dolfinx_mpc_b200 as mpcx

    c = problems.ALL_CASES[name]()
    if c.L is None:
        pytest.skip("no linear form")
    mpc = _mpc(c)
    m = oracle.mpc_from_arrays(c.V, c.data)
    a_lift = c.a_lift if c.a_lift is not None else c.a
    bcs = c.bcs if c.a_lift is not None else []
    A, b = mpcx.assemble_system(c.a, c.L, mpc, bcs=bcs)
    fusable = (len(c.a.integrals) == 1 and len(c.L.integrals) == 1 and c.V.bs == 1 and c.V.degree == 1
               and c.V.mesh.cell_type in ("triangle", "tetrahedron") and int(c.a.integrals[0].kernel) in (0, 1, 4)
               and c.a.integrals[0].cells is None and c.L.integrals[0].cells is None)
    assert A.last_system_fused == fusable, name
    assert_csr_close(*A.getValuesCSR(), *oracle.assemble_matrix(c.a, m, bcs=bcs))
    b_o = oracle.assemble_vector(c.L, m)
    if bcs:
        oracle.apply_lifting(b_o, [a_lift], [bcs], m)
    assert_vec_close(b.array, b_o)
    # again into the same objects (cached plans), with x0 and scale
    x0 = np.random.default_rng(7).random(c.V.num_dofs)
    mpcx.assemble_system(c.a, c.L, mpc, bcs=bcs, A=A, b=b, x0=[x0], scale=0.6)
    assert_csr_close(*A.getValuesCSR(), *oracle.assemble_matrix(c.a, m, bcs=bcs))
    b_o = oracle.assemble_vector(c.L, m)
    if bcs:
        oracle.apply_lifting(b_o, [a_lift], [bcs], m, x0=[x0], scale=0.6)
    assert_vec_close(b.array, b_o)


@pytest.mark.parametrize("switch", [{"MPCX_TILE_CURVE": "hilbert"}, {"MPCX_TILE_ALIGN": "0"},
                                    {"MPCX_TILE_CURVE": "hilbert", "MPCX_TILE_ALIGN": "0"}],
                         ids=["hilbert", "bounding-box-scale", "hilbert+bounding-box-scale"])
@pytest.mark.parametrize("name", ["periodic3d-P1-bs1-4-ax2-bc1", "periodic2d-P1-8-bc1"])
def test_tile_order_switches_keep_parity(oracle, monkeypatch, name, switch):
    """The order of the cells along the space-filling curve (Morton on the cell-aligned lattice by default; Hilbert, or the
    plain bounding-box scale, behind switches: profiles/README.md r02_r) changes the tile plans, never the result."""
    import dolfinx_mpc_b200 as mpcx

    for k, v in switch.items():
        monkeypatch.setenv(k, v)
    c = problems.ALL_CASES[name]()
    mpc = _mpc(c)
    m = oracle.mpc_from_arrays(c.V, c.data)
    a_lift = c.a_lift if c.a_lift is not None else c.a
    bcs = c.bcs if c.a_lift is not None else []
    A, b = mpcx.assemble_system(c.a, c.L, mpc, bcs=bcs)  # fused kernel, or the separate tile kernels: same plans
    assert A._tile_plans, "no tile plan was built"
    assert_csr_close(*A.getValuesCSR(), *oracle.assemble_matrix(c.a, m, bcs=bcs))
    b_o = oracle.assemble_vector(c.L, m)
    if bcs:
        oracle.apply_lifting(b_o, [a_lift], [bcs], m)
    assert_vec_close(b.array, b_o)


def test_fused_system_general_plan_and_scrambled_mesh(oracle):
    """The fused kernel on a mesh with scrambled node / cell numbering, once with the symmetric matrix plan and once
    with the general one (MPCX_TILE_SYM=0), variable-coefficient Laplace + source sharing one coefficient."""
    import os

    import dolfinx_mpc_b200 as mpcx
    from dolfinx_mpc_b200 import fem, generators as gen

    mesh0 = gen.create_unit_cube(11, 9, 10)
    rng = np.random.default_rng(12)
    perm = rng.permutation(mesh0.x.shape[0]).astype(np.int32)
    x = np.empty_like(mesh0.x)
    x[perm] = mesh0.x
    mesh = fem.Mesh(x, perm[mesh0.x_dofmap][rng.permutation(mesh0.num_cells)], "tetrahedron")
    V = gen.functionspace(mesh, 1)
    bc_dofs = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[2], 0))
    bcs = [fem.DirichletBC(V, bc_dofs, 0.3)]
    data = gen.periodic_constraint(V, axes=(0,), exclude_dofs=bc_dofs)
    mpc = mpcx.MultiPointConstraint(V)
    mpc.add_constraint(V, *data)
    mpc.finalize()
    w = fem.Function(V)
    w.interpolate(lambda x: 1.0 + x[0] * x[1])
    a, L = fem.laplace_varcoef(V, w, 2.0), fem.source(V, w, 0.7)
    m = oracle.mpc_from_arrays(V, data)
    ref = oracle.assemble_matrix(a, m, bcs=bcs)
    b_o = oracle.assemble_vector(L, m)
    oracle.apply_lifting(b_o, [a], [bcs], m)
    for sym in ("1", "0"):
        os.environ["MPCX_TILE_SYM"] = sym
        try:
            A, b = mpcx.assemble_system(a, L, mpc, bcs=bcs)
        finally:
            del os.environ["MPCX_TILE_SYM"]
        assert A.last_system_fused and all(info["symmetric"] == int(sym) for _, info in A._tile_plans.values())
        assert_csr_close(*A.getValuesCSR(), *ref)
        assert_vec_close(b.array, b_o)


def test_lifting_rereads_bc_values_changed_in_place(oracle):
    """A Function-valued Dirichlet condition updated IN PLACE between two apply_lifting calls (the usual DOLFINx
    pattern for time-dependent conditions): the second call must see the new values, as the reference re-reads
    them on every call (cpp/lifting.h:166-180)."""
    import dolfinx_mpc_b200 as mpcx
    from dolfinx_mpc_b200 import fem

    c = problems.case_periodic_3d(4, 1, (0, 1), True)
    mpc = _mpc(c)
    m = oracle.mpc_from_arrays(c.V, c.data)
    g = fem.Function(c.V)
    g.interpolate(lambda x: 1.0 + x[0] + 2.0 * x[1])
    bcs = [fem.DirichletBC(c.V, c.bcs[0].dofs, g)]
    for k in range(3):
        if k == 1:
            g.interpolate(lambda x: -3.0 * x[0] + x[1] ** 2)  # Function.interpolate writes g.array in place
        if k == 2:
            g.array[:] = 0.5 * g.array + 0.25
        b = mpcx.assemble_vector(c.L, mpc)
        mpcx.apply_lifting(b, [c.a_lift], [bcs], mpc)
        b_o = oracle.assemble_vector(c.L, m)
        oracle.apply_lifting(b_o, [c.a_lift], [bcs], m)
        assert_vec_close(b.array, b_o)
    # reassigning .value (scalar) is seen as well
    bcs[0].value = 0.125
    b = mpcx.assemble_vector(c.L, mpc)
    mpcx.apply_lifting(b, [c.a_lift], [bcs], mpc)
    b_o = oracle.assemble_vector(c.L, m)
    oracle.apply_lifting(b_o, [c.a_lift], [bcs], m)
    assert_vec_close(b.array, b_o)


def test_async_zero_double_buffer(oracle):
    """Matrix.async_zero: the values alternate between two buffers, the idle one cleared on a side stream during the
    assembly into the other; every assembly must give the full result (nothing left over, nothing cleared late)."""
    import torch

    import dolfinx_mpc_b200 as mpcx

    c = problems.case_periodic_3d(4, 1, (0, 1), True)
    mpc = _mpc(c)
    m = oracle.mpc_from_arrays(c.V, c.data)
    ref = oracle.assemble_matrix(c.a, m, bcs=c.bcs)
    A = mpcx.create_matrix(c.a, mpc)
    A.async_zero = True
    ptrs = set()
    for _ in range(5):
        mpcx.assemble_matrix(c.a, mpc, bcs=c.bcs, A=A)
        ptrs.add(A.val.data_ptr())
        torch.cuda.synchronize()
        assert_csr_close(*A.getValuesCSR(), *ref)
    assert len(ptrs) == 2


def test_deferred_error_is_still_loud():
    """Matrix.deferred_errors: the device error flag travels asynchronously, but an insertion outside the pattern
    still raises -- at synchronize() or at the next assembly into the matrix, whichever comes first."""
    import dolfinx_mpc_b200 as mpcx
    from dolfinx_mpc_b200 import _lib

    c = problems.case_periodic_2d(6, 1, False)
    mpc = _mpc(c)
    empty = mpcx.MultiPointConstraint(c.V)
    empty.finalize()
    A = mpcx.create_matrix(c.a, empty)  # pattern without the master columns
    A.deferred_errors = True
    try:
        mpcx.assemble_matrix(c.a, mpc, A=A)  # may or may not raise here (plan construction checks eagerly)
        with pytest.raises(_lib.MpcxError):
            A.synchronize()
    except _lib.MpcxError:
        pass
    # and a clean matrix stays clean
    B = mpcx.create_matrix(c.a, mpc)
    B.deferred_errors = True
    for _ in range(3):
        mpcx.assemble_matrix(c.a, mpc, A=B)
    B.synchronize()
