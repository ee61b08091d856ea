"""Randomised constraints: arbitrary slave sets, 1-4 masters per slave anywhere in the mesh, random coefficients,
Dirichlet conditions on a random subset -- the structures the structured fixtures never produce (several slaves per
cell, masters shared between slaves, masters far away, slaves in cells with Dirichlet dofs).

CPU: the oracle must satisfy the reference's identities K^T A K == A_mpc / K^T b == b_mpc
(python/src/dolfinx_mpc/utils/test.py:202-265) on every draw.  GPU: the CUDA path must match the oracle entry for
entry on the same draws."""
import numpy as np
import pytest
import scipy.sparse as sp

from dolfinx_mpc_b200 import fem, generators as gen

CONFIGS = [("triangle", 1, 1, 6), ("triangle", 2, 1, 4), ("tetrahedron", 1, 1, 3), ("tetrahedron", 1, 3, 3),
           ("tetrahedron", 2, 1, 2), ("quadrilateral", 1, 2, 5), ("hexahedron", 1, 1, 3)]


def draw(cell, degree, bs, n, seed):
    rng = np.random.default_rng(1000 * seed + 17 * degree + bs)
    if cell in ("triangle", "quadrilateral"):
        mesh = gen.create_unit_square(n, n + 1, cell)
    else:
        mesh = gen.create_unit_cube(n, n, n + 1, cell)
    mesh.x[:, : mesh.tdim] += 0.15 / n * (rng.random((mesh.x.shape[0], mesh.tdim)) - 0.5)  # non-uniform geometry
    V = gen.functionspace(mesh, degree, bs)
    N = V.num_dofs
    perm = rng.permutation(N)
    n_bc, n_sl = max(1, N // 12), max(2, N // 8)
    bc_dofs = np.sort(perm[:n_bc]).astype(np.int32)
    slaves = np.sort(perm[n_bc:n_bc + n_sl]).astype(np.int32)
    free = perm[n_bc + n_sl:]  # masters: neither slave nor Dirichlet (a master on a Dirichlet dof is legal in the
    masters, coeffs, offsets = [], [], [0]  # reference but the K^T A K comparison then needs the bc on K as well)
    for _ in slaves:
        k = int(rng.integers(1, 5))
        masters += list(rng.choice(free, size=k, replace=False))
        coeffs += list(rng.uniform(-1.0, 1.0, size=k))
        offsets.append(len(masters))
    data = (slaves, np.array(masters, np.int64), np.array(coeffs, np.float64), np.zeros(len(masters), np.int32),
            np.array(offsets, np.int32))
    bcs = [fem.DirichletBC(V, bc_dofs, 0.37)]
    if bs == mesh.tdim and bs > 1:
        a = fem.elasticity(V, 1.3, 0.6) + fem.mass(V, 0.2)
    else:
        a = fem.laplace(V, 1.1) + fem.mass(V, 0.4)
    f = fem.Function(V)
    f.array[:] = rng.random(N)
    return V, a, fem.source(V, f, 0.9), data, bcs


@pytest.mark.parametrize("cfg", CONFIGS, ids=[f"{c}-P{d}-bs{b}" for c, d, b, _ in CONFIGS])
@pytest.mark.parametrize("seed", [0, 1])
def test_oracle_identities_on_random_constraints(oracle, cfg, seed):
    V, a, L, data, bcs = draw(*cfg, seed)
    m = oracle.mpc_from_arrays(V, data)
    e = oracle.OracleMPC.empty(V)
    n = V.num_dofs
    A = sp.csr_matrix(oracle.assemble_matrix(a, m, bcs=bcs)[::-1], shape=(n, n))
    A_org = sp.csr_matrix(oracle.assemble_matrix(a, e, bcs=bcs)[::-1], shape=(n, n))
    K = oracle.transformation_matrix(n, data[0], data[1], data[2], data[4])
    oracle.compare_mpc_lhs(A_org, A, K, data[0])
    b = oracle.assemble_vector(L, m)
    oracle.apply_lifting(b, [a], [bcs], m)
    b_org = oracle.assemble_vector(L, e)
    oracle.apply_lifting(b_org, [a], [bcs], e)
    oracle.compare_mpc_rhs(b_org, b, K, data[0])


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", CONFIGS, ids=[f"{c}-P{d}-bs{b}" for c, d, b, _ in CONFIGS])
@pytest.mark.parametrize("seed", [0, 1])
def test_cuda_matches_oracle_on_random_constraints(oracle, cfg, seed):
    import dolfinx_mpc_b200 as mpcx
    from test_gpu_parity import assert_csr_close, assert_vec_close

    V, a, L, data, bcs = draw(*cfg, seed)
    mpc = mpcx.MultiPointConstraint(V)
    mpc.add_constraint(V, *data)
    mpc.finalize()
    m = oracle.mpc_from_arrays(V, data)
    A = mpcx.assemble_matrix(a, mpc, bcs=bcs)
    assert_csr_close(*A.getValuesCSR(), *oracle.assemble_matrix(a, m, bcs=bcs))
    b = mpcx.assemble_vector(L, mpc)
    mpcx.apply_lifting(b, [a], [bcs], mpc)
    b_o = oracle.assemble_vector(L, m)
    oracle.apply_lifting(b_o, [a], [bcs], m)
    assert_vec_close(b.array, b_o)
    # backsubstitution on a random vector
    u = np.random.default_rng(seed).random(V.num_dofs)
    uf = fem.Function(V, u.copy())
    mpc.backsubstitution(uf)
    oracle.backsubstitution(m, u)
    assert np.allclose(uf.array, u, rtol=1e-13, atol=1e-14)  # sums of up to 4 signed terms: fma vs separate rounding


@pytest.mark.parametrize("cfg", CONFIGS, ids=[f"{c}-P{d}-bs{b}" for c, d, b, _ in CONFIGS])
def test_host_data_model_and_pattern_on_random_constraints(oracle, cfg):
    """MultiPointConstraint.finalize (cpp/MultiPointConstraint.h:36-126) and the host pattern builder
    (cpp/utils.h:381-496) against the oracle's restatements, bit for bit, on a random draw."""
    import dolfinx_mpc_b200 as mpcx

    V, a, L, data, bcs = draw(*cfg, seed=3)
    mpc = mpcx.MultiPointConstraint(V)
    mpc.add_constraint(V, *data)
    mpc.finalize()
    m = oracle.mpc_from_arrays(V, data)
    assert np.array_equal(mpc.is_slave, m.is_slave) and np.array_equal(mpc.slaves, m.slaves)
    assert np.array_equal(mpc.masters.offsets, m.offsets) and np.array_equal(mpc.masters.array, m.masters)
    assert np.array_equal(mpc.coefficients()[0], m.coeffs)
    assert np.array_equal(mpc.cell_to_slaves.offsets, m.c2s_offsets) and np.array_equal(mpc.cell_to_slaves.array, m.c2s)
    assert mpc.num_local_slaves == m.num_local_slaves
    rp, col = mpcx.create_sparsity_pattern(a, mpc, num_threads=3)
    rp_o, col_o = oracle.create_pattern(a, m, m)
    assert np.array_equal(rp, rp_o) and np.array_equal(col, col_o)
