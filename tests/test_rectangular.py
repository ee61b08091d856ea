"""Rectangular blocks whose test and trial ELEMENTS differ (python/tests/test_rectangular_assembly.py:25-199: the
Taylor-Hood blocks a01 = -inner(p, div v) dx, a10 = -inner(div u, q) dx with a slip constraint on the velocity and an
empty constraint on the pressure; cpp/assemble_matrix.cpp:99-117 takes dofs[2], bs[2], num_dofs[2]).

CPU part: the oracle's div coupling kernels against an independent numpy quadrature, and the rectangular form of the
reference's identity, K_0^T A K_1 == A_mpc[free_0, free_1] (python/src/dolfinx_mpc/utils/test.py:202-242 with two
transformation matrices).  GPU part (-m gpu): every block of the nest against the oracle entry for entry, the nest
wrappers, and the nest norm against the norm of K_0^T A K_1 (the reference test's final check, :193)."""
import numpy as np
import pytest
import scipy.sparse as sp

import problems


def _K(V, data):
    from oracle import oracle as orc

    slaves, masters, coeffs, _, offsets = data
    return orc.transformation_matrix(V.num_dofs, slaves, masters, coeffs, offsets)


def _free(V, data):
    return np.setdiff1d(np.arange(V.num_dofs), np.asarray(data[0]))


def test_div_kernels_against_independent_quadrature(oracle):
    """A_e[(i, a), j] = c0 * int phi^Q_j d_a phi^V_i on a generic (non-degenerate, non-axis-aligned) simplex, from the
    oracle's tabulated kernel vs a direct numpy evaluation with a much higher quadrature degree; DIV_TRIAL is its
    transpose."""
    from dolfinx_mpc_b200 import elements as el

    rng = np.random.default_rng(3)
    for cell, tdim in (("triangle", 2), ("tetrahedron", 3)):
        X = np.zeros((tdim + 1, 3))
        X[:, :tdim] = el.reference_vertices(cell) + 0.2 * rng.random((tdim + 1, tdim))
        tab = el.mixed_element_tables(cell, 2, 1, 2)
        A01 = oracle.tabulate(5, tab, tdim, X, c=(-1.3,), bs1=1)
        tabT = el.mixed_element_tables(cell, 1, 2, 2)
        A10 = oracle.tabulate(6, tabT, 1, X, c=(-1.3,), bs1=tdim)
        assert A01.shape == (tab.nd * tdim, tdim + 1) and np.allclose(A10, A01.T, rtol=1e-13, atol=1e-15)
        pts, wts = el.make_quadrature(cell, 6)
        phiV, dphiV = el.tabulate(cell, 2, pts)
        phiQ, _ = el.tabulate(cell, 1, pts)
        _, gd = el.tabulate(cell, 1, pts[:1])
        J = X[:, :tdim].T @ gd[0].T  # J[k, a] = sum_g X[g, k] dpsi_g/dxi_a
        Kinv = np.linalg.inv(J)
        ref = np.zeros_like(A01)
        for q in range(len(wts)):
            g = dphiV[q].T @ Kinv  # (nd, gdim): physical gradients
            ref += -1.3 * wts[q] * abs(np.linalg.det(J)) * np.einsum("ia,j->iaj", g, phiQ[q]).reshape(ref.shape)
        assert np.allclose(A01, ref, rtol=1e-12, atol=1e-14)


def _oracle_blocks(oracle, c):
    mv = oracle.mpc_from_arrays(c.V, c.data_v)
    mq = oracle.OracleMPC.empty(c.Q)
    ms = [mv, mq]
    out = [[None, None], [None, None]]
    for i in range(2):
        for j in range(2):
            if c.a[i][j] is not None:
                out[i][j] = oracle.assemble_matrix(c.a[i][j], ms[i], ms[j], bcs=c.bcs, same_space=(i == j))
    return out, ms


def test_rectangular_identity_on_the_oracle(oracle):
    """K_0^T A_ij K_1 == A_mpc,ij[free_0, free_1] for every block of the Stokes nest (bcs applied on both sides of the
    unconstrained assembly as well, as the reference test does with dolfinx.fem.petsc.assemble_matrix(a_nest, bcs))."""
    from dolfinx_mpc_b200 import generators as gen

    c = problems.case_stokes_2d(4)
    blocks, _ = _oracle_blocks(oracle, c)
    spaces, datas = [c.V, c.Q], [c.data_v, gen.empty_constraint()]
    for i in range(2):
        for j in range(2):
            if c.a[i][j] is None:
                continue
            rp, col, val = blocks[i][j]
            A_mpc = sp.csr_matrix((val, col, rp), shape=(spaces[i].num_dofs, spaces[j].num_dofs))
            e0, e1 = oracle.OracleMPC.empty(spaces[i]), oracle.OracleMPC.empty(spaces[j])
            rp0, col0, val0 = oracle.assemble_matrix(c.a[i][j], e0, e1, bcs=c.bcs, same_space=(i == j))
            A_org = sp.csr_matrix((val0, col0, rp0), shape=A_mpc.shape)
            K0, K1 = _K(spaces[i], datas[i]), _K(spaces[j], datas[j])
            red = A_mpc[_free(spaces[i], datas[i]), :][:, _free(spaces[j], datas[j])]
            ref = (K0.T @ A_org @ K1).tocsr()[: red.shape[0], : red.shape[1]]
            assert abs(ref - red).max() < 5e-12 * max(1.0, abs(A_org).max()), (i, j)
            if i != j:  # off-diagonal blocks get no slave diagonal (cpp/assemble_matrix.cpp:711-724)
                assert abs(A_mpc[np.asarray(datas[i][0], dtype=np.int64), :]).max() == 0 if len(datas[i][0]) else True


@pytest.mark.gpu
def test_stokes_nest_blocks_match_oracle(oracle):
    import dolfinx_mpc_b200 as mpcx
    from test_gpu_parity import assert_csr_close, assert_vec_close

    for cell, n in (("triangle", 4), ("triangle", 7)):
        c = problems.case_stokes_2d(n, cell=cell)
        mpc_v = mpcx.MultiPointConstraint(c.V)
        mpc_v.add_constraint(c.V, *c.data_v)
        mpc_v.finalize()
        mpc_q = mpcx.MultiPointConstraint(c.Q)
        mpc_q.finalize()
        mpcs = [mpc_v, mpc_q]
        A = mpcx.create_matrix_nest(c.a, mpcs)
        assert A[1][1] is None and A[0][1].shape == (c.V.num_dofs, c.Q.num_dofs) and A[1][0].shape == (c.Q.num_dofs, c.V.num_dofs)
        mpcx.assemble_matrix_nest(A, c.a, mpcs, c.bcs)
        blocks, ms = _oracle_blocks(oracle, c)
        norm2 = 0.0
        for i in range(2):
            for j in range(2):
                if c.a[i][j] is None:
                    continue
                assert_csr_close(*A[i][j].getValuesCSR(), *blocks[i][j])
                norm2 += A[i][j].norm() ** 2
        # nest norm == norm of the oracle's nest (python/tests/test_rectangular_assembly.py:193 compares nest and monolithic)
        ref2 = sum(np.linalg.norm(blocks[i][j][2]) ** 2 for i in range(2) for j in range(2) if blocks[i][j] is not None)
        assert np.isclose(np.sqrt(norm2), np.sqrt(ref2), rtol=1e-12)
        # a01 == a10^T as matrices (both carry the same constraint on the velocity side)
        assert abs(A[0][1].to_scipy() - A[1][0].to_scipy().T).max() < 1e-12 * abs(A[0][1].to_scipy()).max()
        # nest right-hand side + lifting of the velocity Dirichlet values through BOTH blocks of the first column
        b = mpcx.create_vector_nest(c.L, mpcs)
        mpcx.assemble_vector_nest(b, c.L, mpcs)
        mpcx.apply_lifting(b[0], [c.a[0][0]], [c.bcs], mpc_v)
        mpcx.apply_lifting(b[1], [c.a[1][0]], [c.bcs], mpc_q)  # rectangular lifting: rows in Q, bc columns in V
        b0 = oracle.assemble_vector(c.L[0], ms[0])
        b1 = oracle.assemble_vector(c.L[1], ms[1])
        oracle.apply_lifting(b0, [c.a[0][0]], [c.bcs], ms[0])
        oracle.apply_lifting(b1, [c.a[1][0]], [c.bcs], ms[1])
        assert_vec_close(b[0].array, b0)
        assert_vec_close(b[1].array, b1)
        assert np.abs(b1).max() > 0  # the lifting through a10 is not trivially zero
