"""Replay of dumps made by the REFERENCE itself (tools/dump_reference.py, run on a box with DOLFINx + dolfinx_mpc):
the dumped inputs go through the oracle (CPU) and through libmpcx (GPU, -m gpu) and the results are compared with the
matrix / vector the reference assembled.  No dump can be produced in this repository's build image (neither DOLFINx
nor PETSc is installed), so without files under tests/replay/ these tests skip -- and parity
stays "unpinned against the reference binary", pinned only through the reference tests' identities.

The loader itself (dump layout -> Form / MultiPointConstraint with the constraint arrays taken verbatim) is exercised
on a synthetic dump written in the same layout from this repository's generators.
"""
import glob
import os

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DUMPS = sorted(glob.glob(os.path.join(ROOT, "tests", "replay", "*.npz")))


def load_dump(path):
    """Rebuild the array-backed descriptors from a dump.  The constraint is NOT re-derived: the dumped dof-indexed
    adjacency arrays (masters / coeffs / offsets / is_slave / cell_to_slaves) are installed verbatim."""
    from dolfinx_mpc_b200 import MultiPointConstraint, fem
    from dolfinx_mpc_b200.multipointconstraint import AdjacencyList

    d = np.load(path, allow_pickle=False)
    mesh = fem.Mesh(d["x"], d["x_dofmap"], str(d["cell_type"]))
    bs = int(d["bs"])
    nblocks = len(d["is_slave"]) // bs
    V = fem.FunctionSpace(mesh, int(d["degree"]), d["dofmap"], bs, fem.IndexMap(nblocks), d["dof_coordinates"])
    mpc = MultiPointConstraint(V)
    slaves = np.asarray(d["mpc_slaves"], np.int32)
    off = np.asarray(d["mpc_offsets"], np.int32)
    masters = np.asarray(d["mpc_masters"], np.int32)
    coeffs = np.asarray(d["mpc_coeffs"], np.float64)
    # add_constraint wants per-slave CSR in input order: slice the dof-indexed lists
    per = [(masters[off[s]:off[s + 1]], coeffs[off[s]:off[s + 1]]) for s in slaves]
    o2 = np.concatenate([[0], np.cumsum([len(p[0]) for p in per])]).astype(np.int32)
    mpc.add_constraint(V, slaves, np.concatenate([p[0] for p in per]).astype(np.int64) if per else np.zeros(0, np.int64),
                       np.concatenate([p[1] for p in per]) if per else np.zeros(0), np.zeros(int(o2[-1]), np.int32), o2)
    mpc.finalize()
    # the packed data must come out exactly as the reference packed it
    assert np.array_equal(mpc.is_slave, d["is_slave"]) and np.array_equal(mpc.masters.offsets, off)
    assert np.array_equal(mpc.masters.array, masters) and np.array_equal(mpc.cell_to_slaves.offsets, d["c2s_offsets"])
    assert np.array_equal(mpc.cell_to_slaves.array, d["c2s"]) and mpc.num_local_slaves == int(d["num_local_slaves"])
    bcv = d["bc_value"]
    bcs = [fem.DirichletBC(V, d["bc_dofs"], float(bcv) if bcv.ndim == 0 else bcv)]
    const = d["constants"]
    a = fem.laplace(V, float(const[0])) if str(d["kernel"]) == "laplace" else fem.elasticity(V, float(const[0]), float(const[1]))
    f = fem.Function(V, d["f"])
    L = fem.source(V, f)
    N = V.num_dofs
    A_ref = sp.csr_matrix((d["A_data"], d["A_indices"], d["A_indptr"]), shape=(N, N))
    return dict(V=V, mpc=mpc, bcs=bcs, a=a, L=L, A_ref=A_ref, b_ref=np.asarray(d["b"]), data=d)


def _compare(A, b, P):
    A_ref, b_ref = P["A_ref"], P["b_ref"]
    scale = abs(A_ref).max()
    assert abs(A - A_ref).max() <= 1e-10 * scale, "matrix differs from the reference's"
    # the reference's matrix holds no entry outside our pattern
    ours = sp.csr_matrix((np.ones_like(A.data), A.indices, A.indptr), shape=A.shape)
    ref1 = sp.csr_matrix((np.ones_like(A_ref.data), A_ref.indices, A_ref.indptr), shape=A.shape)
    assert (ref1 - ref1.multiply(ours)).nnz == 0
    assert np.abs(b - b_ref).max() <= 1e-10 * max(1.0, np.abs(b_ref).max())


@pytest.mark.skipif(not DUMPS, reason="no reference dumps under tests/replay (needs a DOLFINx box: tools/dump_reference.py)")
@pytest.mark.parametrize("path", DUMPS)
def test_oracle_reproduces_reference_dump(oracle, path):
    P = load_dump(path)
    mpc = P["mpc"]
    m = oracle.WrappedMPC(P["V"], mpc.is_slave, mpc.masters.array, mpc.coefficients()[0], mpc.masters.offsets,
                          mpc.cell_to_slaves.array, mpc.cell_to_slaves.offsets, mpc.slaves, mpc.num_local_slaves)
    rp, col, val = oracle.assemble_matrix(P["a"], m, bcs=P["bcs"])
    b = oracle.assemble_vector(P["L"], m)
    oracle.apply_lifting(b, [P["a"]], [P["bcs"]], m)
    N = P["V"].num_dofs
    _compare(sp.csr_matrix((val, col, rp), shape=(N, N)), b, P)


@pytest.mark.gpu
@pytest.mark.skipif(not DUMPS, reason="no reference dumps under tests/replay (needs a DOLFINx box: tools/dump_reference.py)")
@pytest.mark.parametrize("path", DUMPS)
def test_cuda_path_reproduces_reference_dump(path):
    import dolfinx_mpc_b200 as mpcx

    P = load_dump(path)
    A, b = mpcx.assemble_system(P["a"], P["L"], P["mpc"], bcs=P["bcs"])
    _compare(A.to_scipy(), b.array, P)


def test_dump_layout_roundtrip(oracle, tmp_path):
    """A dump in the layout of tools/dump_reference.py written from this repository's own generators loads back into
    the same packed constraint and the oracle reproduces the stored system: the replay path works end to end, so a
    real dump only has to be dropped into tests/replay/."""
    import bench

    P = bench.build_problem(6)
    V, mpc = P["V"], P["mpc"]
    m = oracle.mpc_from_arrays(V, P["data"])
    rp, col, val = oracle.assemble_matrix(P["a"], m, bcs=P["bcs"])
    b = oracle.assemble_vector(P["L"], m)
    oracle.apply_lifting(b, [P["a"]], [P["bcs"]], m)
    keep = val != 0.0  # PETSc drops nothing that was inserted, but may hold fewer entries than the pattern
    A = sp.csr_matrix((val, col, rp), shape=(V.num_dofs, V.num_dofs))
    path = tmp_path / "synthetic_n6.npz"
    np.savez_compressed(
        path, x=P["mesh"].x, x_dofmap=P["mesh"].x_dofmap, cell_type="tetrahedron", degree=1, dofmap=V.dofmap, bs=1,
        dof_coordinates=V.tabulate_dof_coordinates(), mpc_slaves=mpc.slaves, mpc_masters=mpc.masters.array,
        mpc_offsets=mpc.masters.offsets, mpc_coeffs=mpc.coefficients()[0], is_slave=mpc.is_slave,
        c2s=mpc.cell_to_slaves.array, c2s_offsets=mpc.cell_to_slaves.offsets, num_local_slaves=mpc.num_local_slaves,
        bc_dofs=P["bcs"][0].dofs, bc_value=np.array(0.25), kernel="laplace", constants=np.array([1.0]), f=P["f"].array,
        A_indptr=A.indptr, A_indices=A.indices, A_data=A.data, b=b, versions="synthetic")
    assert keep.any()
    Q = load_dump(str(path))
    mq = Q["mpc"]
    m2 = oracle.WrappedMPC(Q["V"], mq.is_slave, mq.masters.array, mq.coefficients()[0], mq.masters.offsets,
                           mq.cell_to_slaves.array, mq.cell_to_slaves.offsets, mq.slaves, mq.num_local_slaves)
    rp2, col2, val2 = oracle.assemble_matrix(Q["a"], m2, bcs=Q["bcs"])
    b2 = oracle.assemble_vector(Q["L"], m2)
    oracle.apply_lifting(b2, [Q["a"]], [Q["bcs"]], m2)
    _compare(sp.csr_matrix((val2, col2, rp2), shape=A.shape), b2, Q)
