#!/usr/bin/env python
"""Dump inputs AND outputs of the REFERENCE (jorgensd/dolfinx_mpc) for replay through this repository's C ABI.

Run this where DOLFINx + dolfinx_mpc are installed (it cannot run in the build image of this repository, which
has neither -- SURVEY.md section 8c), in SERIAL:

    python tools/dump_reference.py --out tests/replay --n 8 [--case periodic3d|periodic2d|elasticity_slip]

Every ``<case>.npz`` holds exactly the arrays the hot path reads (the array-level ABI the reference's numba path
states, python/src/dolfinx_mpc/numba/assemble_matrix.py:58-104) and what the reference computed from them:
  x [nn,3], x_dofmap [nc,ng], cell_type, degree, dofmap [nc,nd], bs,
  mpc_slaves, mpc_masters (local, dof-indexed adjacency array), mpc_coeffs, mpc_offsets (ndofs+1), is_slave,
  c2s, c2s_offsets, num_local_slaves, bc_dofs, bc_value, kernel ("laplace"|"elasticity"), constants,
  f (nodal values of the source coefficient), A_indptr / A_indices / A_data (A.getValuesCSR() of the reference's
  assemble_matrix), b (assemble_vector + apply_lifting, before set_bc).
``tests/test_replay.py`` feeds the inputs through the oracle (CPU) and through libmpcx (GPU) and compares with
A_* / b: that turns "parity unpinned" into a comparison against the reference binary.
"""
import argparse
import os

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="tests/replay")
    ap.add_argument("--n", type=int, default=8)
    ap.add_argument("--case", default="periodic3d", choices=["periodic3d", "periodic2d", "elasticity_slip"])
    args = ap.parse_args()

    from mpi4py import MPI
    from petsc4py import PETSc

    import dolfinx
    import dolfinx_mpc
    import ufl
    from dolfinx import default_scalar_type, fem, mesh as dmesh

    assert MPI.COMM_WORLD.size == 1, "dump in serial: the replay compares local == global numbering"
    n = args.n
    if args.case == "periodic2d":
        msh = dmesh.create_unit_square(MPI.COMM_WORLD, n, n, dmesh.CellType.triangle)
    else:
        msh = dmesh.create_unit_cube(MPI.COMM_WORLD, n, n, n, dmesh.CellType.tetrahedron)
    tdim = msh.topology.dim
    vector = args.case == "elasticity_slip"
    degree = 2 if vector else 1
    V = fem.functionspace(msh, ("Lagrange", degree, (tdim,)) if vector else ("Lagrange", degree))
    bs = V.dofmap.index_map_bs
    u, v = ufl.TrialFunction(V), ufl.TestFunction(V)
    x = ufl.SpatialCoordinate(msh)

    if vector:
        E, nu = 1.0e4, 0.1
        mu, lmbda = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))

        def sigma(w):
            return 2.0 * mu * ufl.sym(ufl.grad(w)) + lmbda * ufl.tr(ufl.sym(ufl.grad(w))) * ufl.Identity(tdim)

        a = ufl.inner(sigma(u), ufl.grad(v)) * ufl.dx  # python/benchmarks/bench_elasticity_edge.py:125-135
        constants = [mu, lmbda]
        kernel = "elasticity"
        fixed = fem.locate_dofs_geometrical(V, lambda X: np.isclose(X[0], 0.0))
        bc_value = np.array([0.01, -0.02, 0.03], dtype=default_scalar_type)
        bc = fem.dirichletbc(bc_value, fixed, V)
        fexpr = lambda X: np.stack([X[0] * np.sin(5 * np.pi * X[1]), 0.1 + 0 * X[0], -0.2 * X[2]])  # noqa: E731
    else:
        a = ufl.inner(ufl.grad(u), ufl.grad(v)) * ufl.dx
        constants = [1.0]
        kernel = "laplace"
        last = tdim - 1
        fixed = fem.locate_dofs_geometrical(V, lambda X: np.isclose(X[last], 0) | np.isclose(X[last], 1))
        bc_value = np.array(0.25, dtype=default_scalar_type)
        bc = fem.dirichletbc(default_scalar_type(0.25), fixed, V)
        fexpr = lambda X: X[0] * np.sin(5 * np.pi * X[1]) + np.exp(-((X[0] - 0.5) ** 2 + (X[1] - 0.5) ** 2) / 0.02)  # noqa: E731
    bcs = [bc]
    f = fem.Function(V)
    f.interpolate(fexpr)
    L = ufl.inner(f, v) * ufl.dx

    mpc = dolfinx_mpc.MultiPointConstraint(V)
    if vector:
        # slip u.n = 0 on x = 1 (cpp/SlipConstraint.h:123-140)
        facets = dmesh.locate_entities_boundary(msh, tdim - 1, lambda X: np.isclose(X[0], 1.0))
        mt = dmesh.meshtags(msh, tdim - 1, np.sort(facets), np.full(len(facets), 1, dtype=np.int32))
        nh = dolfinx_mpc.utils.create_normal_approximation(V, mt, 1)
        mpc.create_slip_constraint(V, (mt, 1), nh, bcs=bcs)
    else:
        def indicator(X):
            out = np.isclose(X[0], 1)
            if tdim == 3:
                out |= np.isclose(X[1], 1)
            return out

        def relation(X):  # python/tests/test_stokes_channelflow.py:48-55 style combined map
            out = X.copy()
            out[0] = np.where(np.isclose(X[0], 1), 0.0, X[0])
            if tdim == 3:
                out[1] = np.where(np.isclose(X[1], 1), 0.0, X[1])
            return out

        mpc.create_periodic_constraint_geometrical(V, indicator, relation, bcs)
    mpc.finalize()

    af, Lf = fem.form(a), fem.form(L)
    A = dolfinx_mpc.assemble_matrix(af, mpc, bcs=bcs)
    b = dolfinx_mpc.assemble_vector(Lf, mpc)
    dolfinx_mpc.apply_lifting(b, [af], [bcs], mpc)
    b.ghostUpdate(addv=PETSc.InsertMode.ADD_VALUES, mode=PETSc.ScatterMode.REVERSE)
    indptr, indices, data = A.getValuesCSR()

    Vm = mpc.function_space
    nc = msh.topology.index_map(tdim).size_local
    masters = mpc.masters
    c2s = mpc.cell_to_slaves
    os.makedirs(args.out, exist_ok=True)
    np.savez_compressed(
        os.path.join(args.out, f"{args.case}_n{n}.npz"),
        x=msh.geometry.x, x_dofmap=msh.geometry.dofmap[:nc].astype(np.int32),
        cell_type=msh.topology.cell_name(), degree=degree, dofmap=Vm.dofmap.list[:nc].astype(np.int32), bs=bs,
        dof_coordinates=Vm.tabulate_dof_coordinates(),
        mpc_slaves=mpc.slaves, mpc_masters=masters.array, mpc_offsets=masters.offsets,
        mpc_coeffs=mpc.coefficients()[0], is_slave=mpc.is_slave, c2s=c2s.array, c2s_offsets=c2s.offsets,
        num_local_slaves=mpc.num_local_slaves,
        bc_dofs=bc.dof_indices()[0].astype(np.int32), bc_value=bc_value, kernel=kernel, constants=np.array(constants),
        f=f.x.array, A_indptr=indptr, A_indices=indices, A_data=data, b=b.array,
        versions=f"dolfinx {dolfinx.__version__} dolfinx_mpc {dolfinx_mpc.__version__}")
    print("wrote", os.path.join(args.out, f"{args.case}_n{n}.npz"))


if __name__ == "__main__":
    main()
