#!/usr/bin/env python
"""Writes the golden fixtures tests/golden/*.npz.

The reference (jorgensd/dolfinx_mpc) cannot be imported or built in this image (needs DOLFINx, PETSc, MPI,
FFCx) and its repository holds no golden vectors, so these fixtures are produced by the CPU oracle
(oracle/mpc_oracle.c) AFTER it passed the reference's own test identities (tests/test_oracle.py).  They pin
the oracle and the CUDA path against silent drift: inputs (mesh, dofmap, constraint arrays, bc dofs) and outputs
(CSR, RHS after lifting) are stored, so a box with DOLFINx could replay the same arrays through the reference.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import problems  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CASES = ["general2d-triangle-P1-5x3-m(1, 1)", "general2d-triangle-P1-1x8-m(0, 1)", "periodic2d-P1-8-bc1",
         "periodic3d-P1-bs1-4-ax2-bc1", "slip3d-P2-2", "contact3d", "tie2d-bs2", "lifting-quad", "varcoef-subdomains",
         "surface-traction2d", "surface-robin3d-P1"]


def main():
    orc.build()
    only = sys.argv[1:]  # optional: regenerate just these cases
    for name in (only or CASES):
        c = problems.ALL_CASES[name]()
        m = orc.mpc_from_arrays(c.V, c.data)
        rp, col, val = orc.assemble_matrix(c.a, m, bcs=c.bcs)
        out = dict(x=c.V.mesh.x, x_dofmap=c.V.mesh.x_dofmap, dofmap=c.V.dofmap, bs=c.V.bs, slaves=c.data[0],
                   masters=c.data[1], coeffs=c.data[2], owners=c.data[3], offsets=c.data[4],
                   bc_dofs=np.concatenate([bc.dofs for bc in c.bcs]) if c.bcs else np.zeros(0, np.int32),
                   row_ptr=rp, col=col, val=val)
        if c.L is not None:
            b = orc.assemble_vector(c.L, m)
            if c.a_lift is not None and c.bcs:
                orc.apply_lifting(b, [c.a_lift], [c.bcs], m)
            out["b"] = b
        fn = os.path.join(HERE, name.replace(" ", "").replace("(", "").replace(")", "").replace(",", "_") + ".npz")
        np.savez_compressed(fn, **out)
        print(fn, os.path.getsize(fn))


if __name__ == "__main__":
    main()
