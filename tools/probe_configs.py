#!/usr/bin/env python
"""Throughput of the parity-only BASELINE configs on the generic warp-per-cell kernels (not bench lines):
config 3 (P2 vector elasticity, slip constraint on an inclined boundary) and config 5 (P1 vector elasticity,
contact constraint between two boxes with non-matching grids), at sizes that build in seconds on the host.
    python tools/probe_configs.py [n_slip] [n_contact]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch

import dolfinx_mpc_b200 as mpcx
import problems

n_slip = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n_con = int(sys.argv[2]) if len(sys.argv) > 2 else 24
for name, make in (("config 3: P2 elasticity + slip", lambda: problems.case_slip_elasticity_3d(n_slip, 2)),
                   ("config 5: P1 elasticity + contact", lambda: problems.case_contact_3d((n_con, n_con, n_con // 2),
                                                                                         (n_con + 8, n_con + 8, n_con // 2)))):
    t0 = time.time()
    c = make()
    mpc = mpcx.MultiPointConstraint(c.V)
    mpc.add_constraint(c.V, *c.data)
    mpc.finalize()
    A = mpcx.create_matrix(c.a, mpc)
    t_setup = time.time() - t0
    for _ in range(2):
        mpcx.assemble_matrix(c.a, mpc, bcs=c.bcs, A=A)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5
    for _ in range(reps):
        mpcx.assemble_matrix(c.a, mpc, bcs=c.bcs, A=A)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nc = c.V.mesh.num_cells_local
    print(f"{name}: cells={nc} dofs={c.V.num_dofs} nnz={A.nnz} slaves={len(mpc.slaves)} slave_cells={len(mpc.slave_cells)} "
          f"assemble_matrix {ms:.3f} ms -> {nc / ms / 1e3:.2f} M cells/s (setup {t_setup:.1f} s)", flush=True)
