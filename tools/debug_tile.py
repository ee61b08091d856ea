"""Debug helper: one small matrix + vector assembly through the tile path, compared with the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import problems
import dolfinx_mpc_b200 as mpcx
from oracle import oracle as orc

orc.build()
name = sys.argv[1] if len(sys.argv) > 1 else "general2d-triangle-P1-1x8-m0_1"
c = problems.ALL_CASES[name]() if name in problems.ALL_CASES else None
if c is None:
    n = int(name)  # a cube size
    import bench
    P = bench.build_problem(n)
    class C_: pass
    c = C_(); c.V = P["V"]; c.a = P["a"]; c.L = P["L"]; c.bcs = P["bcs"]; c.data = P["data"]
mpc = mpcx.MultiPointConstraint(c.V); mpc.add_constraint(c.V, *c.data); mpc.finalize()
A = mpcx.assemble_matrix(c.a, mpc, bcs=c.bcs)
torch.cuda.synchronize()
print("matrix plans", [i for _, i in A._tile_plans.values()])
b = mpcx.assemble_vector(c.L, mpc)
torch.cuda.synchronize()
print("vector plans", [v[2] for k, v in c.L.integrals[0]._dev.items() if isinstance(k, tuple) and k[0] == "vector_tile_plan"])
m = orc.mpc_from_arrays(c.V, c.data)
rp, col, val = orc.assemble_matrix(c.a, m, bcs=c.bcs)
bo = orc.assemble_vector(c.L, m)
print("A err", np.abs(A.getValuesCSR()[2] - val).max(), "b err", np.abs(b.array - bo).max())
if np.abs(b.array - bo).max() > 1e-12:
    print("b  ", b.array[:24]); print("b_o", bo[:24])
