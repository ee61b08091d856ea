#!/usr/bin/env python
"""Times the bulk matrix kernel alone (CUDA events inside libmpcx) for one scatter strategy / tile size.
Used for kernel tuning and as the short command ncu wraps.  python tools/probe_matrix.py --n 128 --tile-cells 640"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import dolfinx_mpc_b200 as mpcx
from dolfinx_mpc_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=128)
ap.add_argument("--scatter", default="tile")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
P = bench.build_problem(args.n)
A = mpcx.create_matrix(P["a"], P["mpc"])
A.scatter = args.scatter
lib = _lib.load()
for _ in range(2):
    mpcx.assemble_matrix(P["a"], P["mpc"], bcs=P["bcs"], A=A)
torch.cuda.synchronize()
lib.mpcx_profile_enable(1)
for _ in range(args.reps):
    mpcx.assemble_matrix(P["a"], P["mpc"], bcs=P["bcs"], A=A)
ms, n = C.c_double(0), C.c_longlong(0)
lib.mpcx_profile_read(C.byref(ms), C.byref(n))
nc = P["mesh"].num_cells_local
from dolfinx_mpc_b200 import device as _dev
P["f"].device_array = _dev.to_dev(P["f"].array)
b = mpcx.create_vector(P["mpc"])
for _ in range(2):
    mpcx.assemble_vector(P["L"], P["mpc"], b=b)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    mpcx.assemble_vector(P["L"], P["mpc"], b=b)
e1.record()
torch.cuda.synchronize()
print(f"assemble_vector (zero + kernels): {e0.elapsed_time(e1) / args.reps:.3f} ms")
info = [i for _, i in A._tile_plans.values()]
print(f"n={args.n} scatter={args.scatter} cells={nc} bulk kernel {ms.value / n.value:.3f} ms "
      f"-> {nc / (ms.value / n.value * 1e-3) / 1e9:.2f} Gcells/s  plan={info}")
