#!/usr/bin/env python
"""Times the fused matrix + vector bulk kernel alone (CUDA events inside libmpcx) -- the short command ncu wraps.
python tools/probe_system.py --n 128 [--config 2|4] [--unfused]"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import dolfinx_mpc_b200 as mpcx
from dolfinx_mpc_b200 import _lib, device as _dev

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=128)
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--unfused", action="store_true")
ap.add_argument("--async-zero", action="store_true")
args = ap.parse_args()
P = bench.build_config(args.config, args.n)
a, L, mpc, bcs = P["a"], P["L"], P["mpc"], P["bcs"]
P["f"].device_array = _dev.to_dev(P["f"].array)
A = mpcx.create_matrix(a, mpc)
b = mpcx.create_vector(mpc)
A.async_zero = args.async_zero
lib = _lib.load()


def step():
    if args.unfused:
        mpcx.assemble_matrix(a, mpc, bcs=bcs, A=A)
        mpcx.assemble_vector(L, mpc, b=b)
        mpcx.apply_lifting(b, [a], [bcs], mpc)
    else:
        mpcx.assemble_system(a, L, mpc, bcs=bcs, A=A, b=b)


for _ in range(2):
    step()
torch.cuda.synchronize()
lib.mpcx_profile_enable(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    step()
e1.record()
torch.cuda.synchronize()
ms, n = C.c_double(0), C.c_longlong(0)
lib.mpcx_profile_read(C.byref(ms), C.byref(n))
nc = P["mesh"].num_cells_local
info = [i for ent in A._tile_plans.values() if ent is not None for i in [ent[1]]]
print(f"lib={os.path.basename(_lib.LIB_PATH)} n={args.n} cells={nc} fused={getattr(A, 'last_system_fused', False)} "
      f"bulk kernel {ms.value / max(1, n.value):.3f} ms, step {e0.elapsed_time(e1) / args.reps:.3f} ms "
      f"-> {nc / (e0.elapsed_time(e1) / args.reps * 1e-3) / 1e9:.2f} Gcells/s  plans={info}")
