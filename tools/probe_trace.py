#!/usr/bin/env python
"""Phase timeline of the fused tile kernel from a -DMPCX_TRACE build (make -C dolfinx_mpc_b200/csrc trace):
MPCX_LIB=dolfinx_mpc_b200/libmpcx_trace.so python tools/probe_trace.py --n 128"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
import dolfinx_mpc_b200 as mpcx
from dolfinx_mpc_b200 import _lib, device as _dev

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=128)
ap.add_argument("--config", type=int, default=2)
args = ap.parse_args()
P = bench.build_config(args.config, args.n)
a, L, mpc, bcs = P["a"], P["L"], P["mpc"], P["bcs"]
P["f"].device_array = _dev.to_dev(P["f"].array)
A = mpcx.create_matrix(a, mpc)
b = mpcx.create_vector(mpc)
lib = _lib.load()
for _ in range(3):
    mpcx.assemble_system(a, L, mpc, bcs=bcs, A=A, b=b)
info = [i for ent in A._tile_plans.values() if ent is not None for i in [ent[1]]][0]
W = info["cells_per_tile"] // 32
buf = torch.zeros(8 * 64 * W * 12, dtype=torch.int64, device="cuda")
lib.mpcx_debug_set_trace.argtypes = [C.c_void_p]
_lib.check(lib.mpcx_debug_set_trace(buf.data_ptr()))
mpcx.assemble_system(a, L, mpc, bcs=bcs, A=A, b=b)
torch.cuda.synchronize()
_lib.check(lib.mpcx_debug_set_trace(None))
t = buf.cpu().numpy().reshape(8, 64, W, 12)[:, 4:60]  # skip the first iterations
names = ["top->waitC", "waitC", "phase1", "wait_group(w0)", "sync1", "zero+sync1b", "waitR", "matrix records", "row records+Xs",
         "sync2", "issue runs+TMA R"]
print(f"cells/tile {info['cells_per_tile']}, warps {W}; mean cycles per segment, per warp class (warp 0 / middle / last), "
      f"and the spread over warps")
per_iter = (t[:, 1:, 0, 0] - t[:, :-1, 0, 0]).mean()
print(f"iteration period (warp 0): {per_iter:.0f} cycles")
for k, nm in enumerate(names):
    seg = (t[..., k + 1] - t[..., k]) if k < 10 else None
    if k == 10:
        break
    print(f"  {nm:20s} w0 {seg[:, :, 0].mean():7.0f}  mid {seg[:, :, W // 2].mean():7.0f}  last {seg[:, :, W - 1].mean():7.0f}  "
          f"all {seg.mean():7.0f}  max-over-warps {seg.max(axis=2).mean():7.0f}")
tail = t[:, 1:, :, 0] - t[:, :-1, :, 10]
print(f"  {'loop tail -> top':20s} all {tail.mean():7.0f}")
arrive1 = t[..., 2]  # end of phase 1 per warp
print(f"  skew of phase-1 completion over warps: {(arrive1.max(axis=2) - arrive1.min(axis=2)).mean():.0f} cycles")
arrive2 = t[..., 8]
print(f"  skew of phase-2 completion over warps: {(arrive2.max(axis=2) - arrive2.min(axis=2)).mean():.0f} cycles")
