#!/usr/bin/env python
"""bench.py -- cells assembled/sec (matrix + vector) for MPC-constrained assembly on B200.

One "step" = what LinearProblem.solve does before the linear solve (python/src/dolfinx_mpc/problem.py:539-582):
zero + assemble_matrix (bulk cells, slave-cell elimination, slave and Dirichlet diagonals), zero + assemble_vector,
apply_lifting, ghost-row / ghost-entry reduction when the mesh is partitioned.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config 2|3|4|5] [--n SIZE]

Workloads (BASELINE.json configs[i-1]):
  --config 2  3D unit-cube Poisson P1, 256^3 dofs, periodic x/y MPC, Dirichlet z-faces, one GPU   (default at N = 1)
  --config 3  3D elasticity P2 (bs = 3) on a rotated cube, slip constraint on the inclined face, ~50 M dofs, one GPU
  --config 4  3D Poisson P1, 512 x 512 nodes in x/y, periodic on ALL faces, pre-partitioned in z-slabs of 64 cube
              layers per GPU, ghost rows reduced over NCCL; 8 GPUs = the 512^3 problem          (default at N > 1)
  --config 5  contact constraint between two stacked boxes with non-matching grids, P1 elasticity (bs = 3),
              ~20 M dofs, up to 11 masters per slave, one GPU
N > 1 is launched by torchrun (one rank per GPU); per-GPU work is fixed (weak scaling).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
BULK_REDUCE_PEAK_GADDS = 270.0  # cp.reduce.async.bulk add.f64, whole B200, kernel-shaped operations (see main)
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "traffic.json")  # written by tools/ncu_summary.py from ncu --set full captures

DEFAULT_N = {2: 256, 3: 127, 4: 512, 5: 114}
NZC_CFG4 = 64  # cube layers per GPU of config 4 (8 x 64 = 512 layers = the 512^3 problem)


def f_source(x):
    # python/benchmarks/bench_periodic.py:85-91
    return x[0] * np.sin(5 * np.pi * x[1]) + np.exp(-((x[0] - 0.5) ** 2 + (x[1] - 0.5) ** 2) / 0.02)


def f_vector(x):
    f = f_source(x)
    return np.stack([f, -0.5 * f + 0.1, 0.25 * f - 0.2])


# ------------------------------------------------------------------------------------------------ workloads

def _finish(V, data, bcs, a, L, f, mesh, **extra):
    from dolfinx_mpc_b200 import MultiPointConstraint

    mpc = MultiPointConstraint(V)
    mpc.add_constraint(V, *data)
    mpc.finalize()
    return dict(mesh=mesh, V=V, bcs=bcs, data=data, mpc=mpc, a=a, L=L, f=f, **extra)


def build_problem(n: int, nz: int | None = None):
    """Config 2: unit-cube P1 Poisson with n^2 x (nz or n) nodes, periodic x/y, Dirichlet z-faces (host arrays)."""
    from dolfinx_mpc_b200 import fem, generators as gen

    nz = n if nz is None else nz
    mesh = gen.create_box(n - 1, n - 1, nz - 1, p1=(1.0, 1.0, (nz - 1) / (n - 1)))
    V = gen.functionspace(mesh, 1)
    zmax = mesh.x[:, 2].max()
    bc_dofs = fem.locate_dofs_geometrical(V, lambda x: np.isclose(x[2], 0) | np.isclose(x[2], zmax))
    bcs = [fem.DirichletBC(V, bc_dofs, 0.25)]
    data = gen.periodic_constraint(V, axes=(0, 1), exclude_dofs=bc_dofs)
    f = fem.Function(V)
    f.interpolate(f_source)
    return _finish(V, data, bcs, fem.laplace(V), fem.source(V, f), f, mesh)


def build_slip_elasticity(n: int, theta: float = np.pi / 5):
    """Config 3: P2 vector elasticity (E = 1e4, nu = 0.1: python/benchmarks/bench_elasticity_edge.py:116-135) on
    n^3 cubes x 6 tetrahedra, geometry rotated by theta about (1,1,0)/sqrt(2) (python/tests/test_cube_contact.py:28,45),
    slip u.n = 0 on the (rotated) face x = 1 (cpp/SlipConstraint.h:123-140), Dirichlet on the opposite face."""
    from dolfinx_mpc_b200 import fem, generators as gen

    mesh0 = gen.create_unit_cube(n, n, n)
    V0 = gen.functionspace(mesh0, 2, 3)
    X0 = V0.tabulate_dof_coordinates()
    slip_blocks = np.flatnonzero(np.isclose(X0[:, 0], 1.0))
    fixed = np.flatnonzero(np.isclose(X0[:, 0], 0.0))
    k = np.array([1.0, 1.0, 0.0]) / np.sqrt(2)
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    R = np.eye(3) + np.sin(theta) * Kx + (1 - np.cos(theta)) * (Kx @ Kx)
    mesh = fem.Mesh(mesh0.x @ R.T, mesh0.x_dofmap, mesh0.cell_type)
    V = fem.FunctionSpace(mesh, 2, V0.dofmap, 3, V0.index_map, X0 @ R.T)
    del mesh0, V0
    normal = R @ np.array([1.0, 0.0, 0.0])
    bc_dofs = (fixed[:, None] * 3 + np.arange(3)[None, :]).reshape(-1).astype(np.int32)
    bcs = [fem.DirichletBC(V, bc_dofs, np.array([0.01, -0.02, 0.03]))]
    data = gen.slip_constraint(V, slip_blocks, normal, exclude_dofs=bc_dofs)
    E, nu = 1.0e4, 0.1
    mu, lmbda = E / (2 * (1 + nu)), E * nu / ((1 + nu) * (1 - 2 * nu))
    f = fem.Function(V)
    f.interpolate(f_vector)
    return _finish(V, data, bcs, fem.elasticity(V, mu, lmbda), fem.source(V, f), f, mesh)


def build_contact(n: int):
    """Config 5: two stacked boxes [0,1]^2 x [0,0.5] and [0,1]^2 x [0.5,1] meshed with n and 2n cells per unit
    length (non-matching interface, python/tests/test_cube_contact.py:31-45,145-147), P1 vector elasticity
    (E = 1e3, nu = 0), contact-slip on the interface: every upper interface block is a slave with the other
    components of its block and the three vertices of the facing lower facet (x 3 components) as masters
    (cpp/ContactConstraint.h:87-152); bottom face clamped."""
    from dolfinx_mpc_b200 import fem, generators as gen

    mesh = gen.create_stacked_boxes((n, n, max(1, n // 2)), (2 * n, 2 * n, n))
    V = gen.functionspace(mesh, 1, 3)
    X = V.tabulate_dof_coordinates()
    bottom = np.flatnonzero(np.isclose(X[:, 2], 0.0))
    bc_dofs = (bottom[:, None] * 3 + np.arange(3)[None, :]).reshape(-1).astype(np.int32)
    bcs = [fem.DirichletBC(V, bc_dofs, 0.0)]
    nrm = np.array([0.2, -0.1, 1.0])
    data = gen.contact_constraint(V, 0.5, normal=nrm / np.linalg.norm(nrm))
    f = fem.Function(V)
    f.interpolate(f_vector)
    return _finish(V, data, bcs, fem.elasticity(V, 1.0e3 / 2, 0.0), fem.source(V, f), f, mesh)


def build_config(cfg: int, n: int, rank: int = 0, world: int = 1):
    if cfg == 2:
        if world > 1:  # config 2 stacked in z: per-GPU work fixed, Dirichlet on the two end planes
            from dolfinx_mpc_b200 import distributed

            return distributed.build_slab_problem(n, rank, world, f_source)
        return build_problem(n)
    if cfg == 4:
        from dolfinx_mpc_b200 import distributed

        nzc = NZC_CFG4 if n == DEFAULT_N[4] else max(2, n // 8)
        return distributed.build_slab_problem(n, rank, world, f_source, periodic_z=True, nzc=nzc)
    if world > 1:
        raise SystemExit(f"--config {cfg} is a single-GPU workload here (its mesh is not pre-partitioned); use --gpus 1")
    return build_slip_elasticity(n) if cfg == 3 else build_contact(n)


def workload_config(cfg, n, gpus):
    desc = {
        2: f"3D unit-cube Poisson P1 (Kuhn tets), {n}^3 dofs per GPU, periodic x/y MPC, Dirichlet z-faces, fp64"
           " (BASELINE.json configs[1])",
        3: f"3D linear elasticity P2 (bs 3) on {n}^3 x 6 tets, rotated cube, slip constraint on the inclined face, "
           f"{3 * (2 * n + 1) ** 3} dofs, fp64 (BASELINE.json configs[2])",
        4: f"3D Poisson P1 (Kuhn tets), {n} x {n} x ({gpus} x {NZC_CFG4 if n == DEFAULT_N[4] else max(2, n // 8)} + 1) nodes, "
           f"periodic on all faces, pre-partitioned into {gpus} z-slab(s), ghost-row NCCL reduce, fp64 (BASELINE.json "
           "configs[3]: 8 GPUs = 512^3)",
        5: f"contact constraint between two stacked boxes ({n} / {2 * n} cells per unit length), P1 elasticity (bs 3), "
           "fp64 (BASELINE.json configs[4])",
    }[cfg]
    nzc = NZC_CFG4 if n == DEFAULT_N[4] else max(2, n // 8)
    cells, dofs = {2: (6 * (n - 1) ** 3, n ** 3), 3: (6 * n ** 3, 3 * (2 * n + 1) ** 3),
                   4: (6 * (n - 1) ** 2 * nzc, n * n * (nzc + 1)),
                   5: (6 * (n * n * max(1, n // 2) + 4 * n * n * n), 3 * ((n + 1) ** 2 * (max(1, n // 2) + 1) + (2 * n + 1) ** 2 * (n + 1)))}[cfg]
    # the same dictionary in both arms (ours / reference): it describes the workload, nothing about the implementation
    return {"workload": desc, "config_id": cfg, "n": n, "gpus": gpus, "cells_per_gpu": cells, "dofs_per_gpu": dofs,
            "l2": "inputs (> 3 GB per step) far larger than the 126 MB L2; no explicit flush",
            "partition": "z-slabs by ownership, ghost-row NCCL reduce" if gpus > 1 else "single GPU"}


def algorithmic_bytes(P, nnz: int, with_rhs: bool) -> float:
    """SURVEY.md section 8(d) compulsory-traffic formula, exact nnz of the built pattern."""
    V, mesh = P["V"], P["mesh"]
    nc = mesh.num_cells_local
    nd, ng = V.nd, mesh.x_dofmap.shape[1]
    ndofs = V.num_dofs
    per_cell = 4 * nd + 4 * ng + 4
    glob = 24 * mesh.x.shape[0] + 2 * ndofs + nnz * 12 + 8 * (ndofs + 1) + (8 * ndofs if with_rhs else 0)
    return per_cell * nc + glob


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe).  The sampler is
    started before the warm-up (nvidia-smi needs ~0.1 s to deliver its first line) and samples every 20 ms; only
    the samples whose timestamp falls inside the timed window [t0, t1] are used."""

    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self, t0: float, t1: float) -> dict:
        import datetime

        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        rows = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            p = [q.strip() for q in line.split(",")]
            if len(p) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(p[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(p[2]), float(p[3]), [n for n, v in zip(names, p[6:10]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        inside = [r for r in rows if t0 - 0.01 <= r[0] <= t1 + 0.01]
        used = inside or rows[-3:]
        reasons = sorted({n for r in used for n in r[3]})
        return {"sm_mhz": float(np.median([r[1] for r in used])) if used else None,
                "sm_max_mhz": max(r[2] for r in used) if used else None, "samples": len(inside),
                "samples_total": len(rows), "reasons": reasons}


# ------------------------------------------------------------------------------------------------ CPU reference arm

LAST_ONE_RANK = None


def _oracle_rank(P):
    """Oracle-side objects of one rank's problem: packed constraint (masters in the rank's extended local
    numbering, cpp/MultiPointConstraint.h:117-125) and the local sparsity pattern."""
    from oracle import oracle as orc

    mpc = P["mpc"]
    V = mpc.function_space
    m = orc.WrappedMPC(V, mpc.is_slave, mpc.masters.array, mpc.coefficients()[0], mpc.masters.offsets,
                       mpc.cell_to_slaves.array, mpc.cell_to_slaves.offsets, mpc.slaves, mpc.num_local_slaves)
    pat = orc.create_pattern(P["a"], m, m)
    return m, pat


def _divisor_at_most(k: int, limit: int) -> int:
    return max(d for d in range(1, max(1, limit) + 1) if k % d == 0)


def cpu_rank_problems(cfg: int, n: int, cores: int, budget_cells: float):
    """The per-rank problems of the CPU arm: R ranks, EACH WITH ITS OWN ARRAYS (mesh, dofmap, constraint, pattern,
    matrix, vector), as `mpirun -n R` gives the reference (README.md:30 of the reference: owner-local cells, no
    communication during assembly).

    Configs 2 and 4 (slab-partitionable P1 Poisson): the ranks' slabs together cover one GPU's whole workload when
    `budget_cells` allows it (R = largest divisor of the number of cube layers <= cores); otherwise thinner slabs of
    the same cross-section (a bounded sample).  Configs 3 and 5: R independent smaller instances of the same
    problem (same element, constraint type and slaves-per-cell density), sized by the budget."""
    from dolfinx_mpc_b200 import distributed

    if cfg in (2, 4):
        layers = (n - 1) if cfg == 2 else (NZC_CFG4 if n == DEFAULT_N[4] else max(2, n // 8))
        per_layer = 6 * (n - 1) ** 2
        R = _divisor_at_most(layers, cores)
        nzc = layers // R
        full = True
        if R * nzc * per_layer > budget_cells:  # thinner slabs, all cores
            R, full = cores, False
            nzc = int(max(2, min(layers // R, budget_cells / (R * per_layer))))
        probs = [distributed.build_slab_problem(n, r, R, f_source, periodic_z=(cfg == 4), nzc=nzc) for r in range(R)]
        what = (f"{R} ranks x {n}x{n}x{nzc + 1}-node z-slabs, each rank its own arrays; "
                + ("together the whole single-GPU workload" if full else "a bounded sample of the workload's slabs"))
        return probs, what, full
    R = cores
    if cfg == 3:
        ns = int(max(2, min(n, round((budget_cells / (6 * R)) ** (1 / 3)))))
        probs = [build_slip_elasticity(ns) for _ in range(R)]
        what = f"{R} ranks x an independent {ns}^3-cube instance of the P2 slip-elasticity problem (bounded sample)"
    else:
        ns = int(max(2, min(n, round((budget_cells / (6 * 4.5 * R)) ** (1 / 3)))))
        probs = [build_contact(ns) for _ in range(R)]
        what = f"{R} ranks x an independent instance of the contact problem with {ns} / {2 * ns} cells per unit length (bounded sample)"
    return probs, what, False


CPU_RATE_GUESS = {2: 2.5e6, 3: 2.5e4, 4: 2.5e6, 5: 3.0e5}  # cells/s per rank, only used to size the sample


def cpu_reference(cfg: int, n: int, steps: int, warmup: int, budget_s: float, label: str):
    """The reference's algorithm (oracle port, oracle/mpc_oracle.c -- the reference itself cannot be built in
    this image) on the host cores: R independent ranks, each assembling its own partition into its own local
    matrix / vector with no communication, as `mpirun -n R` does; time of a step = slowest rank."""
    from oracle import oracle as orc

    # the CPU arm runs a -march=native build of the oracle; the tests keep the portable one
    orc.LIB_PATH = orc.build(march="native", out=os.path.join(os.path.dirname(orc.LIB_PATH), "libmpc_oracle_native.so"))
    cores = os.cpu_count() or 1
    per_step = budget_s / max(1, steps + warmup)
    probs, what, full = cpu_rank_problems(cfg, n, cores, CPU_RATE_GUESS[cfg] * cores * per_step)
    R = len(probs)
    setup = [None] * R

    def prepare(r):
        setup[r] = _oracle_rank(probs[r])

    th = [threading.Thread(target=prepare, args=(r,)) for r in range(R)]
    [t.start() for t in th]
    [t.join() for t in th]
    bs_ = [np.zeros(p["mpc"].function_space.num_dofs) for p in probs]
    cells = sum(p["mesh"].num_cells_local for p in probs)
    times = np.zeros(R)

    def work(r):
        P = probs[r]
        m, pat = setup[r]
        t = time.perf_counter()
        orc.assemble_matrix(P["a"], m, bcs=P["bcs"], pattern=pat)
        orc.assemble_vector(P["L"], m, bs_[r])
        if P["bcs"]:
            orc.apply_lifting(bs_[r], [P["a"]], [P["bcs"]], m)
        times[r] = time.perf_counter() - t

    def one_step():
        th = [threading.Thread(target=work, args=(r,)) for r in range(R)]
        t = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        return time.perf_counter() - t

    for _ in range(warmup):
        one_step()
    work(0)  # one rank alone (BASELINE.json configs[0] is quoted on 1 CPU rank): no neighbours on the cores
    global LAST_ONE_RANK
    LAST_ONE_RANK = probs[0]["mesh"].num_cells_local / times[0]
    dts = [one_step() for _ in range(steps)]
    total = float(sum(dts))
    value = cells * steps / total
    sample = (f"{label}: {what} ({cells} cells per step in total), GIL released in C, pattern cached, "
              f"zero+matrix+vector+lifting timed, oracle built with -march=native")
    return value, total / steps * 1e3, R, sample, cells, full


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, ms, R, sample, cells, full = cpu_reference(args.config, args.n, args.steps, args.warmup, args.ref_budget,
                                                      "reference arm")
    line = {
        "impl": "reference", "metric": "cells assembled/sec (matrix+vector)", "value": value, "unit": "cells/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.config, args.n, args.gpus),
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": R, "kind": "port", "sample": sample,
                         "one_rank_value": LAST_ONE_RANK, "whole_single_gpu_workload": full},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm

def bind_to_gpu_numa_node(local_rank: int):
    """Pin this rank's host threads (and with them the first-touch placement of its pinned staging buffers) to the
    NUMA node its GPU hangs off: with 8 ranks downloading 2 GB per step each, buffers placed on the far socket were
    what capped the end-to-end rate of round 1 near 88 GB/s in aggregate.  Best effort; returns what it did."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return {"numa_node": None}
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed)}
    except Exception as e:  # no sysfs entry, containerised cpuset, ...
        return {"numa_node": None, "why": type(e).__name__}


def run_ours(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as ge
    import dolfinx_mpc_b200 as mpcx
    from dolfinx_mpc_b200 import _lib, device as dev

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        ge.build()
    if world > 1:
        dist.barrier()
    lib = _lib.load()

    cfg, n = args.config, args.n
    P = build_config(cfg, n, rank, world)
    mesh, V, mpc, a, L, bcs, f = (P[k] for k in ("mesh", "V", "mpc", "a", "L", "bcs", "f"))
    nc = mesh.num_cells_local
    f.device_array = dev.to_dev(f.array)
    if world > 1:
        from dolfinx_mpc_b200 import distributed

        A = distributed.create_matrix(a, mpc)
        b = mpcx.create_vector(mpc)
        distributed.attach_ghost_exchange(A, b, P)
    else:
        A = mpcx.create_matrix(a, mpc)
        b = mpcx.create_vector(mpc)

    A.async_zero = bool(args.async_zero) and not args.unfused  # zero-fill of A overlapped with the previous assembly

    A.deferred_errors = True  # cached, validated plans: the per-step error flag travels asynchronously (la.Matrix)

    def step():
        if args.unfused:
            mpcx.assemble_matrix(a, mpc, bcs=bcs, A=A)
            mpcx.assemble_vector(L, mpc, b=b)
            mpcx.apply_lifting(b, [a], [bcs], mpc)
            b.ghostUpdate()
        else:
            mpcx.assemble_system(a, L, mpc, bcs=bcs, A=A, b=b)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        step()
    barrier()
    lib.mpcx_profile_enable(1)
    launches0 = lib.mpcx_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    t_wall1 = time.time()
    A.synchronize()  # raises if any step of the loop reported a device error
    clk = clocks.stop(t_wall0, t_wall1) if rank == 0 else None  # samples inside the timed window, 20 ms apart
    ms_total = ev0.elapsed_time(ev1)
    launches = lib.mpcx_launch_count() - launches0
    kms, kn = C.c_double(0), C.c_longlong(0)
    _lib.check(lib.mpcx_profile_read(C.byref(kms), C.byref(kn)))
    lib.mpcx_profile_enable(0)
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        cells_t = torch.tensor([nc], device="cuda", dtype=torch.int64)
        dist.all_reduce(cells_t)
        total_cells = int(cells_t.item())
    else:
        total_cells = nc
    value = total_cells * args.steps / (ms_total * 1e-3)

    # roofline of the dominant kernel (the bulk-cell kernel bracketed by CUDA events inside libmpcx), rank 0
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        try:
            peak = float(json.load(open(peaks_path))["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    k_ms = kms.value / max(1, kn.value)
    fused = not args.unfused and getattr(A, "last_system_fused", False)
    alg = algorithmic_bytes(P, A.nnz, with_rhs=fused)
    achieved = alg / (k_ms * 1e-3) / 1e9
    traffic, traffic_src = args.traffic, "command line"
    if traffic is None and os.path.exists(TRAFFIC_FILE):
        try:
            ent = json.load(open(TRAFFIC_FILE)).get(f"cfg{cfg}_n{n}")
            if ent:
                traffic, traffic_src = float(ent["dram_bytes_per_launch"]), ent["source"]
        except Exception:
            pass
    roofline = {"bound": "hbm", "kernel": "fused matrix + vector bulk-cell kernel" if fused else "matrix bulk-cell kernel",
                "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
                "traffic_source": traffic_src if traffic is not None else None,
                "algorithmic_bytes_per_launch": alg, "bytes_per_cell": alg / nc, "kernel_ms": k_ms,
                "kernel_share_of_step": kms.value / ms_total,
                "step_algorithmic_frac": algorithmic_bytes(P, A.nnz, True) / (ms_total / args.steps * 1e-3) / 1e9 / peak}

    # second bound of the tile kernels: the fp64 additions the copy engine performs at L2 (cp.reduce.async.bulk).  Its
    # peak was measured in isolation with the kernel's operation shape (tools/micro/bulk_rate.cu,
    # profiles/r02_n_bulk_rate.txt: 33 operations of 672 B per tile and CTA, two CTAs per SM -> 270 G fp64 adds/s)
    slots = [e[1].get("stage_slots", 0) for e in getattr(A, "_tile_plans", {}).values() if e is not None and not e[1].get("vec")]
    if fused and slots and slots[0] > 0:
        roofline["reduce_engine"] = {"fp64_adds_per_launch": int(slots[0]), "achieved": slots[0] / (k_ms * 1e-3) / 1e9,
                                     "peak": BULK_REDUCE_PEAK_GADDS, "unit": "G fp64 adds/s",
                                     "frac": slots[0] / (k_ms * 1e-3) / 1e9 / BULK_REDUCE_PEAK_GADDS,
                                     "peak_source": "measured in isolation, tools/micro/bulk_rate.cu (profiles/r02_n_bulk_rate.txt)"}

    # matrix-only / vector-only split of the step, timed separately after the main loop
    def timed(fn, reps=3):
        fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / reps

    breakdown = {"assemble_matrix_ms": timed(lambda: mpcx.assemble_matrix(a, mpc, bcs=bcs, A=A)),
                 "assemble_vector_ms": timed(lambda: mpcx.assemble_vector(L, mpc, b=b)),
                 "apply_lifting_ms": timed(lambda: mpcx.apply_lifting(b, [a], [bcs], mpc))}

    # end to end through the public API with host (pinned) buffers (see run_e2e)
    e2e = run_e2e(args, P, A, b, step, world, barrier)
    # same pipeline when the assembled system is consumed on the device (Matrix.dlpack / to_torch_sparse_csr) and
    # only two norms travel back: reported beside the headline e2e, not instead of it
    e2e_dev = run_e2e(args, P, A, b, step, world, barrier, download="norms")
    # symmetric systems (one constraint, one space): the upper triangle + RHS is all a host solver needs -- half the bytes
    e2e_upper = run_e2e(args, P, A, b, step, world, barrier, download="upper") if A.nnz < (1 << 31) else None

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, ms, R, sample, _, full = cpu_reference(cfg, n, 1, 1, args.cpu_budget, "cpu_baseline")
        cpu = {"value": v, "unit": "cells/s", "cores": R, "kind": "port", "sample": sample,
               "one_rank_value": LAST_ONE_RANK, "whole_single_gpu_workload": full}

    plans = [info for ent in A._tile_plans.values() if ent is not None for info in [ent[1]]]
    for it_ in L.integrals:
        plans += [v[2] for k_, v in it_._dev.items() if isinstance(k_, tuple) and k_[0] == "vector_tile_plan" and v is not None]
    if rank == 0:
        line = {
            "metric": "cells assembled/sec (matrix+vector)", "value": value, "unit": "cells/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(cfg, n, world),
            "detail": dict(cells_rank0=nc, dofs_rank0=V.num_dofs, nnz_rank0=A.nnz, slaves_rank0=len(mpc.slaves),
                           step="fused assemble_system" if fused else "assemble_matrix + assemble_vector + apply_lifting",
                           zero_fill="second value buffer cleared on a side stream during the previous step" if A.async_zero else "on the assembly stream"),
            "roofline": roofline, "breakdown": breakdown, "cpu_baseline": cpu, "e2e": e2e,
            "e2e_device_consumer": e2e_dev, "e2e_upper_triangle": e2e_upper, "gpu_launches": int(launches),
            "clocks": clk, "numa": numa, "tile_plans": plans, "lib": os.path.basename(_lib.LIB_PATH),
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, P, A, b, step, world, barrier, download="full"):
    """End to end through the public API with HOST buffers (dolfinx_mpc_b200.hostio.StepIO).  Every step uploads
    the step's INPUT VALUES from pinned host memory (vertex coordinates, the coefficient, constraint coefficients;
    Dirichlet values travel inside apply_lifting when they can change) and downloads the assembled CSR values and the
    RHS to pinned host memory.  The mesh topology / dofmaps / constraint structure stay on the device together with
    the sparsity pattern and the tile plans derived from them -- the state the reference keeps in its Form /
    FunctionSpace / cached Mat between assemblies (python/src/dolfinx_mpc/assemble_matrix.py:49-51).  Steps are
    pipelined over three streams (upload of step k+1 and download of step k overlap; value arrays double-buffered
    when they fit); the time is K steps start to finish."""
    import torch

    from dolfinx_mpc_b200.hostio import StepIO

    io = StepIO(P["mesh"], [P["f"]], P["mpc"], A, b, download=download)
    cur = torch.cuda.current_stream()

    def run(k_steps):
        ev_done, ev_out = [], []
        for k in range(k_steps):
            i = k % io.nbuf
            ev_in = io.upload(after=ev_done[k - 1] if k > 0 else None)  # the previous step has read its inputs
            cur.wait_event(ev_in)
            if k >= io.nbuf:
                cur.wait_event(ev_out[k - io.nbuf])  # this set of output buffers has been downloaded
            io.bind(i)
            step()
            e = torch.cuda.Event()
            e.record(cur)
            ev_done.append(e)
            ev_out.append(io.download(i, after=e))
        for e in ev_out[-io.nbuf:]:
            cur.wait_event(e)

    run(2)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = max(2, min(args.steps, 8 if io.d2h_bytes < (4 << 30) else 3))
    ev0.record()
    run(k)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    io.release()
    nc = P["mesh"].num_cells_local
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        c = torch.tensor([nc], device="cuda", dtype=torch.int64)
        dist.all_reduce(c)
        nc = int(c.item())
    return {"value": nc * k / (ms * 1e-3), "unit": "cells/s", "h2d_bytes_per_step": int(io.h2d_bytes),
            "d2h_bytes_per_step": int(io.d2h_bytes), "ms_per_step": ms / k, "steps": k,
            "download": io.download_desc,
            "note": "per step: pinned-host upload of the input VALUES (vertex coordinates, coefficient, constraint "
                    "coefficients), assembly, download to pinned host memory; topology (dofmaps, constraint "
                    "structure), sparsity pattern and tile plans are cached on the device like the reference's Form / "
                    "cached Mat; steps pipelined over 3 streams"}


_REAL_STDOUT = None


def _quiet_stdout():
    """Route everything libraries print to fd 1 (NCCL's version banner, build output) to stderr, so that stdout
    carries exactly one line: the JSON result, written by emit()."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=None, choices=[2, 3, 4, 5],
                    help="BASELINE.json config (1-based); default 2 on one GPU, 4 on several")
    ap.add_argument("--n", type=int, default=None, help="size parameter of the config (nodes per side for 2 / 4, "
                    "cubes per side for 3, lower-box cells per unit length for 5)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="seconds of CPU work for --impl reference")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--unfused", action="store_true", help="step = three separate calls instead of assemble_system")
    ap.add_argument("--async-zero", action="store_true", help="clear a second value buffer on a side stream during the "
                    "previous step instead of zeroing A on the assembly stream (measured: no gain, profiles/README.md)")
    ap.add_argument("--traffic", type=float, default=None, help="dram bytes per launch of the dominant kernel "
                    "from an ncu --set full capture, reported as roofline.traffic (default: profiles/traffic.json)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.config is None:
        args.config = 2 if max(world, args.gpus) == 1 else 4
    if args.n is None:
        args.n = DEFAULT_N[args.config]
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        if args.steps > 20:  # the CPU arm's step is seconds, not milliseconds: keep the run within minutes
            args.steps = 20
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
