#!/usr/bin/env python
"""Times the pieces of one distributed step (torchrun, one rank per GPU): assembly kernels vs ghost-row exchanges."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
import dolfinx_mpc_b200 as mpcx
from dolfinx_mpc_b200 import device as dev, distributed

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
P = distributed.build_slab_problem(n, rank, world, bench.f_source)
mesh, V, mpc, a, L, bcs, f = (P[k] for k in ("mesh", "V", "mpc", "a", "L", "bcs", "f"))
f.device_array = dev.to_dev(f.array)
A = distributed.create_matrix(a, mpc)
b = mpcx.create_vector(mpc)
distributed.attach_ghost_exchange(A, b, P)
ex = A.ghost_exchange
A.ghost_exchange = None  # time the exchange separately


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


t_mat = timed(lambda: mpcx.assemble_matrix(a, mpc, bcs=bcs, A=A))
t_exm = timed(lambda: ex.reduce_matrix(A))
t_vec = timed(lambda: (mpcx.assemble_vector(L, mpc, b=b), mpcx.apply_lifting(b, [a], [bcs], mpc)))
t_exv = timed(lambda: b.ghost_exchange.reduce_vector(b))
t_a2a = timed(lambda: dist.all_to_all_single(ex.mat.recv_buf, A.val[ex.mat.send_start:ex.mat.send_start + ex.mat.n_send],
                                             ex.mat.recv_counts, ex.mat.send_counts))
t_sc = timed(lambda: ex.mat._scatter_add(A.val, ex.mat.recv_pos, ex.mat.recv_buf))
print(f"rank {rank}: a2a alone {t_a2a:.3f} ms, scatter_add alone {t_sc:.3f} ms", flush=True)
print(f"rank {rank}: matrix {t_mat:.3f} ms, matrix exchange {t_exm:.3f} ms (send {ex.mat.n_send} recv {ex.mat.n_recv} "
      f"contiguous {ex.mat.contiguous}), vector+lifting {t_vec:.3f} ms, vector exchange {t_exv:.3f} ms", flush=True)
dist.destroy_process_group()
