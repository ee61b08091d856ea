"""ctypes binding of the C ABI declared in ``include/mpcx.h`` (libmpcx.so).

The library is built in-tree (``dolfinx_mpc_b200/csrc/Makefile``).  There is no
fallback: if it cannot be loaded, every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# MPCX_LIB: alternative build of the same library (kernel tuning experiments, tools/)
LIB_PATH = os.environ.get("MPCX_LIB") or os.path.join(_HERE, "libmpcx.so")

MAX_CONSTANTS = 8

OK, ERR_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_PATTERN, ERR_ALLOC = range(6)


class Tables(C.Structure):
    _fields_ = [("tdim", C.c_int32), ("gdim", C.c_int32), ("nd", C.c_int32), ("ng", C.c_int32),
                ("nq", C.c_int32), ("bs", C.c_int32), ("weights", C.c_void_p), ("phi", C.c_void_p),
                ("dphi", C.c_void_p), ("gdphi", C.c_void_p), ("nfacets", C.c_int32), ("facet_tangents", C.c_void_p),
                ("nd1", C.c_int32), ("bs1", C.c_int32), ("phi1", C.c_void_p), ("dphi1", C.c_void_p)]


class MeshS(C.Structure):
    _fields_ = [("x", C.c_void_p), ("x_dofmap", C.c_void_p), ("num_nodes", C.c_int64), ("ng", C.c_int32),
                ("x_stride", C.c_int32)]


class DofmapS(C.Structure):
    _fields_ = [("map", C.c_void_p), ("nd", C.c_int32), ("bs", C.c_int32), ("num_dofs", C.c_int64),
                ("num_owned_dofs", C.c_int64)]


class MpcS(C.Structure):
    _fields_ = [("is_slave", C.c_void_p), ("masters", C.c_void_p), ("coeffs", C.c_void_p),
                ("offsets", C.c_void_p), ("cell_to_slaves", C.c_void_p), ("cell_to_slaves_offsets", C.c_void_p),
                ("slaves", C.c_void_p), ("num_slaves", C.c_int32), ("num_local_slaves", C.c_int32),
                ("num_dofs", C.c_int64)]


class CsrS(C.Structure):
    _fields_ = [("row_ptr", C.c_void_p), ("col", C.c_void_p), ("val", C.c_void_p), ("num_rows", C.c_int64),
                ("nnz", C.c_int64)]


class IntegralS(C.Structure):
    _fields_ = [("kernel", C.c_int32), ("tables", C.POINTER(Tables)), ("cells", C.c_void_p),
                ("num_cells", C.c_int64), ("coeffs", C.c_void_p), ("cstride", C.c_int32),
                ("coeff_nodal", C.c_void_p), ("coeff_dofmap", C.c_void_p), ("coeff_nd", C.c_int32),
                ("coeff_bs", C.c_int32), ("num_constants", C.c_int32), ("constants", C.c_double * MAX_CONSTANTS),
                ("slave_cells", C.c_void_p), ("num_slave_cells", C.c_int64), ("local_facets", C.c_void_p),
                ("custom", C.c_void_p)]


class PlanS(C.Structure):
    _fields_ = [("lpos", C.c_void_p), ("width", C.c_int32)]


class MpcHostS(C.Structure):
    _fields_ = [("masters", C.c_void_p), ("offsets", C.c_void_p), ("cell_to_slaves", C.c_void_p),
                ("cell_to_slaves_offsets", C.c_void_p)]


# every symbol include/mpcx.h declares (tests check the library exports all of them)
SYMBOLS = (
    "mpcx_last_error", "mpcx_abi_version", "mpcx_device_error", "mpcx_device_error_async", "mpcx_zero_f64", "mpcx_custom_kernel_create", "mpcx_custom_kernel_destroy",
    "mpcx_assemble_matrix_f64",
    "mpcx_add_diagonal_f64", "mpcx_build_plan", "mpcx_assemble_vector_f64", "mpcx_apply_lifting_f64",
    "mpcx_backsubstitution_f64", "mpcx_homogenize_f64", "mpcx_gather_f64", "mpcx_scatter_add_f64",
    "mpcx_create_pattern_host", "mpcx_free_host", "mpcx_profile_enable", "mpcx_launch_count", "mpcx_profile_read",
    "mpcx_flag_cells", "mpcx_tile_plan_create", "mpcx_tile_plan_destroy", "mpcx_tile_plan_info",
    "mpcx_assemble_matrix_tiled_f64", "mpcx_vector_tile_plan_create", "mpcx_assemble_vector_tiled_f64",
    "mpcx_pattern_create", "mpcx_pattern_export", "mpcx_pattern_destroy", "mpcx_assemble_system_tiled_f64", "mpcx_assemble_system_tiled_part_f64", "mpcx_nccl_load", "mpcx_comm_unique_id", "mpcx_comm_create",
    "mpcx_comm_destroy", "mpcx_ghost_reduce_f64", "mpcx_tile_plan_add_slave_cells", "mpcx_row_plan_create", "mpcx_row_plan_destroy",
    "mpcx_assemble_matrix_rowgather_f64", "mpcx_slave_plan_create", "mpcx_slave_plan_destroy",
    "mpcx_assemble_slave_cells_f64",
)

_lib = None


class MpcxError(RuntimeError):
    """Raised for any non-zero status of the C ABI (the reference surfaces std::runtime_error as
    RuntimeError through nanobind, ``cpp/assemble_matrix.cpp:315,464,605,659``)."""


def load():
    """Load libmpcx.so; raises if it is missing (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MpcxError(
            f"{LIB_PATH} not found: build it with `make -C dolfinx_mpc_b200/csrc` "
            "(or __graft_entry__.build()). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.mpcx_last_error.restype = C.c_char_p
    lib.mpcx_abi_version.restype = C.c_int
    vp, i32, i64, f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    P = C.POINTER
    lib.mpcx_device_error.argtypes = [vp]
    lib.mpcx_device_error_async.argtypes = [vp, vp]
    lib.mpcx_zero_f64.argtypes = [vp, C.c_int64, vp]
    lib.mpcx_custom_kernel_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(vp)]
    lib.mpcx_custom_kernel_destroy.argtypes = [vp]
    lib.mpcx_custom_kernel_destroy.restype = None
    lib.mpcx_assemble_matrix_f64.argtypes = [P(IntegralS), P(MeshS), P(DofmapS), P(DofmapS), vp, vp, P(MpcS),
                                             P(MpcS), P(CsrS), P(PlanS), vp]
    lib.mpcx_add_diagonal_f64.argtypes = [P(CsrS), vp, i64, f64, vp]
    lib.mpcx_build_plan.argtypes = [P(DofmapS), P(DofmapS), vp, i64, P(CsrS), vp, i32, vp]
    lib.mpcx_assemble_vector_f64.argtypes = [P(IntegralS), P(MeshS), P(DofmapS), P(MpcS), vp, vp]
    lib.mpcx_apply_lifting_f64.argtypes = [P(IntegralS), P(MeshS), P(DofmapS), P(DofmapS), vp, vp, vp, f64,
                                           P(MpcS), vp, i64, vp, vp]
    lib.mpcx_tile_plan_create.argtypes = [P(MeshS), P(DofmapS), P(DofmapS), vp, i64, vp, vp, vp, P(CsrS), vp, P(vp)]
    lib.mpcx_tile_plan_destroy.argtypes = [vp]
    lib.mpcx_tile_plan_destroy.restype = None
    lib.mpcx_tile_plan_info.argtypes = [vp, P(i64), i32]
    lib.mpcx_assemble_matrix_tiled_f64.argtypes = [P(IntegralS), P(MeshS), P(DofmapS), P(DofmapS), vp, vp, P(MpcS),
                                                   P(MpcS), P(CsrS), vp, vp]
    lib.mpcx_vector_tile_plan_create.argtypes = [P(MeshS), P(DofmapS), vp, i64, vp, vp, P(vp)]
    lib.mpcx_assemble_vector_tiled_f64.argtypes = [P(IntegralS), P(MeshS), P(DofmapS), P(MpcS), vp, vp, vp]
    lib.mpcx_assemble_system_tiled_f64.argtypes = [P(IntegralS), P(IntegralS), P(MeshS), P(DofmapS), vp, P(MpcS), P(CsrS),
                                                   vp, vp, vp, vp]
    lib.mpcx_tile_plan_add_slave_cells.argtypes = [vp, P(IntegralS), P(DofmapS), P(DofmapS), vp, vp, P(MpcS), P(MpcS), P(CsrS), vp]
    lib.mpcx_row_plan_create.argtypes = [P(DofmapS), vp, i64, vp, P(CsrS), vp, P(vp)]
    lib.mpcx_row_plan_destroy.argtypes = [vp]
    lib.mpcx_row_plan_destroy.restype = None
    lib.mpcx_assemble_matrix_rowgather_f64.argtypes = [P(IntegralS), P(MeshS), P(DofmapS), vp, P(MpcS), P(CsrS), vp, vp, vp]
    lib.mpcx_slave_plan_create.argtypes = [P(IntegralS), P(DofmapS), P(DofmapS), vp, vp, P(MpcS), P(MpcS), P(CsrS), vp, P(vp)]
    lib.mpcx_slave_plan_destroy.argtypes = [vp]
    lib.mpcx_slave_plan_destroy.restype = None
    lib.mpcx_assemble_slave_cells_f64.argtypes = [P(IntegralS), P(MeshS), P(MpcS), P(MpcS), P(CsrS), vp, vp]
    lib.mpcx_nccl_load.argtypes = [C.c_char_p]
    lib.mpcx_comm_unique_id.argtypes = [vp]
    lib.mpcx_comm_create.argtypes = [vp, i32, i32, P(vp)]
    lib.mpcx_comm_destroy.argtypes = [vp]
    lib.mpcx_comm_destroy.restype = None
    lib.mpcx_ghost_reduce_f64.argtypes = [vp, vp, vp, i64, P(i64), vp, P(i64), vp, vp, vp]
    lib.mpcx_assemble_system_tiled_part_f64.argtypes = [P(IntegralS), P(IntegralS), P(MeshS), P(DofmapS), vp, P(MpcS),
                                                        P(CsrS), vp, vp, vp, i32, vp]
    lib.mpcx_flag_cells.argtypes = [P(DofmapS), vp, i64, vp, vp, vp]
    lib.mpcx_backsubstitution_f64.argtypes = [P(MpcS), vp, vp]
    lib.mpcx_homogenize_f64.argtypes = [P(MpcS), vp, vp]
    lib.mpcx_gather_f64.argtypes = [vp, vp, i64, vp, vp]
    lib.mpcx_scatter_add_f64.argtypes = [vp, vp, i64, vp, vp]
    lib.mpcx_create_pattern_host.argtypes = [vp, i32, i32, vp, i32, i32, i64, i64, P(MpcHostS), P(MpcHostS), i32,
                                             P(P(C.c_int64)), P(P(C.c_int32)), P(i64)]
    lib.mpcx_pattern_create.argtypes = [P(DofmapS), P(DofmapS), i64, P(MpcS), P(MpcS), vp, P(vp), P(i64)]
    lib.mpcx_pattern_export.argtypes = [vp, vp, vp, vp]
    lib.mpcx_pattern_destroy.argtypes = [vp]
    lib.mpcx_pattern_destroy.restype = None
    lib.mpcx_profile_enable.argtypes = [C.c_int]
    lib.mpcx_launch_count.restype = C.c_longlong
    lib.mpcx_profile_read.argtypes = [P(C.c_double), P(C.c_longlong)]
    lib.mpcx_free_host.argtypes = [vp]
    lib.mpcx_free_host.restype = None
    _lib = lib
    return lib


def check(status: int):
    if status != OK:
        msg = load().mpcx_last_error().decode()
        err = MpcxError(f"mpcx status {status}: {msg}")
        err.status = status
        raise err
