"""B200-native MPC-constrained finite-element assembly (drop-in for dolfinx_mpc's hot path).

Exports the names of ``python/src/dolfinx_mpc/__init__.py:14-41`` that belong to the assembly path.
Importing the package does not need a GPU; calling any assembly routine does (no CPU fallback).
"""
from .assemble_matrix import (assemble_matrix, assemble_matrix_nest, create_matrix, create_matrix_nest,
                              create_sparsity_pattern, create_sparsity_pattern_device)
from .assemble_vector import (apply_lifting, assemble_vector, assemble_vector_nest, create_vector,
                              create_vector_nest, set_bc)
from .multipointconstraint import MultiPointConstraint
from .problem import assemble_system

__all__ = [
    "assemble_matrix", "create_matrix", "create_matrix_nest", "assemble_matrix_nest", "assemble_vector",
    "apply_lifting", "assemble_vector_nest", "create_vector_nest", "create_vector", "set_bc",
    "MultiPointConstraint", "create_sparsity_pattern", "create_sparsity_pattern_device", "assemble_system",
]
__version__ = "0.1.0"
