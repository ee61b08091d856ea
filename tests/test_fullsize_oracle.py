"""The slab-parallel oracle driver used by the full-size GPU parity test (oracle.assemble_system_slabs) gives exactly
what the serial oracle routines give -- checked here on a size the serial oracle does in a blink."""
import numpy as np


def test_slab_parallel_oracle_equals_serial_oracle(oracle):
    import bench

    n = 11
    P = bench.build_problem(n)
    m = oracle.mpc_from_arrays(P["V"], P["data"])
    pat = oracle.create_pattern(P["a"], m, m)
    _, _, val = oracle.assemble_matrix(P["a"], m, bcs=P["bcs"], pattern=pat)
    b = oracle.assemble_vector(P["L"], m)
    oracle.apply_lifting(b, [P["a"]], [P["bcs"]], m)
    val_p, b_p = oracle.assemble_system_slabs(P["a"], P["L"], m, P["bcs"], pat, 6 * (n - 1) ** 2, nthreads=3)
    # same per-cell arithmetic; only the order in which cells of different slabs reach a shared row differs
    assert np.abs(val_p - val).max() <= 1e-13 * np.abs(val).max()
    assert np.abs(b_p - b).max() <= 1e-13 * np.abs(b).max()
