"""Custom element kernels: the reference takes an opaque ``tabulate_tensor`` pointer from the form
(``cpp/assemble_matrix.cpp:438-439, 505-506, 620-636``); here the same C source is compiled with NVRTC for the device
(``mpcx_custom_kernel_create``) and with gcc for the oracle (``orc_set_custom_kernel``), and both sides must agree --
on a form the registry also has (P1 Laplace on triangles) and on one it has not (a non-symmetric convection-diffusion-
reaction operator with a vector-valued velocity coefficient, plus its load vector)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import problems
from test_gpu_parity import _mpc, assert_csr_close, assert_vec_close

# UFCx signature; closed forms (affine P1), so no tables are needed
LAPLACE_P1_TRI = r"""
#include <math.h>
#include <stdint.h>
void laplace_p1_tri(double* restrict A, const double* restrict w, const double* restrict c,
                    const double* restrict x, const int* restrict entity_local_index,
                    const uint8_t* restrict quadrature_permutation)
{
  const double J00 = x[3] - x[0], J01 = x[6] - x[0], J10 = x[4] - x[1], J11 = x[7] - x[1];
  const double det = J00 * J11 - J01 * J10;
  /* gradients of the barycentric coordinates */
  const double g[3][2] = {{(J10 - J11) / det, (J01 - J00) / det}, {J11 / det, -J01 / det}, {-J10 / det, J00 / det}};
  const double area = 0.5 * fabs(det);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[3 * i + j] += c[0] * area * (g[i][0] * g[j][0] + g[i][1] * g[j][1]);
}
"""

# a(u, v) = kappa grad u . grad v + (beta . grad u) v + sigma u v on P1 tetrahedra, beta a P1 vector field (w[4][3]);
# L(v) = s * f v with f a P1 scalar field (w[4])
CDR_P1_TET = r"""
#include <math.h>
#include <stdint.h>
static void tet_geometry(const double* x, double g[4][3], double* vol)
{
  double J[3][3];
  for (int k = 0; k < 3; ++k)
    for (int a = 0; a < 3; ++a) J[k][a] = x[3 * (a + 1) + k] - x[k];
  const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0])
                     + J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
  /* rows of J^-1 = gradients of lambda_1..3 */
  g[1][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) / det; g[1][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) / det;
  g[1][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) / det;
  g[2][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) / det; g[2][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) / det;
  g[2][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) / det;
  g[3][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) / det; g[3][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) / det;
  g[3][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) / det;
  for (int k = 0; k < 3; ++k) g[0][k] = -(g[1][k] + g[2][k] + g[3][k]);
  *vol = fabs(det) / 6.0;
}
void cdr_p1_tet(double* restrict A, const double* restrict w, const double* restrict c, const double* restrict x,
                const int* restrict entity_local_index, const uint8_t* restrict quadrature_permutation)
{
  double g[4][3], vol;
  tet_geometry(x, g, &vol);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
    {
      double v = c[0] * (g[i][0] * g[j][0] + g[i][1] * g[j][1] + g[i][2] * g[j][2]);
      double conv = 0.0; /* sum_k (beta_k . grad phi_j) int phi_k phi_i / vol */
      for (int k = 0; k < 4; ++k)
        conv += (w[3 * k] * g[j][0] + w[3 * k + 1] * g[j][1] + w[3 * k + 2] * g[j][2]) * ((k == i ? 2.0 : 1.0) / 20.0);
      v += conv + c[1] * ((i == j ? 2.0 : 1.0) / 20.0);
      A[4 * i + j] += vol * v;
    }
}
void load_p1_tet(double* restrict b, const double* restrict w, const double* restrict c, const double* restrict x,
                 const int* restrict entity_local_index, const uint8_t* restrict quadrature_permutation)
{
  double g[4][3], vol;
  tet_geometry(x, g, &vol);
  for (int i = 0; i < 4; ++i)
    for (int k = 0; k < 4; ++k) b[i] += c[0] * vol * w[k] * ((k == i ? 2.0 : 1.0) / 20.0);
}
"""


def _host_fn(tmp_path, source: str, entry: str):
    """The same source compiled for the host: the function pointer the reference would get from the form."""
    src = tmp_path / f"{entry}.c"
    src.write_text(source)
    so = tmp_path / f"{entry}.so"
    subprocess.run(["gcc", "-O2", "-std=c99", "-shared", "-fPIC", "-o", str(so), str(src), "-lm"], check=True)
    lib = C.CDLL(str(so))
    return lib, C.cast(getattr(lib, entry), C.c_void_p)


def test_custom_kernel_compiles_and_reports_errors():
    """NVRTC needs no GPU: the handle is created here; a source that does not compile raises with the compiler's log."""
    from dolfinx_mpc_b200 import _lib

    lib = _lib.load()
    h = C.c_void_p()
    _lib.check(lib.mpcx_custom_kernel_create(LAPLACE_P1_TRI.encode(), b"laplace_p1_tri", 9, 3, 0, C.byref(h)))
    assert h.value
    lib.mpcx_custom_kernel_destroy(h)
    bad = LAPLACE_P1_TRI.replace("const double area", "const double area_ = undefined_symbol; const double area")
    with pytest.raises(_lib.MpcxError) as e:
        _lib.check(lib.mpcx_custom_kernel_create(bad.encode(), b"laplace_p1_tri", 9, 3, 0, C.byref(h)))
    assert "undefined_symbol" in str(e.value)


def test_oracle_custom_kernel_equals_registry(oracle, tmp_path):
    """The oracle's K_CUSTOM path (the reference's opaque fn) reproduces its registry kernel for the same form."""
    from dolfinx_mpc_b200 import fem

    c = problems.ALL_CASES["periodic2d-P1-8-bc1"]()
    m = oracle.mpc_from_arrays(c.V, c.data)
    keep, fn = _host_fn(tmp_path, LAPLACE_P1_TRI, "laplace_p1_tri")
    oracle.lib().orc_set_custom_kernel(fn)
    a_c = fem.custom_form((c.V, c.V), fem.CustomKernel(LAPLACE_P1_TRI, "laplace_p1_tri"), constants=[1.0])
    a_r = fem.laplace(c.V, 1.0)
    rp, col, val = oracle.assemble_matrix(a_c, m, bcs=c.bcs)
    rp_r, col_r, val_r = oracle.assemble_matrix(a_r, m, bcs=c.bcs)
    assert_csr_close(rp, col, val, rp_r, col_r, val_r)
    del keep


@pytest.mark.gpu
def test_custom_laplace_matches_registry_and_oracle(oracle, tmp_path):
    import dolfinx_mpc_b200 as mpcx
    from dolfinx_mpc_b200 import fem

    c = problems.ALL_CASES["periodic2d-P1-8-bc1"]()
    mpc = _mpc(c)
    m = oracle.mpc_from_arrays(c.V, c.data)
    a_c = fem.custom_form((c.V, c.V), fem.CustomKernel(LAPLACE_P1_TRI, "laplace_p1_tri"), constants=[1.0])
    A = mpcx.assemble_matrix(a_c, mpc, bcs=c.bcs)
    A_r = mpcx.assemble_matrix(fem.laplace(c.V, 1.0), mpc, bcs=c.bcs)
    assert_csr_close(*A.getValuesCSR(), *A_r.getValuesCSR())
    keep, fn = _host_fn(tmp_path, LAPLACE_P1_TRI, "laplace_p1_tri")
    oracle.lib().orc_set_custom_kernel(fn)
    assert_csr_close(*A.getValuesCSR(), *oracle.assemble_matrix(a_c, m, bcs=c.bcs))
    mpcx.assemble_matrix(a_c, mpc, bcs=c.bcs, A=A)  # cached handle, plans and scratch
    assert_csr_close(*A.getValuesCSR(), *A_r.getValuesCSR())
    del keep


@pytest.mark.gpu
@pytest.mark.parametrize("chunk_cells", [None, "37"])
@pytest.mark.parametrize("name", ["periodic3d-P1-bs1-4-ax2-bc1", "periodic3d-P1-bs1-3-ax3-bc0"])
def test_custom_convection_diffusion_matches_oracle(oracle, tmp_path, monkeypatch, name, chunk_cells):
    """A form the registry does not have: non-symmetric matrix with a vector coefficient, its lifting and a load vector
    with a scalar coefficient, through MPC elimination and Dirichlet conditions, against the oracle running the SAME
    source compiled for the host.  ``chunk_cells``: the element tensors evaluated 37 cells at a time (the scratch array
    of a custom kernel is bounded, MPCX_CUSTOM_SCRATCH_MB), every chunk followed by the elimination / scatter kernels."""
    import dolfinx_mpc_b200 as mpcx
    from dolfinx_mpc_b200 import fem, generators as gen

    if chunk_cells:
        monkeypatch.setenv("MPCX_CUSTOM_CHUNK_CELLS", chunk_cells)
    c = problems.ALL_CASES[name]()
    V = c.V
    mpc = _mpc(c)
    m = oracle.mpc_from_arrays(V, c.data)
    W = gen.functionspace(V.mesh, 1, 3)
    beta = fem.Function(W)
    xb = np.asarray(V.mesh.x)[:, :3]
    beta.array[:] = np.stack([1.0 + xb[:, 1], -0.5 + xb[:, 0] * xb[:, 2], 0.3 * np.cos(xb[:, 0])], axis=1).reshape(-1)[: len(beta.array)]
    f = fem.Function(V)
    f.array[:] = (np.sin(3 * xb[:, 0]) + xb[:, 1] * xb[:, 2])[: len(f.array)]
    ka = fem.CustomKernel(CDR_P1_TET, "cdr_p1_tet")
    kl = fem.CustomKernel(CDR_P1_TET, "load_p1_tet")
    a = fem.custom_form((V, V), ka, constants=[0.7, 2.0], coefficients=[beta])
    L = fem.custom_form((V,), kl, constants=[1.5], coefficients=[f])

    keep_a, fn_a = _host_fn(tmp_path, CDR_P1_TET, "cdr_p1_tet")
    A = mpcx.assemble_matrix(a, mpc, bcs=c.bcs)
    oracle.lib().orc_set_custom_kernel(fn_a)
    rp_o, col_o, val_o = oracle.assemble_matrix(a, m, bcs=c.bcs)
    assert_csr_close(*A.getValuesCSR(), rp_o, col_o, val_o)
    asym = A.to_scipy()
    assert abs(asym - asym.T).max() > 1e-3  # the operator is not symmetric: nothing relied on symmetry

    b = mpcx.assemble_vector(L, mpc)
    keep_l = C.CDLL(keep_a._name)
    oracle.lib().orc_set_custom_kernel(C.cast(keep_l.load_p1_tet, C.c_void_p))
    b_o = oracle.assemble_vector(L, m)
    assert_vec_close(b.array, b_o)
    if c.bcs:
        mpcx.apply_lifting(b, [a], [c.bcs], mpc)
        oracle.lib().orc_set_custom_kernel(fn_a)
        oracle.apply_lifting(b_o, [a], [c.bcs], m)
        assert_vec_close(b.array, b_o)


# Robin term int u v ds on a facet of a P1 tetrahedron (local facet f is opposite vertex f): the exterior-facet call
# of cpp/assemble_matrix.cpp:361-362, entity_local_index = the local facet
FACET_MASS_P1_TET = r"""
#include <math.h>
void facet_mass_p1_tet(double* restrict A, const double* restrict w, const double* restrict c, const double* restrict x,
                       const int* restrict entity_local_index, const unsigned char* restrict quadrature_permutation)
{
  const int f = entity_local_index[0];
  int v[3], n = 0;
  for (int k = 0; k < 4; ++k)
    if (k != f) v[n++] = k;
  double e1[3], e2[3];
  for (int k = 0; k < 3; ++k) { e1[k] = x[3 * v[1] + k] - x[3 * v[0] + k]; e2[k] = x[3 * v[2] + k] - x[3 * v[0] + k]; }
  const double cx = e1[1] * e2[2] - e1[2] * e2[1], cy = e1[2] * e2[0] - e1[0] * e2[2], cz = e1[0] * e2[1] - e1[1] * e2[0];
  const double area = 0.5 * sqrt(cx * cx + cy * cy + cz * cz);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[4 * v[i] + v[j]] += c[0] * area * ((i == j ? 2.0 : 1.0) / 12.0);
}
"""

# plumbing check for a block between DIFFERENT elements (P2 vector test space x P1 scalar trial space): a kernel with
# recognisable entries, the same on both sides
PATTERN_RECT = r"""
void pattern_rect(double* restrict A, const double* restrict w, const double* restrict c, const double* restrict x,
                  const int* restrict entity_local_index, const unsigned char* restrict quadrature_permutation)
{
  const double s = x[3] - x[0] + 2.0 * (x[7] - x[1]);  /* depends on the cell */
  for (int i = 0; i < 12; ++i)
    for (int j = 0; j < 3; ++j) A[3 * i + j] += c[0] * (1.0 + i + 10.0 * j) * s;
}
"""


@pytest.mark.gpu
def test_custom_exterior_facet_kernel_matches_registry_and_oracle(oracle, tmp_path):
    import dolfinx_mpc_b200 as mpcx
    from dolfinx_mpc_b200 import fem

    c = problems.ALL_CASES["surface-robin3d-P1"]()
    V = c.V
    mpc = _mpc(c)
    m = oracle.mpc_from_arrays(V, c.data)
    allf = fem.locate_exterior_facets(V.mesh)
    k = fem.CustomKernel(FACET_MASS_P1_TET, "facet_mass_p1_tet")
    a_c = fem.laplace(V) + fem.custom_form((V, V), k, constants=[1.5], facets=allf)
    a_r = fem.laplace(V) + fem.mass(V, 1.5, facets=allf)
    A = mpcx.assemble_matrix(a_c, mpc, bcs=c.bcs)
    A_r = mpcx.assemble_matrix(a_r, mpc, bcs=c.bcs)
    assert_csr_close(*A.getValuesCSR(), *A_r.getValuesCSR())
    keep, fn = _host_fn(tmp_path, FACET_MASS_P1_TET, "facet_mass_p1_tet")
    oracle.lib().orc_set_custom_kernel(fn)
    assert_csr_close(*A.getValuesCSR(), *oracle.assemble_matrix(a_c, m, bcs=c.bcs))
    del keep


@pytest.mark.gpu
def test_custom_kernel_between_different_elements(oracle, tmp_path):
    import dolfinx_mpc_b200 as mpcx
    from dolfinx_mpc_b200 import MultiPointConstraint, fem

    sc = problems.case_stokes_2d()
    mpc_v = MultiPointConstraint(sc.V)
    mpc_v.add_constraint(sc.V, *sc.data_v)
    mpc_v.finalize()
    mpc_q = MultiPointConstraint(sc.Q)
    mpc_q.finalize()
    a01 = fem.custom_form((sc.V, sc.Q), fem.CustomKernel(PATTERN_RECT, "pattern_rect"), constants=[0.5])
    A = mpcx.assemble_matrix(a01, [mpc_v, mpc_q], bcs=sc.bcs)
    keep, fn = _host_fn(tmp_path, PATTERN_RECT, "pattern_rect")
    oracle.lib().orc_set_custom_kernel(fn)
    mv = oracle.mpc_from_arrays(sc.V, sc.data_v)
    mq = oracle.OracleMPC.empty(sc.Q)
    assert_csr_close(*A.getValuesCSR(), *oracle.assemble_matrix(a01, mv, mq, bcs=sc.bcs, same_space=False))
    del keep


# In the STYLE FFCx generates (function-scope static tables, quadrature loop, `restrict`, ufcx.h include): the mass
# matrix of P1 on triangles with a 3-point degree-2 rule.  (FFCx itself is not installed here: hand-written after its
# output format.)
FFCX_STYLE_MASS_P1_TRI = r"""
#include <math.h>
#include <stdint.h>
#include <ufcx.h>

void tabulate_tensor_integral_mass_p1(double* restrict A,
                                      const double* restrict w,
                                      const double* restrict c,
                                      const double* restrict coordinate_dofs,
                                      const int* restrict entity_local_index,
                                      const uint8_t* restrict quadrature_permutation)
{
// Quadrature rules
static const double weights_e1b[3] = {0.1666666666666667, 0.1666666666666667, 0.1666666666666667};
// Precomputed values of basis functions and precomputations
// FE* dimensions: [permutation][entities][points][dofs]
static const double FE0_C0_D10_Qe1b[1][1][1][3] = {{{{-1.0, 1.0, 0.0}}}};
static const double FE0_C0_D01_Qe1b[1][1][1][3] = {{{{-1.0, 0.0, 1.0}}}};
static const double FE1_C0_Qe1b[1][1][3][3] = {{{{0.6666666666666667, 0.1666666666666667, 0.1666666666666667},
  {0.1666666666666667, 0.1666666666666667, 0.6666666666666667},
  {0.1666666666666667, 0.6666666666666667, 0.1666666666666667}}}};
// ------------------------
// Section: Jacobian
double J_c0 = 0.0;
double J_c3 = 0.0;
double J_c1 = 0.0;
double J_c2 = 0.0;
for (int ic = 0; ic < 3; ++ic)
{
  J_c0 += coordinate_dofs[(ic) * 3] * FE0_C0_D10_Qe1b[0][0][0][ic];
  J_c3 += coordinate_dofs[(ic) * 3 + 1] * FE0_C0_D01_Qe1b[0][0][0][ic];
  J_c1 += coordinate_dofs[(ic) * 3] * FE0_C0_D01_Qe1b[0][0][0][ic];
  J_c2 += coordinate_dofs[(ic) * 3 + 1] * FE0_C0_D10_Qe1b[0][0][0][ic];
}
// ------------------------
double sp_e1b_0 = J_c0 * J_c3;
double sp_e1b_1 = J_c1 * J_c2;
double sp_e1b_2 = -sp_e1b_1;
double sp_e1b_3 = sp_e1b_0 + sp_e1b_2;
double sp_e1b_4 = fabs(sp_e1b_3);
double sp_e1b_5 = c[0] * sp_e1b_4;
for (int iq = 0; iq < 3; ++iq)
{
  const double fw0 = sp_e1b_5 * weights_e1b[iq];
  double t0[3];
  for (int i = 0; i < 3; ++i)
    t0[i] = fw0 * FE1_C0_Qe1b[0][0][iq][i];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      A[3 * (i) + (j)] += FE1_C0_Qe1b[0][0][iq][j] * t0[i];
}
}
"""


def test_ffcx_style_source_compiles():
    from dolfinx_mpc_b200 import _lib

    h = C.c_void_p()
    _lib.check(_lib.load().mpcx_custom_kernel_create(FFCX_STYLE_MASS_P1_TRI.encode(), b"tabulate_tensor_integral_mass_p1",
                                                    9, 3, 0, C.byref(h)))
    _lib.load().mpcx_custom_kernel_destroy(h)


@pytest.mark.gpu
def test_ffcx_style_mass_matches_registry(oracle):
    import dolfinx_mpc_b200 as mpcx
    from dolfinx_mpc_b200 import fem

    c = problems.ALL_CASES["periodic2d-P1-8-bc1"]()
    mpc = _mpc(c)
    k = fem.CustomKernel(FFCX_STYLE_MASS_P1_TRI, "tabulate_tensor_integral_mass_p1")
    A = mpcx.assemble_matrix(fem.custom_form((c.V, c.V), k, constants=[2.5]), mpc, bcs=c.bcs)
    A_r = mpcx.assemble_matrix(fem.mass(c.V, 2.5), mpc, bcs=c.bcs)
    assert_csr_close(*A.getValuesCSR(), *A_r.getValuesCSR())
