/*
 * mpc_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the MPC-constrained assembly algorithm of
 * jorgensd/dolfinx_mpc @ 8dd7891e.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference leg may load this library; the
 * product path (dolfinx_mpc_b200 + libmpcx.so) never does.
 *
 * PARITY STATUS: "parity unpinned" for the element-tensor values.  The
 * reference cannot be built or imported in this image (needs DOLFINx, PETSc,
 * MPI, FFCx) and its repository holds no golden vectors; the element kernel
 * (FFCx tabulate_tensor) is third-party generated code that is absent from
 * /root/reference.  It is restated here as tabulated-basis quadrature, the
 * scheme FFCx generates (fenics-ffcx, pinned only as "matches
 * fenics-dolfinx>=0.12.0.dev0", python/pyproject.toml:23).  What IS pinned:
 * the elimination + scatter, through the reference's own test identities
 * (K^H A K == A_mpc[free,free], K^H b == b_mpc[free], b_mpc[slaves] == 0,
 * python/src/dolfinx_mpc/utils/test.py:202-265) checked in tests/ on
 * re-creations of the reference's test fixtures.
 *
 * Each function cites the reference lines it follows (paths relative to
 * /root/reference).  The loop structure, the order of insertions and the
 * per-cell heap allocations of modify_mpc_cell are kept on purpose: this file
 * is also the CPU baseline timed by bench.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_ERR_PATTERN 1  /* insertion outside the sparsity pattern */
#define ORC_ERR_KERNEL 2   /* unknown kernel id / unsupported combination */
#define ORC_ERR_ALLOC 3
#define ORC_ERR_GEOM 4

/* kernel ids -- must match include/mpcx.h */
enum
{
  K_LAPLACE = 0,
  K_MASS = 1,
  K_ELASTICITY = 2,
  K_SOURCE = 3,
  K_LAPLACE_VARCOEF = 4,
  K_DIV_TEST = 5,  /* c0 * inner(p, div v) dx: test = vector element, trial = scalar element */
  K_DIV_TRIAL = 6, /* c0 * inner(div u, q) dx: test = scalar element, trial = vector element */
  K_CUSTOM = 7     /* the function set with orc_set_custom_kernel: the reference's opaque `fn` itself */
};

/* Tabulated element: the tables an FFCx kernel has baked into its source. */
typedef struct
{
  int32_t tdim, gdim, nd, ng, nq, bs;
  const double* weights; /* [nq] */
  const double* phi;     /* [nq][nd] */
  const double* dphi;    /* [nq][tdim][nd]  reference derivatives */
  const double* gdphi;   /* [nq][tdim][ng]  geometry-map reference derivatives */
  /* exterior-facet integrals: the arrays above hold nfacets consecutive tables, one per local facet; ftan are the
   * tangents of the reference facet map, [nfacets][tdim-1][tdim].  nfacets == 0 for cell integrals. */
  int32_t nfacets;
  const double* ftan;
  /* rectangular forms (test and trial elements differ; dofs[2] / bs[2] / num_dofs[2] of modify_mpc_cell,
   * cpp/assemble_matrix.cpp:99-117): the trial element at the same quadrature points; nd1 == 0: same element. */
  int32_t nd1, bs1;
  const double* phi1;  /* [nq][nd1] */
  const double* dphi1; /* [nq][tdim][nd1] */
} orc_tables;

/* The tables of one entity: a cell (e == NULL) or local facet e[0] of a cell -- the entity_local_index argument of
 * the UFCx kernel, passed at cpp/assemble_matrix.cpp:361-362 for exterior facets. */
static orc_tables entity_view(const orc_tables* t, const int* e)
{
  orc_tables v = *t;
  if (e && t->nfacets > 0)
  {
    const int f = e[0];
    v.phi += (size_t)f * t->nq * t->nd;
    v.dphi += (size_t)f * t->nq * t->tdim * t->nd;
    v.gdphi += (size_t)f * t->nq * t->tdim * t->ng;
    v.ftan = t->ftan + (size_t)f * (t->tdim - 1) * t->tdim;
  }
  else
    v.ftan = NULL;
  return v;
}

typedef void (*ufcx_kernel)(double* A, const double* w, const double* c,
                            const double* coordinate_dofs,
                            const int* entity_local_index,
                            const uint8_t* quadrature_permutation,
                            void* custom_data);

/* ---- geometry at one quadrature point: J = sum_g X_g (x) dpsi_g, K = J^-1 */
/* On return *detJ holds the signed determinant for a cell integral and the (positive) surface measure
 * |J t| (2-D) / |J t1 x J t2| (3-D) for a facet view, so that weights[q] * fabs(*detJ) is the scale in both cases. */
static int jacobian_raw(const orc_tables* t, int q, const double* X, double* K,
                        double* detJ, double* J);
static int jacobian(const orc_tables* t, int q, const double* X, double* K,
                    double* detJ)
{
  double J[9];
  if (jacobian_raw(t, q, X, K, detJ, J)) return 1;
  if (t->ftan)
  {
    const int td = t->tdim;
    double a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
    for (int k = 0; k < td; ++k)
      for (int c = 0; c < td; ++c)
      {
        a[k] += J[k * 3 + c] * t->ftan[c];
        if (td == 3) b[k] += J[k * 3 + c] * t->ftan[td + c];
      }
    if (td == 2)
      *detJ = sqrt(a[0] * a[0] + a[1] * a[1]);
    else
    {
      const double cx = a[1] * b[2] - a[2] * b[1], cy = a[2] * b[0] - a[0] * b[2], cz = a[0] * b[1] - a[1] * b[0];
      *detJ = sqrt(cx * cx + cy * cy + cz * cz);
    }
  }
  return 0;
}
static int jacobian_raw(const orc_tables* t, int q, const double* X, double* K,
                        double* detJ, double* J)
{
  for (int i = 0; i < 9; ++i) J[i] = 0; /* J[k][a], k<gdim, a<tdim */
  const int td = t->tdim, gd = t->gdim;
  for (int a = 0; a < td; ++a)
    for (int g = 0; g < t->ng; ++g)
    {
      const double d = t->gdphi[(q * td + a) * t->ng + g];
      for (int k = 0; k < gd; ++k)
        J[k * 3 + a] += X[3 * g + k] * d;
    }
  if (td == 2 && gd == 2)
  {
    const double det = J[0] * J[4] - J[1] * J[3];
    *detJ = det;
    /* K[a][k] */
    K[0] = J[4] / det;
    K[1] = -J[1] / det;
    K[3] = -J[3] / det;
    K[4] = J[0] / det;
    return 0;
  }
  if (td == 3 && gd == 3)
  {
    const double c00 = J[4] * J[8] - J[5] * J[7];
    const double c01 = J[5] * J[6] - J[3] * J[8];
    const double c02 = J[3] * J[7] - J[4] * J[6];
    const double det = J[0] * c00 + J[1] * c01 + J[2] * c02;
    *detJ = det;
    K[0] = c00 / det;
    K[1] = (J[2] * J[7] - J[1] * J[8]) / det;
    K[2] = (J[1] * J[5] - J[2] * J[4]) / det;
    K[3] = c01 / det;
    K[4] = (J[0] * J[8] - J[2] * J[6]) / det;
    K[5] = (J[2] * J[3] - J[0] * J[5]) / det;
    K[6] = c02 / det;
    K[7] = (J[1] * J[6] - J[0] * J[7]) / det;
    K[8] = (J[0] * J[4] - J[1] * J[3]) / det;
    return 0;
  }
  return 1;
}

/* Affine cells (simplex, ng == tdim + 1) have a constant Jacobian: FFCx-generated kernels
 * evaluate it once, outside the quadrature loop. */
static int is_affine(const orc_tables* t) { return t->ng == t->tdim + 1; }
#define GEOM_AT(q) \
  if ((q) == 0 || !affine) { if (jacobian(t, (q), X, K, &detJ)) return; }

/* physical gradients g[i][k] = sum_a K[a][k] dphi[q][a][i] */
static void phys_grads(const orc_tables* t, int q, const double* K, double* g)
{
  for (int i = 0; i < t->nd; ++i)
    for (int k = 0; k < t->gdim; ++k)
    {
      double s = 0;
      for (int a = 0; a < t->tdim; ++a)
        s += K[a * 3 + k] * t->dphi[(q * t->tdim + a) * t->nd + i];
      g[i * 3 + k] = s;
    }
}

/* The element kernels.  ABI and accumulate-into-zeroed-buffer convention of the
 * UFCx tabulate_tensor call site, cpp/assemble_matrix.cpp:438-439,504-506.
 * A_e is row-major [nd*bs][nd*bs], interleaved components (i*bs + c). */
static void k_laplace(double* A, const double* w, const double* c,
                      const double* X, const int* e, const uint8_t* p, void* cd)
{
  (void)w; (void)p;
  const orc_tables tv = entity_view((const orc_tables*)cd, e);
  const orc_tables* t = &tv;
  const int n = t->nd * t->bs, bs = t->bs;
  double K[9], detJ = 0, g[3 * 64];
  const int affine = is_affine(t);
  for (int q = 0; q < t->nq; ++q)
  {
    GEOM_AT(q)
    phys_grads(t, q, K, g);
    const double s = c[0] * t->weights[q] * fabs(detJ);
    for (int i = 0; i < t->nd; ++i)
      for (int j = 0; j < t->nd; ++j)
      {
        double d = 0;
        for (int k = 0; k < t->gdim; ++k) d += g[i * 3 + k] * g[j * 3 + k];
        d *= s;
        for (int b = 0; b < bs; ++b) A[(i * bs + b) * n + j * bs + b] += d;
      }
  }
}

static void k_laplace_varcoef(double* A, const double* w, const double* c,
                              const double* X, const int* e, const uint8_t* p,
                              void* cd)
{
  (void)p;
  const orc_tables tv = entity_view((const orc_tables*)cd, e);
  const orc_tables* t = &tv;
  const int n = t->nd * t->bs, bs = t->bs;
  double K[9], detJ = 0, g[3 * 64];
  const int affine = is_affine(t);
  for (int q = 0; q < t->nq; ++q)
  {
    GEOM_AT(q)
    phys_grads(t, q, K, g);
    double kap = 0;
    for (int k = 0; k < t->nd; ++k) kap += t->phi[q * t->nd + k] * w[k];
    const double s = c[0] * kap * t->weights[q] * fabs(detJ);
    for (int i = 0; i < t->nd; ++i)
      for (int j = 0; j < t->nd; ++j)
      {
        double d = 0;
        for (int k = 0; k < t->gdim; ++k) d += g[i * 3 + k] * g[j * 3 + k];
        d *= s;
        for (int b = 0; b < bs; ++b) A[(i * bs + b) * n + j * bs + b] += d;
      }
  }
}

static void k_mass(double* A, const double* w, const double* c, const double* X,
                   const int* e, const uint8_t* p, void* cd)
{
  (void)w; (void)p;
  const orc_tables tv = entity_view((const orc_tables*)cd, e);
  const orc_tables* t = &tv;
  const int n = t->nd * t->bs, bs = t->bs;
  double K[9], detJ = 0;
  const int affine = is_affine(t);
  for (int q = 0; q < t->nq; ++q)
  {
    GEOM_AT(q)
    const double s = c[0] * t->weights[q] * fabs(detJ);
    for (int i = 0; i < t->nd; ++i)
      for (int j = 0; j < t->nd; ++j)
      {
        const double d = s * t->phi[q * t->nd + i] * t->phi[q * t->nd + j];
        for (int b = 0; b < bs; ++b) A[(i * bs + b) * n + j * bs + b] += d;
      }
  }
}

/* inner(sigma(u), grad(v)) dx, sigma = 2 mu sym(grad u) + lambda tr(sym grad u) I
 * (python/benchmarks/bench_elasticity_edge.py:125-135) */
static void k_elasticity(double* A, const double* w, const double* c,
                         const double* X, const int* e, const uint8_t* p,
                         void* cd)
{
  (void)w; (void)p;
  const orc_tables tv = entity_view((const orc_tables*)cd, e);
  const orc_tables* t = &tv;
  const int bs = t->bs, n = t->nd * bs, gd = t->gdim;
  const double mu = c[0], lmbda = c[1];
  double K[9], detJ = 0, g[3 * 64];
  const int affine = is_affine(t);
  for (int q = 0; q < t->nq; ++q)
  {
    GEOM_AT(q)
    phys_grads(t, q, K, g);
    const double s = t->weights[q] * fabs(detJ);
    for (int i = 0; i < t->nd; ++i)
      for (int j = 0; j < t->nd; ++j)
      {
        double dot = 0;
        for (int k = 0; k < gd; ++k) dot += g[i * 3 + k] * g[j * 3 + k];
        for (int a = 0; a < bs; ++a)
          for (int b = 0; b < bs; ++b)
          {
            double v = mu * g[i * 3 + b] * g[j * 3 + a]
                       + lmbda * g[i * 3 + a] * g[j * 3 + b];
            if (a == b) v += mu * dot;
            A[(i * bs + a) * n + j * bs + b] += s * v;
          }
      }
  }
}

/* inner(f, v) dx with f in the same (blocked) space; w = f at the cell dofs */
static void k_source(double* b, const double* w, const double* c,
                     const double* X, const int* e, const uint8_t* p, void* cd)
{
  (void)p;
  const orc_tables tv = entity_view((const orc_tables*)cd, e);
  const orc_tables* t = &tv;
  const int bs = t->bs;
  double K[9], detJ = 0;
  const int affine = is_affine(t);
  for (int q = 0; q < t->nq; ++q)
  {
    GEOM_AT(q)
    const double s = c[0] * t->weights[q] * fabs(detJ);
    for (int a = 0; a < bs; ++a)
    {
      double fq = 0;
      for (int j = 0; j < t->nd; ++j) fq += t->phi[q * t->nd + j] * w[j * bs + a];
      for (int i = 0; i < t->nd; ++i)
        b[i * bs + a] += s * fq * t->phi[q * t->nd + i];
    }
  }
}

/* The off-diagonal blocks of a Taylor-Hood Stokes system (python/tests/test_rectangular_assembly.py:83-86:
 * a01 = -inner(p, div(v)) dx, a10 = -inner(div(u), q) dx).  V = vector element (nv scalar basis functions x gdim
 * components, interleaved), Q = scalar element (ns basis functions); A_e row-major [n0][n1] as the UFCx ABI has it. */
static void div_coupling(double* A, const double* c, const double* X, const orc_tables* t, int v_is_test)
{
  const int nv = v_is_test ? t->nd : t->nd1, ns = v_is_test ? t->nd1 : t->nd, gd = t->gdim, td = t->tdim;
  const int n1 = v_is_test ? ns : nv * gd;
  double K[9], detJ = 0, g[3 * 64];
  for (int q = 0; q < t->nq; ++q)
  {
    if (jacobian(t, q, X, K, &detJ)) return;
    const double* dphv = v_is_test ? t->dphi + (size_t)q * td * nv : t->dphi1 + (size_t)q * td * nv;
    const double* phs = v_is_test ? t->phi1 + (size_t)q * ns : t->phi + (size_t)q * ns;
    for (int i = 0; i < nv; ++i)
      for (int k = 0; k < gd; ++k)
      {
        double sum = 0;
        for (int a = 0; a < td; ++a) sum += K[a * 3 + k] * dphv[a * nv + i];
        g[i * 3 + k] = sum;
      }
    const double s = c[0] * t->weights[q] * fabs(detJ);
    for (int i = 0; i < nv; ++i)
      for (int a = 0; a < gd; ++a)
        for (int j = 0; j < ns; ++j)
        {
          const double v = s * phs[j] * g[i * 3 + a];
          if (v_is_test) A[(i * gd + a) * n1 + j] += v;
          else A[j * n1 + i * gd + a] += v;
        }
  }
}
static void k_div_test(double* A, const double* w, const double* c, const double* X, const int* e, const uint8_t* p,
                       void* cd)
{
  (void)w; (void)p; (void)e;
  div_coupling(A, c, X, (const orc_tables*)cd, 1);
}
static void k_div_trial(double* A, const double* w, const double* c, const double* X, const int* e, const uint8_t* p,
                        void* cd)
{
  (void)w; (void)p; (void)e;
  div_coupling(A, c, X, (const orc_tables*)cd, 0);
}

/* A tabulate_tensor function handed in by the caller (UFCx signature without custom_data) -- what the reference
 * receives from a.kernel(...) (cpp/assemble_matrix.cpp:438-439, 620-636).  Process-global: the tests set it before
 * assembling a form of kind K_CUSTOM. */
typedef void (*orc_custom_fn)(double* A, const double* w, const double* c, const double* coordinate_dofs,
                              const int* entity_local_index, const uint8_t* quadrature_permutation);
static orc_custom_fn g_custom_fn = NULL;
void orc_set_custom_kernel(void* fn) { g_custom_fn = (orc_custom_fn)fn; }
static void k_custom(double* A, const double* w, const double* c, const double* X, const int* e, const uint8_t* p, void* cd)
{
  (void)cd;
  if (g_custom_fn) g_custom_fn(A, w, c, X, e, p);
}

static ufcx_kernel pick_kernel(int id)
{
  switch (id)
  {
  case K_CUSTOM: return k_custom;
  case K_LAPLACE: return k_laplace;
  case K_MASS: return k_mass;
  case K_ELASTICITY: return k_elasticity;
  case K_SOURCE: return k_source;
  case K_LAPLACE_VARCOEF: return k_laplace_varcoef;
  case K_DIV_TEST: return k_div_test;
  case K_DIV_TRIAL: return k_div_trial;
  default: return NULL;
  }
}

/* Tabulate one element tensor (used by tests to pin the kernels on their own). */
int orc_tabulate(int kernel, const orc_tables* t, const double* w,
                 const double* c, const double* X, double* out, int out_size)
{
  ufcx_kernel k = pick_kernel(kernel);
  if (!k) return ORC_ERR_KERNEL;
  memset(out, 0, sizeof(double) * (size_t)out_size);
  k(out, w, c, X, NULL, NULL, (void*)t);
  return ORC_OK;
}

/* ------------------------------------------------------------------------
 * Constraint data: the packed arrays of cpp/MultiPointConstraint.h:201-223. */
typedef struct
{
  const int8_t* is_slave;       /* [num_dofs] */
  const int32_t* masters;       /* dof-indexed adjacency values (local ids) */
  const double* coeffs;         /* same offsets */
  const int32_t* offsets;       /* [num_dofs + 1] */
  const int32_t* c2s;           /* cell_to_slaves values */
  const int32_t* c2s_offsets;   /* [num_cells + 1] */
  const int32_t* slaves;        /* sorted, owned first */
  int32_t num_slaves, num_local_slaves;
} orc_mpc;

typedef struct
{
  const int64_t* row_ptr;
  const int32_t* col;
  double* val;
  int64_t num_rows;
} orc_csr;

/* PETSc MatSetValuesLocal(ADD_VALUES) stand-in (python/src/dolfinx_mpc/mpc.cpp:
 * 285-286): per-entry search of the sorted row segment, then +=. */
static int mat_add(orc_csr* A, int nr, const int32_t* rows, int nc,
                   const int32_t* cols, const double* vals)
{
  for (int i = 0; i < nr; ++i)
  {
    const int64_t lo0 = A->row_ptr[rows[i]], hi0 = A->row_ptr[rows[i] + 1];
    for (int j = 0; j < nc; ++j)
    {
      int64_t lo = lo0, hi = hi0;
      const int32_t cj = cols[j];
      while (lo < hi)
      {
        const int64_t mid = (lo + hi) >> 1;
        if (A->col[mid] < cj) lo = mid + 1; else hi = mid;
      }
      if (lo >= hi0 || A->col[lo] != cj) return ORC_ERR_PATTERN;
      A->val[lo] += vals[i * nc + j];
    }
  }
  return ORC_OK;
}

/* MatSetValuesBlockedLocal stand-in: blocked indices, (nd0*bs0)x(nd1*bs1) block */
static int mat_add_block(orc_csr* A, int nd0, const int32_t* d0, int bs0, int nd1,
                         const int32_t* d1, int bs1, const double* Ae,
                         int32_t* rows, int32_t* cols)
{
  for (int i = 0; i < nd0; ++i)
    for (int k = 0; k < bs0; ++k) rows[i * bs0 + k] = d0[i] * bs0 + k;
  for (int j = 0; j < nd1; ++j)
    for (int k = 0; k < bs1; ++k) cols[j * bs1 + k] = d1[j] * bs1 + k;
  return mat_add(A, nd0 * bs0, rows, nd1 * bs1, cols, Ae);
}

/* cpp/assemble_utils.cpp:10-28 (returns a fresh heap vector, like the reference) */
static int32_t* compute_local_slave_index(const int32_t* slaves, int ns,
                                          int num_dofs, int bs,
                                          const int32_t* cell_dofs,
                                          const int8_t* is_slave)
{
  int32_t* local_index = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ns ? ns : 1));
  for (int i = 0; i < num_dofs; ++i)
    for (int j = 0; j < bs; ++j)
    {
      const int32_t dof = cell_dofs[i] * bs + j;
      if (is_slave[dof])
      {
        int it = 0;
        while (it < ns && slaves[it] != dof) ++it; /* std::ranges::find */
        if (it < ns) local_index[it] = i * bs + j;
      }
    }
  return local_index;
}

/* cpp/assemble_matrix.cpp:33-77 */
static void fill_stripped_matrix(double* S, const double* Ae, int nd0, int nd1,
                                 int bs0, int bs1, const int8_t* sl0,
                                 const int8_t* sl1, const int32_t* d0,
                                 const int32_t* d1)
{
  const int ndim1 = nd1 * bs1;
  for (int i = 0; i < nd0; ++i)
    for (int r = 0; r < bs0; ++r)
    {
      const int slave_row = sl0[d0[i] * bs0 + r];
      const int l_row = i * bs0 + r;
      for (int j = 0; j < nd1; ++j)
        for (int cc = 0; cc < bs1; ++cc)
        {
          const int slave_col = sl1[d1[j] * bs1 + cc];
          const int l_col = j * bs1 + cc;
          S[l_row * ndim1 + l_col]
              = (slave_row && slave_col) ? 0.0 : Ae[l_row * ndim1 + l_col];
        }
    }
}

/* cpp/assemble_matrix.cpp:99-268 */
static int modify_mpc_cell(orc_csr* A, int nd0, int nd1, double* Ae,
                           const int32_t* d0, const int32_t* d1, int bs0, int bs1,
                           const int32_t* sl0, int ns0, const int32_t* sl1,
                           int ns1, const orc_mpc* m0, const orc_mpc* m1,
                           double* scratch)
{
  const int nd[2] = {nd0, nd1}, bs[2] = {bs0, bs1}, ns[2] = {ns0, ns1};
  const int32_t* dofs[2] = {d0, d1};
  const int32_t* slaves[2] = {sl0, sl1};
  const orc_mpc* mpc[2] = {m0, m1};
  size_t nflat[2] = {0, 0};
  int32_t* local_index[2];
  for (int ax = 0; ax < 2; ++ax) /* :121-139 */
  {
    local_index[ax] = compute_local_slave_index(
        slaves[ax], ns[ax], nd[ax], bs[ax], dofs[ax], mpc[ax]->is_slave);
    for (int i = 0; i < nd[ax]; ++i)
      for (int j = 0; j < bs[ax]; ++j)
      {
        const int32_t dof = dofs[ax][i] * bs[ax] + j;
        if (mpc[ax]->is_slave[dof])
          nflat[ax] += (size_t)(mpc[ax]->offsets[dof + 1] - mpc[ax]->offsets[dof]);
      }
  }
  const int ndim0 = bs0 * nd0, ndim1 = bs1 * nd1;
  memset(scratch, 0, sizeof(double) * (size_t)(2 * ndim0 * ndim1 + ndim0 + ndim1));
  double* Ae_original = scratch;                       /* :147-153 */
  memcpy(Ae_original, Ae, sizeof(double) * (size_t)(ndim0 * ndim1));
  double* Ae_stripped = scratch + ndim0 * ndim1;       /* :155-161 */
  fill_stripped_matrix(Ae_stripped, Ae, nd0, nd1, bs0, bs1, m0->is_slave,
                       m1->is_slave, d0, d1);
  for (int i = 0; i < ns0; ++i)                        /* :165-171 */
    memset(Ae + ndim1 * local_index[0][i], 0, sizeof(double) * (size_t)ndim1);
  for (int i = 0; i < ns1; ++i)                        /* :173-178 */
    for (int r = 0; r < ndim0; ++r) Ae[r * ndim1 + local_index[1][i]] = 0.0;

  /* :182-201 flatten (three heap vectors per axis, as the reference) */
  int32_t* fm[2]; int32_t* fs[2]; double* fc[2];
  for (int ax = 0; ax < 2; ++ax)
  {
    fm[ax] = (int32_t*)malloc(sizeof(int32_t) * (nflat[ax] ? nflat[ax] : 1));
    fs[ax] = (int32_t*)malloc(sizeof(int32_t) * (nflat[ax] ? nflat[ax] : 1));
    fc[ax] = (double*)malloc(sizeof(double) * (nflat[ax] ? nflat[ax] : 1));
    size_t k = 0;
    for (int i = 0; i < ns[ax]; ++i)
    {
      const int32_t s = slaves[ax][i];
      for (int32_t j = mpc[ax]->offsets[s]; j < mpc[ax]->offsets[s + 1]; ++j)
      {
        fs[ax][k] = local_index[ax][i];
        fm[ax][k] = mpc[ax]->masters[j];
        fc[ax][k] = mpc[ax]->coeffs[j];
        ++k;
      }
    }
  }
  int err = ORC_OK;
  double* Arow = scratch + 2 * ndim0 * ndim1;          /* :209-210 */
  double* Acol = scratch + 2 * ndim0 * ndim1 + ndim0;
  int32_t* unrolled = (int32_t*)malloc(
      sizeof(int32_t) * (size_t)(ndim0 > ndim1 ? ndim0 : ndim1)); /* :213 */
  for (size_t i = 0; i < nflat[0] && !err; ++i)        /* :214-246 */
  {
    const double coeff_i = fc[0][i]; /* real T: no conj */
    for (int j = 0; j < nd1; ++j)
      for (int k = 0; k < bs1; ++k)
      {
        Acol[j * bs1 + k] = coeff_i * Ae_stripped[fs[0][i] * ndim1 + j * bs1 + k];
        unrolled[j * bs1 + k] = d1[j] * bs1 + k;
      }
    const int32_t row = fm[0][i];
    err = mat_add(A, 1, &row, ndim1, unrolled, Acol);
    for (size_t j = 0; j < nflat[1] && !err; ++j)
    {
      const int32_t col = fm[1][j];
      const double A0 = coeff_i * fc[1][j] * Ae_original[fs[0][i] * ndim1 + fs[1][j]];
      err = mat_add(A, 1, &row, 1, &col, &A0);
    }
  }
  for (size_t i = 0; i < nflat[1] && !err; ++i)        /* :251-267 */
  {
    for (int j = 0; j < nd0; ++j)
      for (int k = 0; k < bs0; ++k)
      {
        Arow[j * bs0 + k] = fc[1][i] * Ae_stripped[(j * bs0 + k) * ndim1 + fs[1][i]];
        unrolled[j * bs0 + k] = d0[j] * bs0 + k;
      }
    const int32_t col = fm[1][i];
    err = mat_add(A, ndim0, unrolled, 1, &col, Arow);
  }
  free(unrolled);
  for (int ax = 0; ax < 2; ++ax)
  {
    free(fm[ax]); free(fs[ax]); free(fc[ax]); free(local_index[ax]);
  }
  return err;
}

typedef struct
{
  const double* x;          /* [num_nodes][3] */
  const int32_t* x_dofmap;  /* [num_cells][ng] */
  int32_t ng;
} orc_mesh;

typedef struct
{
  const int32_t* map; /* [num_cells][nd] blocked */
  int32_t nd, bs;
} orc_dofmap;

/* One cell integral: cpp/assemble_matrix.cpp:417-548 (assemble_cells_impl).
 * cells == NULL means 0..num_cells-1.  coeffs is indexed by position in the
 * active list (:505).  bc0 / bc1 may be NULL (= empty marker vectors :513,526). */
int orc_assemble_cells_matrix(int kernel, const orc_tables* tab, const orc_mesh* mesh,
                              const int32_t* cells, int64_t num_cells,
                              const double* coeffs, int cstride,
                              const double* constants, const orc_dofmap* dm0,
                              const orc_dofmap* dm1, const int8_t* bc0,
                              const int8_t* bc1, const orc_mpc* m0,
                              const orc_mpc* m1, orc_csr* A, const int32_t* local_facets)
{
  /* local_facets != NULL: exterior-facet integral over the (cells[i], local_facets[i]) pairs --
   * assemble_exterior_facets, cpp/assemble_matrix.cpp:271-415: same steps as the cell loop, the local facet index
   * handed to the kernel (:361-362) */
  ufcx_kernel fn = pick_kernel(kernel);
  if (!fn || kernel == K_SOURCE) return ORC_ERR_KERNEL;
  const int nd0 = dm0->nd, nd1 = dm1->nd, bs0 = dm0->bs, bs1 = dm1->bs;
  const int ndim0 = nd0 * bs0, ndim1 = nd1 * bs1, ng = mesh->ng;
  double* X = (double*)malloc(sizeof(double) * 3 * (size_t)ng);
  double* Ae = (double*)malloc(sizeof(double) * (size_t)(ndim0 * ndim1));
  double* scratch = (double*)malloc(
      sizeof(double) * (size_t)(2 * ndim0 * ndim1 + ndim0 + ndim1));
  int32_t* rows = (int32_t*)malloc(sizeof(int32_t) * (size_t)ndim0);
  int32_t* cols = (int32_t*)malloc(sizeof(int32_t) * (size_t)ndim1);
  int err = ORC_OK;
  for (int64_t index = 0; index < num_cells && !err; ++index)
  {
    const int32_t cell = cells ? cells[index] : (int32_t)index;
    const int32_t* xd = mesh->x_dofmap + (int64_t)cell * ng;
    for (int i = 0; i < ng; ++i)                       /* :495-501 */
      memcpy(X + 3 * i, mesh->x + 3 * (int64_t)xd[i], 3 * sizeof(double));
    memset(Ae, 0, sizeof(double) * (size_t)(ndim0 * ndim1)); /* :504 */
    const int lf = local_facets ? local_facets[index] : 0;
    fn(Ae, coeffs ? coeffs + index * cstride : NULL, constants, X, local_facets ? &lf : NULL, NULL,
       (void*)tab);                                    /* :505-506 */
    const int32_t* d0 = dm0->map + (int64_t)cell * nd0;
    const int32_t* d1 = dm1->map + (int64_t)cell * nd1;
    if (bc0)                                           /* :513-525 */
      for (int i = 0; i < nd0; ++i)
        for (int k = 0; k < bs0; ++k)
          if (bc0[bs0 * d0[i] + k])
            memset(Ae + ndim1 * (bs0 * i + k), 0, sizeof(double) * (size_t)ndim1);
    if (bc1)                                           /* :526-533 */
      for (int j = 0; j < nd1; ++j)
        for (int k = 0; k < bs1; ++k)
          if (bc1[bs1 * d1[j] + k])
            for (int l = 0; l < ndim0; ++l) Ae[l * ndim1 + bs1 * j + k] = 0;
    const int ns0 = m0->c2s_offsets[cell + 1] - m0->c2s_offsets[cell];
    const int ns1 = m1->c2s_offsets[cell + 1] - m1->c2s_offsets[cell];
    if (ns0 > 0 || ns1 > 0)                            /* :537-545 */
      err = modify_mpc_cell(A, nd0, nd1, Ae, d0, d1, bs0, bs1,
                            m0->c2s + m0->c2s_offsets[cell], ns0,
                            m1->c2s + m1->c2s_offsets[cell], ns1, m0, m1, scratch);
    if (!err)                                          /* :546 */
      err = mat_add_block(A, nd0, d0, bs0, nd1, d1, bs1, Ae, rows, cols);
  }
  free(X); free(Ae); free(scratch); free(rows); free(cols);
  return err;
}

/* Slave diagonal, cpp/assemble_matrix.cpp:711-724 (caller checks equal spaces);
 * also used for the Dirichlet diagonal of python/.../assemble_matrix.py:59-62. */
int orc_add_diagonal(orc_csr* A, const int32_t* dofs, int32_t n, double diagval)
{
  for (int32_t i = 0; i < n; ++i)
  {
    const int err = mat_add(A, 1, dofs + i, 1, dofs + i, &diagval);
    if (err) return err;
  }
  return ORC_OK;
}

/* cpp/assemble_vector.h:35-69 */
static void modify_mpc_vec(double* b, double* b_local, const double* b_local_copy,
                           const int32_t* dofs, int num_dofs, int bs,
                           const int32_t* slaves, int ns, const orc_mpc* m)
{
  int32_t* local_index
      = compute_local_slave_index(slaves, ns, num_dofs, bs, dofs, m->is_slave);
  for (int i = 0; i < ns; ++i)
    for (int32_t j = m->offsets[slaves[i]]; j < m->offsets[slaves[i] + 1]; ++j)
    {
      b[m->masters[j]] += m->coeffs[j] * b_local_copy[local_index[i]];
      b_local[local_index[i]] = 0; /* inside the master loop, :66 */
    }
  free(local_index);
}

/* cpp/assemble_vector.cpp:34-91 with the cell lambda of :163-185 */
int orc_assemble_cells_vector(int kernel, const orc_tables* tab, const orc_mesh* mesh,
                              const int32_t* cells, int64_t num_cells,
                              const double* coeffs, int cstride,
                              const double* constants, const orc_dofmap* dm,
                              const orc_mpc* m, double* b, const int32_t* local_facets)
{
  /* local_facets != NULL: exterior facets, cpp/assemble_vector.cpp:196-240 */
  ufcx_kernel fn = pick_kernel(kernel);
  if (!fn || (kernel != K_SOURCE && kernel != K_CUSTOM)) return ORC_ERR_KERNEL;
  const int nd = dm->nd, bs = dm->bs, n = nd * bs, ng = mesh->ng;
  double* X = (double*)malloc(sizeof(double) * 3 * (size_t)ng);
  double* be = (double*)malloc(sizeof(double) * (size_t)n);
  double* be_copy = (double*)malloc(sizeof(double) * (size_t)n);
  for (int64_t e = 0; e < num_cells; ++e)
  {
    const int32_t cell = cells ? cells[e] : (int32_t)e;
    const int32_t* xd = mesh->x_dofmap + (int64_t)cell * ng;
    for (int i = 0; i < ng; ++i)
      memcpy(X + 3 * i, mesh->x + 3 * (int64_t)xd[i], 3 * sizeof(double));
    memset(be, 0, sizeof(double) * (size_t)n);
    const int lf = local_facets ? local_facets[e] : 0;
    fn(be, coeffs ? coeffs + e * cstride : NULL, constants, X, local_facets ? &lf : NULL, NULL,
       (void*)tab);
    const int32_t* dofs = dm->map + (int64_t)cell * nd;
    const int ns = m->c2s_offsets[cell + 1] - m->c2s_offsets[cell];
    if (ns > 0)                                        /* :76-84 */
    {
      memcpy(be_copy, be, sizeof(double) * (size_t)n);
      modify_mpc_vec(b, be, be_copy, dofs, nd, bs, m->c2s + m->c2s_offsets[cell],
                     ns, m);
    }
    for (int i = 0; i < nd; ++i)                       /* :87-89 */
      for (int k = 0; k < bs; ++k) b[bs * dofs[i] + k] += be[bs * i + k];
  }
  free(X); free(be); free(be_copy);
  return ORC_OK;
}

/* dolfinx::fem::pack_coefficients as called at cpp/assemble_matrix.cpp:587-589 and
 * cpp/assemble_vector.cpp:113-116 (DOLFINx code, not in /root/reference): for every
 * active entity, in iteration order, the cell-local dof values of one coefficient are
 * copied into its column range [offset, offset + nd*bs) of the packed row. */
void orc_pack_coefficient(const double* u, const orc_dofmap* dm, const int32_t* cells,
                          int64_t num_cells, double* w, int cstride, int offset)
{
  const int nd = dm->nd, bs = dm->bs;
  for (int64_t e = 0; e < num_cells; ++e)
  {
    const int32_t cell = cells ? cells[e] : (int32_t)e;
    const int32_t* d = dm->map + (int64_t)cell * nd;
    double* row = w + e * cstride + offset;
    for (int i = 0; i < nd; ++i)
      for (int k = 0; k < bs; ++k) row[i * bs + k] = u[(int64_t)d[i] * bs + k];
  }
}

/* cpp/lifting.h:45-134 (lift_bc_entities) + :250-301 (lift_bcs_cell).
 * x0 may be NULL (= empty span, :295). */
int orc_apply_lifting_cells(int kernel, const orc_tables* tab, const orc_mesh* mesh,
                            const int32_t* cells, int64_t num_cells,
                            const double* coeffs, int cstride,
                            const double* constants, const orc_dofmap* dm0,
                            const orc_dofmap* dm1, const int8_t* bc_markers1,
                            const double* bc_values1, const double* x0,
                            double scale, const orc_mpc* m0, double* b,
                            const int32_t* local_facets)
{
  /* local_facets != NULL: exterior facets, cpp/lifting.h:316-397 (same per-entity steps) */
  ufcx_kernel fn = pick_kernel(kernel);
  if (!fn || kernel == K_SOURCE) return ORC_ERR_KERNEL;
  const int nd0 = dm0->nd, nd1 = dm1->nd, bs0 = dm0->bs, bs1 = dm1->bs;
  const int num_rows = nd0 * bs0, num_cols = nd1 * bs1, ng = mesh->ng;
  double* X = (double*)malloc(sizeof(double) * 3 * (size_t)ng);
  double* Ae = (double*)malloc(sizeof(double) * (size_t)(num_rows * num_cols));
  double* be = (double*)malloc(sizeof(double) * (size_t)num_rows);
  double* be_copy = (double*)malloc(sizeof(double) * (size_t)num_rows);
  for (int64_t e = 0; e < num_cells; ++e)
  {
    const int32_t cell = cells ? cells[e] : (int32_t)e;
    const int32_t* dmap0 = dm0->map + (int64_t)cell * nd0;
    const int32_t* dmap1 = dm1->map + (int64_t)cell * nd1;
    int has_bc = 0;                                    /* :93-109 */
    for (int j = 0; j < nd1 && !has_bc; ++j)
      for (int k = 0; k < bs1; ++k)
        if (bc_markers1[bs1 * dmap1[j] + k]) { has_bc = 1; break; }
    if (!has_bc) continue;
    const int32_t* xd = mesh->x_dofmap + (int64_t)cell * ng;
    for (int i = 0; i < ng; ++i)
      memcpy(X + 3 * i, mesh->x + 3 * (int64_t)xd[i], 3 * sizeof(double));
    memset(Ae, 0, sizeof(double) * (size_t)(num_rows * num_cols));
    const int lf = local_facets ? local_facets[e] : 0;
    fn(Ae, coeffs ? coeffs + e * cstride : NULL, constants, X, local_facets ? &lf : NULL, NULL,
       (void*)tab);
    memset(be, 0, sizeof(double) * (size_t)num_rows);  /* :276-299 */
    for (int j = 0; j < nd1; ++j)
      for (int k = 0; k < bs1; ++k)
      {
        const int32_t jj = bs1 * dmap1[j] + k;
        if (bc_markers1[jj])
        {
          const double bc = bc_values1[jj];
          const double _x0 = x0 ? x0[jj] : 0.0;
          for (int mrow = 0; mrow < num_rows; ++mrow)
            be[mrow] -= Ae[mrow * num_cols + bs1 * j + k] * scale * (bc - _x0);
        }
      }
    const int ns = m0->c2s_offsets[cell + 1] - m0->c2s_offsets[cell]; /* :117 */
    if (ns > 0)
    {
      memcpy(be_copy, be, sizeof(double) * (size_t)num_rows);
      modify_mpc_vec(b, be, be_copy, dmap0, nd0, bs0,
                     m0->c2s + m0->c2s_offsets[cell], ns, m0);
    }
    for (int i = 0; i < nd0; ++i)                      /* :130-132 */
      for (int k = 0; k < bs0; ++k) b[bs0 * dmap0[i] + k] += be[bs0 * i + k];
  }
  free(X); free(Ae); free(be); free(be_copy);
  return ORC_OK;
}

/* cpp/MultiPointConstraint.h:129-145 / :148-152 */
void orc_backsubstitution(const orc_mpc* m, double* v)
{
  for (int32_t i = 0; i < m->num_slaves; ++i)
  {
    const int32_t s = m->slaves[i];
    v[s] = 0.0;
    for (int32_t k = m->offsets[s]; k < m->offsets[s + 1]; ++k)
      v[s] += m->coeffs[k] * v[m->masters[k]];
  }
}
void orc_homogenize(const orc_mpc* m, double* v)
{
  for (int32_t i = 0; i < m->num_slaves; ++i) v[m->slaves[i]] = 0.0;
}

/* ------------------------------------------------------------------------
 * Constraint construction: cpp/MultiPointConstraint.h:36-126 with
 * create_cell_to_dofs_map (cpp/mpc_helpers.h:19-94).  Masters arrive already in
 * local (extended-map) numbering: the global->local step (:117-125) needs the
 * DOLFINx IndexMap and is done by the caller.
 * Outputs (caller-allocated): is_slave[num_dofs], m_offsets[num_dofs+1],
 * m_masters/m_coeffs/m_owners[offsets_in[ns]], sorted_slaves[ns],
 * c2s_offsets[num_cells+1]; c2s is malloc'ed here (*c2s_out, free with orc_free). */
int orc_mpc_build(int32_t num_dofs, int32_t num_local_dofs, int32_t ns,
                  const int32_t* slaves, const int32_t* masters_local,
                  const double* coeffs, const int32_t* owners,
                  const int32_t* offsets_in, const int32_t* dofmap,
                  int32_t num_cells, int32_t nd, int32_t bs, int8_t* is_slave,
                  int32_t* m_offsets, int32_t* m_masters, double* m_coeffs,
                  int32_t* m_owners, int32_t* sorted_slaves,
                  int32_t* num_local_slaves, int32_t* c2s_offsets,
                  int32_t** c2s_out)
{
  memset(is_slave, 0, (size_t)num_dofs);
  for (int32_t i = 0; i < ns; ++i) is_slave[slaves[i]] = 1;       /* :55-63 */
  int32_t* num_masters = (int32_t*)calloc((size_t)num_dofs + 1, sizeof(int32_t));
  for (int32_t i = 0; i < ns; ++i)                                 /* :70-72 */
    num_masters[slaves[i]] = offsets_in[i + 1] - offsets_in[i];
  m_offsets[0] = 0;                                                /* :73-77 */
  for (int32_t d = 0; d < num_dofs; ++d) m_offsets[d + 1] = m_offsets[d] + num_masters[d];
  memset(num_masters, 0, sizeof(int32_t) * (size_t)num_dofs);
  for (int32_t i = 0; i < ns; ++i)                                 /* :86-98 */
    for (int32_t j = 0; j < offsets_in[i + 1] - offsets_in[i]; ++j)
    {
      const int32_t p = m_offsets[slaves[i]] + num_masters[slaves[i]];
      m_masters[p] = masters_local[offsets_in[i] + j];
      m_coeffs[p] = coeffs[offsets_in[i] + j];
      m_owners[p] = owners[offsets_in[i] + j];
      num_masters[slaves[i]]++;
    }
  free(num_masters);
  int32_t c = 0;                                                   /* :105-110 */
  for (int32_t d = 0; d < num_dofs; ++d)
    if (is_slave[d]) sorted_slaves[c++] = d;
  int32_t nl = 0;                                                  /* :112-115 */
  while (nl < c && sorted_slaves[nl] < num_local_dofs) ++nl;
  *num_local_slaves = nl;

  /* create_cell_to_dofs_map, mpc_helpers.h:40-93: dof -> cells for slave dofs,
   * then inverted; yields slaves ascending within each cell */
  int32_t* in_num_cells = (int32_t*)calloc((size_t)num_dofs, sizeof(int32_t));
  for (int32_t i = 0; i < num_cells; ++i)
    for (int32_t k = 0; k < nd; ++k)
      for (int32_t j = 0; j < bs; ++j) in_num_cells[dofmap[(int64_t)i * nd + k] * bs + j]++;
  int32_t* num_slave_cells = (int32_t*)calloc((size_t)num_dofs, sizeof(int32_t));
  for (int32_t i = 0; i < ns; ++i) num_slave_cells[slaves[i]] = in_num_cells[slaves[i]];
  free(in_num_cells);
  int64_t* cell_offsets = (int64_t*)malloc(sizeof(int64_t) * ((size_t)num_dofs + 1));
  cell_offsets[0] = 0;
  for (int32_t d = 0; d < num_dofs; ++d) cell_offsets[d + 1] = cell_offsets[d] + num_slave_cells[d];
  int32_t* cell_data = (int32_t*)malloc(sizeof(int32_t) * (size_t)(cell_offsets[num_dofs] + 1));
  int32_t* insert_position = (int32_t*)calloc((size_t)num_dofs, sizeof(int32_t));
  for (int32_t i = 0; i < num_cells; ++i)
    for (int32_t k = 0; k < nd; ++k)
      for (int32_t j = 0; j < bs; ++j)
      {
        const int32_t dof = dofmap[(int64_t)i * nd + k] * bs + j;
        if (num_slave_cells[dof] > 0) cell_data[cell_offsets[dof] + insert_position[dof]++] = i;
      }
  free(insert_position);
  memset(c2s_offsets, 0, sizeof(int32_t) * ((size_t)num_cells + 1));
  for (int32_t d = 0; d < num_dofs; ++d)
    for (int64_t k = cell_offsets[d]; k < cell_offsets[d + 1]; ++k) c2s_offsets[cell_data[k] + 1]++;
  for (int32_t i = 0; i < num_cells; ++i) c2s_offsets[i + 1] += c2s_offsets[i];
  int32_t* c2s = (int32_t*)malloc(sizeof(int32_t) * (size_t)(c2s_offsets[num_cells] + 1));
  int32_t* pos = (int32_t*)calloc((size_t)num_cells, sizeof(int32_t));
  for (int32_t d = 0; d < num_dofs; ++d)
    for (int64_t k = cell_offsets[d]; k < cell_offsets[d + 1]; ++k)
    {
      const int32_t cell = cell_data[k];
      c2s[c2s_offsets[cell] + pos[cell]++] = d;
    }
  free(pos); free(cell_data); free(cell_offsets); free(num_slave_cells);
  *c2s_out = c2s;
  return ORC_OK;
}

void orc_free(void* p) { free(p); }

/* ------------------------------------------------------------------------
 * Sparsity pattern: cpp/utils.h:381-496 (create_sparsity_pattern) with the
 * standard cell pattern of :276-361.  Block pairs are collected, sorted and
 * made unique (what dolfinx::la::SparsityPattern::finalize does), then
 * expanded by bs0 x bs1 to a scalar CSR with ascending columns. */
typedef struct { int64_t* d; size_t n, cap; } keyvec;
static int kv_push(keyvec* v, int64_t k)
{
  if (v->n == v->cap)
  {
    size_t nc = v->cap ? 2 * v->cap : 1024;
    int64_t* nd = (int64_t*)realloc(v->d, sizeof(int64_t) * nc);
    if (!nd) return 1;
    v->d = nd; v->cap = nc;
  }
  v->d[v->n++] = k;
  return 0;
}
static int cmp_i64(const void* a, const void* b)
{
  const int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
  return (x > y) - (x < y);
}

/* Returns the scalar CSR through malloc'ed *row_ptr_out / *col_out. */
int orc_create_pattern(const orc_dofmap* dm0, const orc_dofmap* dm1,
                       const int32_t* cells, int64_t num_active,
                       int32_t num_cells, const orc_mpc* m0, const orc_mpc* m1,
                       int64_t num_rows, int64_t** row_ptr_out, int32_t** col_out)
{
  const int nd0 = dm0->nd, nd1 = dm1->nd, bs0 = dm0->bs, bs1 = dm1->bs;
  const int64_t nbc = ((int64_t)1) << 32;
  keyvec kv = {NULL, 0, 0};
  int err = 0;
  /* standard pattern, utils.h:318-327 (sparsitybuild::cells) */
  for (int64_t e = 0; e < num_active && !err; ++e)
  {
    const int32_t c = cells ? cells[e] : (int32_t)e;
    for (int i = 0; i < nd0; ++i)
      for (int j = 0; j < nd1; ++j)
        err |= kv_push(&kv, dm0->map[(int64_t)c * nd0 + i] * nbc + dm1->map[(int64_t)c * nd1 + j]);
  }
  /* MPC additions, utils.h:456-489 (pattern_populator), every owned cell */
  int32_t* colset = NULL; size_t colcap = 0;
  for (int32_t c = 0; c < num_cells && !err; ++c)
  {
    size_t need = (size_t)nd1;
    for (int32_t k = m1->c2s_offsets[c]; k < m1->c2s_offsets[c + 1]; ++k)
      need += (size_t)(m1->offsets[m1->c2s[k] + 1] - m1->offsets[m1->c2s[k]]);
    if (need > colcap) { colcap = 2 * need; colset = (int32_t*)realloc(colset, sizeof(int32_t) * colcap); }
    size_t ncol = 0;
    for (int j = 0; j < nd1; ++j) colset[ncol++] = dm1->map[(int64_t)c * nd1 + j];
    for (int32_t k = m1->c2s_offsets[c]; k < m1->c2s_offsets[c + 1]; ++k)   /* :427-428 */
      for (int32_t q = m1->offsets[m1->c2s[k]]; q < m1->offsets[m1->c2s[k] + 1]; ++q)
        colset[ncol++] = m1->masters[q] / bs1;
    for (int i = 0; i < nd0; ++i)                                          /* :471 */
      for (size_t j = 0; j < ncol; ++j)
        err |= kv_push(&kv, dm0->map[(int64_t)c * nd0 + i] * nbc + colset[j]);
    for (int32_t k = m0->c2s_offsets[c]; k < m0->c2s_offsets[c + 1]; ++k)   /* :481-488 */
      for (int32_t q = m0->offsets[m0->c2s[k]]; q < m0->offsets[m0->c2s[k] + 1]; ++q)
        for (size_t j = 0; j < ncol; ++j)
          err |= kv_push(&kv, (int64_t)(m0->masters[q] / bs0) * nbc + colset[j]);
  }
  free(colset);
  if (err) { free(kv.d); return ORC_ERR_ALLOC; }
  qsort(kv.d, kv.n, sizeof(int64_t), cmp_i64);
  size_t nu = 0;
  for (size_t i = 0; i < kv.n; ++i)
    if (i == 0 || kv.d[i] != kv.d[i - 1]) kv.d[nu++] = kv.d[i];
  /* expand blocks */
  int64_t* row_ptr = (int64_t*)calloc((size_t)num_rows + 1, sizeof(int64_t));
  for (size_t i = 0; i < nu; ++i)
  {
    const int64_t br = kv.d[i] / nbc;
    for (int k = 0; k < bs0; ++k) row_ptr[br * bs0 + k + 1] += bs1;
  }
  for (int64_t r = 0; r < num_rows; ++r) row_ptr[r + 1] += row_ptr[r];
  int32_t* col = (int32_t*)malloc(sizeof(int32_t) * (size_t)(row_ptr[num_rows] + 1));
  int64_t* fill = (int64_t*)malloc(sizeof(int64_t) * ((size_t)num_rows + 1));
  memcpy(fill, row_ptr, sizeof(int64_t) * ((size_t)num_rows + 1));
  for (size_t i = 0; i < nu; ++i)
  {
    const int64_t br = kv.d[i] / nbc, bc = kv.d[i] % nbc;
    for (int k = 0; k < bs0; ++k)
      for (int l = 0; l < bs1; ++l) col[fill[br * bs0 + k]++] = (int32_t)(bc * bs1 + l);
  }
  free(fill); free(kv.d);
  *row_ptr_out = row_ptr; *col_out = col;
  return ORC_OK;
}
