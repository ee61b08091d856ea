/*
 * mpcx.h -- C ABI of the B200-native MPC-constrained assembly engine (libmpcx.so).
 *
 * This is the drop-in boundary for the hot path of jorgensd/dolfinx_mpc: the
 * entry points below are what the reference's FFI layer for this path
 * (nanobind module dolfinx_mpc.cpp.mpc, python/src/dolfinx_mpc/mpc.cpp:261-345)
 * would bind instead of its own C++ loops.  Plain pointers and sizes only; no
 * torch / DOLFINx / PETSc types.  The array layouts are the ones the reference
 * states explicitly in its numba path (python/src/dolfinx_mpc/numba/
 * assemble_matrix.py:58-104): blocked int32 dofmaps, 3-padded float64 geometry,
 * dof-indexed master/coefficient adjacency lists, int8 slave / bc markers.
 *
 * Conventions
 *  - every function returns 0 (MPCX_OK) or an mpcx_status; mpcx_last_error()
 *    gives the message (thread-local).  Reference: C++ std::runtime_error
 *    surfaced by nanobind (cpp/assemble_matrix.cpp:315,464,605,659).
 *  - pointers inside the structs are DEVICE pointers unless the name ends in
 *    _host.  They are borrowed for the call.  Outputs are written in place.
 *  - `stream` is a cudaStream_t passed as void*; kernels are asynchronous.
 *  - a slave/entry outside the sparsity pattern raises a device-side flag read
 *    back by mpcx_device_error() (one 4-byte copy + stream sync).
 *  - T = float64 (suffix _f64).  Integer maps int32, row_ptr int64.
 */
#ifndef MPCX_H
#define MPCX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MPCX_ABI_VERSION 6
#define MPCX_MAX_CONSTANTS 8

typedef enum mpcx_status
{
  MPCX_OK = 0,
  MPCX_ERR_ARG = 1,          /* bad argument (null pointer, negative size ...) */
  MPCX_ERR_UNSUPPORTED = 2,  /* kernel/element combination without a device kernel */
  MPCX_ERR_CUDA = 3,         /* CUDA runtime error */
  MPCX_ERR_PATTERN = 4,      /* an insertion fell outside the sparsity pattern */
  MPCX_ERR_ALLOC = 5
} mpcx_status;

/* Element kernels: replace the opaque FFCx tabulate_tensor function pointer
 * (cpp/assemble_matrix.cpp:438-439,505-506; numba/assemble_matrix.py:138-145). */
typedef enum mpcx_kernel
{
  MPCX_KERNEL_LAPLACE = 0,         /* c[0] * inner(grad u, grad v) dx, block-diagonal for bs > 1 */
  MPCX_KERNEL_MASS = 1,            /* c[0] * inner(u, v) dx */
  MPCX_KERNEL_ELASTICITY = 2,      /* inner(sigma(u), grad v) dx; c = {mu, lambda}; bs == gdim */
  MPCX_KERNEL_SOURCE = 3,          /* c[0] * inner(f, v) dx (linear form); w = f at the cell dofs */
  MPCX_KERNEL_LAPLACE_VARCOEF = 4, /* c[0] * w * inner(grad u, grad v) dx; w scalar, same element */
  MPCX_KERNEL_DIV_TEST = 5,        /* c[0] * inner(p, div v) dx; test = vector element (bs == gdim), trial = scalar */
  MPCX_KERNEL_DIV_TRIAL = 6,       /* c[0] * inner(div u, q) dx; test = scalar element, trial = vector (bs1 == gdim) */
  MPCX_KERNEL_CUSTOM = 7           /* mpcx_integral.custom: a tabulate_tensor function compiled at run time (below) */
} mpcx_kernel;

/* The reference's opaque element kernel (`fn` of cpp/assemble_matrix.cpp:438-439, called at :505-506 and
 * cpp/assemble_vector.cpp / cpp/lifting.h alike) for forms outside the registry above.  `source` is C / CUDA source
 * that defines
 *     void <entry>(double* A, const double* w, const double* c, const double* coordinate_dofs,
 *                  const int* entity_local_index, const uint8_t* quadrature_permutation);
 * -- the UFCx tabulate_tensor signature, so FFCx output can be passed as generated (its #include lines are dropped;
 * functions without an execution space become device functions; `restrict` is accepted).  It is compiled once with
 * NVRTC for sm_100a; an assembly first evaluates it for every active entity (thread per entity: coordinate_dofs
 * [num_coordinate_dofs][3], w = the packed coefficients of the entity, A zeroed, num_entries values) into a scratch
 * array owned by the handle -- in chunks of entities bounded by MPCX_CUSTOM_SCRATCH_MB (default 4096) --, each chunk
 * followed by the generic kernels, which eliminate and scatter as for a registry kernel.  The NVRTC log of
 * a failed compilation is what mpcx_last_error() returns. */
typedef struct mpcx_custom_kernel mpcx_custom_kernel;
int mpcx_custom_kernel_create(const char* source, const char* entry, int32_t num_entries, int32_t num_coordinate_dofs,
                              int32_t num_coefficient_values, mpcx_custom_kernel** kernel_out);
void mpcx_custom_kernel_destroy(mpcx_custom_kernel* kernel);

/* Tabulated element (what FFCx bakes into the generated kernel). */
typedef struct mpcx_tables
{
  int32_t tdim, gdim, nd, ng, nq, bs;
  const double* weights; /* [nq] */
  const double* phi;     /* [nq][nd] */
  const double* dphi;    /* [nq][tdim][nd] */
  const double* gdphi;   /* [nq][tdim][ng] */
  /* exterior-facet integrals (cpp/assemble_matrix.cpp:271-415): the arrays above hold nfacets consecutive tables,
   * one per local facet (facet quadrature points mapped into the cell); facet_tangents [nfacets][tdim-1][tdim] are
   * the tangents of the reference facet map, the surface measure being |J t| (2-D) or |J t1 x J t2| (3-D).
   * nfacets == 0 for cell integrals. */
  int32_t nfacets;
  const double* facet_tangents;
  /* rectangular forms -- test and trial ELEMENTS differ (dofs[2], bs[2], num_dofs[2] of modify_mpc_cell,
   * cpp/assemble_matrix.cpp:99-117; python/tests/test_rectangular_assembly.py:83-86): nd / bs / phi / dphi above
   * describe the test element, nd1 / bs1 / phi1 / dphi1 the trial element at the same quadrature points.
   * nd1 == 0: one element on both sides. */
  int32_t nd1, bs1;
  const double* phi1;  /* [nq][nd1] */
  const double* dphi1; /* [nq][tdim][nd1] */
} mpcx_tables;

/* Geometry: mesh.geometry().x() / dofmaps().front() (cpp/assemble_matrix.cpp:465-470).
 * x_stride is 3 (the reference layout) or 4 (padded device mirror, 32-byte rows). */
typedef struct mpcx_mesh
{
  const double* x;
  const int32_t* x_dofmap; /* [num_cells][ng] */
  int64_t num_nodes;
  int32_t ng;
  int32_t x_stride;
} mpcx_mesh;

/* dofmap.map() / bs() (cpp/assemble_matrix.cpp:474-478). */
typedef struct mpcx_dofmap
{
  const int32_t* map; /* [num_cells][nd], blocked */
  int32_t nd, bs;
  int64_t num_dofs; /* unrolled, owned + ghost */
  /* unrolled dofs owned by this process (owned first, ghosts after: cpp/MultiPointConstraint.h:112-115); 0 = all.  The
   * tile plans order the cells that touch a ghost row first, so that the rows to be sent to their owners are complete
   * early (mpcx_assemble_system_tiled_part_f64). */
  int64_t num_owned_dofs;
} mpcx_dofmap;

/* MultiPointConstraint data (cpp/MultiPointConstraint.h:201-223; tuple `mpc_data`
 * of numba/assemble_matrix.py:60-74). */
typedef struct mpcx_mpc
{
  const int8_t* is_slave;          /* [num_dofs] */
  const int32_t* masters;          /* adjacency values, local (extended-map) dof ids */
  const double* coeffs;            /* same offsets */
  const int32_t* offsets;          /* [num_dofs + 1] */
  const int32_t* cell_to_slaves;   /* adjacency values (ascending slave dofs per cell) */
  const int32_t* cell_to_slaves_offsets; /* [num_cells + 1] */
  const int32_t* slaves;           /* sorted, first num_local_slaves are owned */
  int32_t num_slaves, num_local_slaves;
  int64_t num_dofs;
} mpcx_mpc;

/* Device CSR the element tensors accumulate into (replaces the PETSc Mat behind
 * mat_add_block_values / mat_add_values, python/src/dolfinx_mpc/mpc.cpp:284-287). */
typedef struct mpcx_csr
{
  const int64_t* row_ptr; /* [num_rows + 1] */
  const int32_t* col;     /* ascending within a row */
  double* val;
  int64_t num_rows, nnz;
} mpcx_csr;

/* One cell integral: kernel, integration domain, packed coefficients, constants
 * (cpp/assemble_matrix.cpp:620-636). */
typedef struct mpcx_integral
{
  int32_t kernel;
  const mpcx_tables* tables;  /* host struct holding device pointers */
  const int32_t* cells;       /* active cells or NULL for 0..num_cells-1 */
  int64_t num_cells;
  const double* coeffs;       /* packed [num_cells][cstride], by position (:505), or NULL */
  int32_t cstride;
  /* alternative to `coeffs`: gather w from a nodal array through a dofmap inside the
   * kernel (device-side pack_coefficients, cpp/assemble_matrix.cpp:587-589) */
  const double* coeff_nodal;
  const int32_t* coeff_dofmap; /* [num_cells][coeff_nd] */
  int32_t coeff_nd, coeff_bs;
  int32_t num_constants;
  double constants[MPCX_MAX_CONSTANTS];
  /* optional: compacted list of the active cells that hold a slave dof (row or column
   * side).  When given together with a scatter plan the bulk kernel skips those cells
   * and the elimination kernel runs on this list only. */
  const int32_t* slave_cells;      /* values are POSITIONS into the active list */
  int64_t num_slave_cells;
  /* exterior-facet integral: entity i is local facet local_facets[i] of cell cells[i] (the (cell, local facet)
   * pairs of cpp/assemble_matrix.cpp:343-348); NULL for a cell integral.  Needs tables->nfacets > 0. */
  const int32_t* local_facets;
  /* kernel == MPCX_KERNEL_CUSTOM: the compiled element kernel (NULL otherwise) */
  const mpcx_custom_kernel* custom;
} mpcx_integral;

/* Scatter plan: for every active cell and local entry (p, q) the offset of column
 * d1(q) inside CSR row d0(p) (blocked: offset of the block).  Built once per
 * (pattern, dofmaps) pair -- the role PETSc's per-call row search plays in the
 * reference, hoisted out of the assembly loop. */
typedef struct mpcx_plan
{
  const void* lpos;     /* [num_cells][nd0*nd1] uint8 or uint16 */
  int32_t width;        /* 1 or 2 bytes per entry */
} mpcx_plan;

const char* mpcx_last_error(void);
int mpcx_abi_version(void);

/* Instrumentation used by bench.py.  mpcx_launch_count: kernels launched by this library since load.
 * mpcx_profile_enable(1): every following mpcx_assemble_matrix_f64 call brackets its dominant (bulk) kernel
 * with CUDA events on the launch stream; mpcx_profile_read waits for them and returns the summed duration
 * in milliseconds and the number of bracketed launches, then forgets them. */
int mpcx_profile_enable(int on);
long long mpcx_launch_count(void);
int mpcx_profile_read(double* ms_sum, long long* n_timed);

/* Reads and clears the device-side error flag (syncs `stream`). 0 = none,
 * MPCX_ERR_PATTERN = an insertion missed the pattern. */
int mpcx_device_error(void* stream);
/* The same flag copied to pinned host memory in stream order, WITHOUT synchronising: the caller examines
 * *pinned_host_flag later (after an event / at its next call), so that a time loop keeps the device queue filled.  The
 * flag is not cleared. */
int mpcx_device_error_async(int32_t* pinned_host_flag, void* stream);

/* data[0 .. n) <- 0 in stream order (cudaMemsetAsync): `A.zeroEntries()` before an assembly
 * (python/src/dolfinx_mpc/assemble_matrix.py:51, problem.py:539) and `b_local.set(0.0)`
 * (python/src/dolfinx_mpc/assemble_vector.py:100-101) in the reference's callers. */
int mpcx_zero_f64(double* data, int64_t n, void* stream);

/* A += integral, with BC row/column zeroing and MPC elimination K^T A_e K.
 * Replaces assemble_cells_impl + modify_mpc_cell (cpp/assemble_matrix.cpp:417-548,
 * 99-268).  bc0 / bc1 may be NULL.  plan may be NULL (row search per entry). */
int mpcx_assemble_matrix_f64(const mpcx_integral* integral, const mpcx_mesh* mesh,
                             const mpcx_dofmap* dofmap0, const mpcx_dofmap* dofmap1,
                             const int8_t* bc0, const int8_t* bc1,
                             const mpcx_mpc* mpc0, const mpcx_mpc* mpc1,
                             const mpcx_csr* A, const mpcx_plan* plan, void* stream);

/* Tile plan for the cells without slave dofs (see csrc/mpcx_tile.cuh): cells ordered along a Morton curve, cut
 * into tiles of 512; per tile its vertices, CSR destinations and the element entries summing into each.  Built
 * once per (pattern, dofmaps, active cells, bc markers) on the device -- the role MatSetPreallocationCOO plays
 * for PETSc's device assembly.  `skip[i] != 0` excludes active cell i (cells holding slave dofs: the
 * elimination kernel handles them). */
typedef struct mpcx_tile_plan mpcx_tile_plan;
int mpcx_tile_plan_create(const mpcx_mesh* mesh, const mpcx_dofmap* dofmap0, const mpcx_dofmap* dofmap1,
                          const int32_t* cells, int64_t num_cells, const int8_t* skip,
                          const int8_t* bc0, const int8_t* bc1, const mpcx_csr* A, void* stream,
                          mpcx_tile_plan** plan_out);
void mpcx_tile_plan_destroy(mpcx_tile_plan* plan);
/* out[0..13] = tiles, cells per tile, bulk cells, max vertices / dest records per tile, total tile vertices,
 * total dest records, plan bytes read per assembly, max / total element-buffer slots, max / total runs
 * (= TMA bulk reductions per assembly), max staging positions per tile, 1 for a symmetric plan (same dofmap
 * and bc markers on both sides: upper-triangular records feed entry (r, c) and entry (c, r)); out[14] = number of
 * leading tiles that hold every cell touching a ghost row (0 without ghosts); out[15] = staging positions over all
 * tiles = fp64 additions the copy engine performs per assembly (entries + the zero padding inside runs) */
int mpcx_tile_plan_info(const mpcx_tile_plan* plan, int64_t* out, int32_t n);

/* Optional: scatter plan for the cells holding slaves (integral->slave_cells), stored inside a matrix tile plan of a
 * scalar P1 space.  For every insertion modify_mpc_cell makes for such a cell (cpp/assemble_matrix.cpp:214-267: master
 * rows, master columns, master x master, plus the unconstrained entries) it records the element entry, the CSR
 * position and the indices of the row / column coefficients, so that the elimination kernel of the tiled routines
 * needs neither constraint lookups nor row searches.  Coefficient VALUES are read at assembly time. */
int mpcx_tile_plan_add_slave_cells(mpcx_tile_plan* plan, const mpcx_integral* integral, const mpcx_dofmap* dofmap0,
                                   const mpcx_dofmap* dofmap1, const int8_t* bc0, const int8_t* bc1, const mpcx_mpc* mpc0,
                                   const mpcx_mpc* mpc1, const mpcx_csr* A, void* stream);

/* Same contract as mpcx_assemble_matrix_f64 (A += integral; the caller zeroes A), bulk cells through the
 * tile plan: element entries are combined per tile in shared memory and added with one reduction per
 * (tile, CSR entry), added to A.val by TMA bulk reductions over runs of consecutive entries; slave cells are
 * eliminated afterwards by the same kernel as in mpcx_assemble_matrix_f64.  A.val must be 16-byte aligned and
 * have room for nnz rounded up to an even count (a run may carry one zero of padding past the last entry). */
int mpcx_assemble_matrix_tiled_f64(const mpcx_integral* integral, const mpcx_mesh* mesh,
                                   const mpcx_dofmap* dofmap0, const mpcx_dofmap* dofmap1,
                                   const int8_t* bc0, const int8_t* bc1,
                                   const mpcx_mpc* mpc0, const mpcx_mpc* mpc1, const mpcx_csr* A,
                                   const mpcx_tile_plan* plan, void* stream);

/* Tile plan for the load vector (same tiling; dests are the row dofs, built once per (dofmap, active cells,
 * skip flags)) and the tiled counterpart of mpcx_assemble_vector_f64: element entries of the constraint-free
 * cells are combined per tile, one reduction per (tile, row); cells holding slaves (integral->slave_cells) go
 * through the elimination path of mpcx_assemble_vector_f64 (cpp/assemble_vector.h:35-69).  b is NOT zeroed; it
 * must be 16-byte aligned with room for num_dofs rounded up to an even count. */
int mpcx_vector_tile_plan_create(const mpcx_mesh* mesh, const mpcx_dofmap* dofmap, const int32_t* cells,
                                 int64_t num_cells, const int8_t* skip, void* stream, mpcx_tile_plan** plan_out);
int mpcx_assemble_vector_tiled_f64(const mpcx_integral* integral, const mpcx_mesh* mesh, const mpcx_dofmap* dofmap,
                                   const mpcx_mpc* mpc, double* b, const mpcx_tile_plan* plan, void* stream);

/* Matrix AND load vector of the same cells in ONE pass over the bulk cells (the assembly block of
 * LinearProblem.solve, python/src/dolfinx_mpc/problem.py:539-566, calls assemble_matrix and assemble_vector back to
 * back; each of them gathers the cell geometry again, cpp/assemble_matrix.cpp:488-506 and
 * cpp/assemble_vector.cpp:163-185).  Same contract as mpcx_assemble_matrix_tiled_f64 followed by
 * mpcx_assemble_vector_tiled_f64 on one space (test == trial == the space of L, one bc marker array, one
 * constraint): A += a_integral, b += L_integral; the caller zeroes A and b.  Both tile plans must have been created
 * for the same cells and skip flags (they then share one tiling).  Cells holding slaves go through the elimination
 * kernels of both routines. */
int mpcx_assemble_system_tiled_f64(const mpcx_integral* a_integral, const mpcx_integral* L_integral, const mpcx_mesh* mesh,
                                   const mpcx_dofmap* dofmap, const int8_t* bc, const mpcx_mpc* mpc, const mpcx_csr* A,
                                   double* b, const mpcx_tile_plan* matrix_plan, const mpcx_tile_plan* vector_plan,
                                   void* stream);

/* "Row gather" assembly of a blocked space (bs == gdim, P1 simplices, isotropic elasticity): one warp owns one block
 * row, evaluates the cells around its node and STORES the finished row -- no atomics and no zero-fill: unlike every other
 * routine here this one OVERWRITES A (all rows, zeros included) before the cells holding slaves are added by the
 * elimination kernel, so it must be the first integral assembled into A and the caller must not zero A.  Replaces
 * assemble_cells_impl + MatSetValuesBlockedLocal (cpp/assemble_matrix.cpp:417-548) for that element; the row plan
 * (cells around every node, and for every block entry the (cell, i, j) contributions) is built once per pattern /
 * dofmap / active cells / skip flags on the device.  One constraint and one bc marker array on both sides. */
typedef struct mpcx_row_plan mpcx_row_plan;
typedef struct mpcx_slave_plan mpcx_slave_plan;
int mpcx_row_plan_create(const mpcx_dofmap* dofmap, const int32_t* cells, int64_t num_cells, const int8_t* skip,
                         const mpcx_csr* A, void* stream, mpcx_row_plan** plan_out);
void mpcx_row_plan_destroy(mpcx_row_plan* plan);
int mpcx_assemble_matrix_rowgather_f64(const mpcx_integral* integral, const mpcx_mesh* mesh, const mpcx_dofmap* dofmap,
                                       const int8_t* bc, const mpcx_mpc* mpc, const mpcx_csr* A, const mpcx_row_plan* plan,
                                       const mpcx_slave_plan* slave_plan /* optional */, void* stream);

/* Scatter plan of the cells holding slaves for ANY element (the general form of mpcx_tile_plan_add_slave_cells): the
 * (element entry, CSR position, coefficient indices) of every insertion modify_mpc_cell makes for them
 * (cpp/assemble_matrix.cpp:214-267, explicit zeros of bc rows / columns left out), found once per pattern / constraint /
 * bc set.  mpcx_assemble_slave_cells_f64 tabulates the element matrix of every such cell and walks its list:
 * A += K^T A_e K for those cells only. */
int mpcx_slave_plan_create(const mpcx_integral* integral, const mpcx_dofmap* dofmap0, const mpcx_dofmap* dofmap1,
                           const int8_t* bc0, const int8_t* bc1, const mpcx_mpc* mpc0, const mpcx_mpc* mpc1,
                           const mpcx_csr* A, void* stream, mpcx_slave_plan** plan_out);
void mpcx_slave_plan_destroy(mpcx_slave_plan* plan);
int mpcx_assemble_slave_cells_f64(const mpcx_integral* integral, const mpcx_mesh* mesh, const mpcx_mpc* mpc0,
                                  const mpcx_mpc* mpc1, const mpcx_csr* A, const mpcx_slave_plan* plan, void* stream);

/* The same in two parts, for overlapping the ghost-row exchange with the assembly of the interior: part 1 = the cells
 * holding slaves and the tiles that touch ghost rows (tiles [0, interface_tiles) of the plans, see mpcx_tile_plan_info
 * out[14]) -- after it every ghost row of A is complete and can travel (mpcx_ghost_reduce_f64 on another stream);
 * part 2 = the remaining tiles; part 0 = everything (= mpcx_assemble_system_tiled_f64). */
int mpcx_assemble_system_tiled_part_f64(const mpcx_integral* a_integral, const mpcx_integral* L_integral, const mpcx_mesh* mesh,
                                        const mpcx_dofmap* dofmap, const int8_t* bc, const mpcx_mpc* mpc, const mpcx_csr* A,
                                        double* b, const mpcx_tile_plan* matrix_plan, const mpcx_tile_plan* vector_plan,
                                        int32_t part, void* stream);

/* A[d, d] += diagval for the listed unrolled dofs.  Slave diagonal
 * (cpp/assemble_matrix.cpp:711-724) and Dirichlet diagonal
 * (python/src/dolfinx_mpc/assemble_matrix.py:59-62). */
int mpcx_add_diagonal_f64(const mpcx_csr* A, const int32_t* dofs, int64_t n,
                          double diagval, void* stream);

/* Builds the scatter plan on the device (one row search per entry). */
int mpcx_build_plan(const mpcx_dofmap* dofmap0, const mpcx_dofmap* dofmap1,
                    const int32_t* cells, int64_t num_cells, const mpcx_csr* A,
                    void* lpos_out, int32_t width, void* stream);

/* b += integral with K^T b_e.  Replaces _assemble_entities_impl + modify_mpc_vec
 * (cpp/assemble_vector.cpp:34-91, cpp/assemble_vector.h:35-69).  b is NOT zeroed. */
int mpcx_assemble_vector_f64(const mpcx_integral* integral, const mpcx_mesh* mesh,
                             const mpcx_dofmap* dofmap, const mpcx_mpc* mpc,
                             double* b, void* stream);

/* b -= scale * K^T A_e (g - x0) on cells with a Dirichlet column.  Replaces
 * lift_bc_entities + lift_bcs_cell (cpp/lifting.h:45-134,250-301).  x0 may be NULL.
 * bc_cells (optional): positions in the active list of the cells to visit -- normally the
 * cells with a marked column dof, found once by mpcx_flag_cells; NULL = visit every active
 * cell and test it, as the reference does (cpp/lifting.h:93-109). */
int mpcx_apply_lifting_f64(const mpcx_integral* integral, const mpcx_mesh* mesh,
                           const mpcx_dofmap* dofmap0, const mpcx_dofmap* dofmap1,
                           const int8_t* bc_markers1, const double* bc_values1,
                           const double* x0, double scale, const mpcx_mpc* mpc0,
                           const int32_t* bc_cells, int64_t num_bc_cells,
                           double* b, void* stream);

/* flags_out[i] = 1 when active cell i (cells[i], or i when cells == NULL) has a dof d with
 * marker[d] != 0.  Setup helper for the list above. */
int mpcx_flag_cells(const mpcx_dofmap* dofmap, const int32_t* cells, int64_t num_cells,
                    const int8_t* marker, int8_t* flags_out, void* stream);

/* u[s] = sum_k coeff_k u[master_k] / u[s] = 0 for every slave.  Replaces
 * MultiPointConstraint::backsubstitution / homogenize (cpp/MultiPointConstraint.h:129-152). */
int mpcx_backsubstitution_f64(const mpcx_mpc* mpc, double* u, void* stream);
int mpcx_homogenize_f64(const mpcx_mpc* mpc, double* u, void* stream);

/* Ghost-row exchange helpers (PETSc MatAssembly / VecGhostUpdate(ADD, REVERSE) in the
 * reference's callers, python/src/dolfinx_mpc/assemble_matrix.py:60-64):
 * dst[i] = src[idx[i]]  and  dst[idx[i]] += src[i]. */
int mpcx_gather_f64(const double* src, const int64_t* idx, int64_t n, double* dst, void* stream);
int mpcx_scatter_add_f64(double* dst, const int64_t* idx, int64_t n, const double* src, void* stream);

/* The exchange step of the distributed path in ONE call over NCCL (PETSc MatAssemblyBegin/End and
 * VecGhostUpdate(ADD_VALUES, SCATTER_REVERSE) in the reference's callers, python/src/dolfinx_mpc/assemble_matrix.py:64,
 * python/src/dolfinx_mpc/problem.py:566): pack the ghost values (values[send_idx[i]], ordered by destination rank; or
 * the contiguous slice values[send_start ...] when send_idx == NULL -- ghost rows are numbered last), grouped
 * ncclSend / ncclRecv with every peer that has a non-zero count, then values[recv_pos[i]] += received[i] at the owner.
 * send_counts / recv_counts are HOST arrays of `world` entries; send_idx / recv_pos / buffers are device pointers.
 * NCCL is bound at run time: mpcx_nccl_load(path) (NULL: the libnccl.so.2 already loaded in the process, else the
 * default search path).  A communicator is created collectively from an id made on one rank (mpcx_comm_unique_id ->
 * broadcast by the caller -> mpcx_comm_create on every rank). */
typedef struct mpcx_comm mpcx_comm;
int mpcx_nccl_load(const char* libnccl_path);
int mpcx_comm_unique_id(void* id128_out);
int mpcx_comm_create(const void* id128, int32_t rank, int32_t world, mpcx_comm** comm_out);
void mpcx_comm_destroy(mpcx_comm* comm);
int mpcx_ghost_reduce_f64(mpcx_comm* comm, double* values, const int64_t* send_idx, int64_t send_start,
                          const int64_t* send_counts, const int64_t* recv_pos, const int64_t* recv_counts,
                          double* send_buf, double* recv_buf, void* stream);

/* Sparsity pattern with the MPC additions, on the HOST (cold path; replaces
 * create_sparsity_pattern, cpp/utils.h:381-496).  All pointers are host pointers.
 * The scalar CSR is returned in malloc'ed arrays released with mpcx_free_host. */
typedef struct mpcx_mpc_host
{
  const int32_t* masters;
  const int32_t* offsets;
  const int32_t* cell_to_slaves;
  const int32_t* cell_to_slaves_offsets;
} mpcx_mpc_host;

/* The same pattern built on the DEVICE (all pointers inside the structs are device pointers; mpc0 / mpc1 may be
 * NULL): one 64-bit key per (block row, block column) coupling of every cell, CUB radix sort, duplicate removal.
 * Two steps because the caller owns the CSR arrays: create returns the scalar nnz, export fills
 * row_ptr_out [num_rows + 1] and col_out [nnz].  Fails with MPCX_ERR_UNSUPPORTED beyond 2^31 couplings per device
 * (then mpcx_create_pattern_host applies). */
typedef struct mpcx_pattern mpcx_pattern;
int mpcx_pattern_create(const mpcx_dofmap* dofmap0, const mpcx_dofmap* dofmap1, int64_t num_cells,
                        const mpcx_mpc* mpc0, const mpcx_mpc* mpc1, void* stream, mpcx_pattern** pattern_out,
                        int64_t* nnz_out);
int mpcx_pattern_export(const mpcx_pattern* pattern, int64_t* row_ptr_out, int32_t* col_out, void* stream);
void mpcx_pattern_destroy(mpcx_pattern* pattern);

int mpcx_create_pattern_host(const int32_t* dofmap0, int32_t nd0, int32_t bs0,
                             const int32_t* dofmap1, int32_t nd1, int32_t bs1,
                             int64_t num_cells, int64_t num_block_rows,
                             const mpcx_mpc_host* mpc0, const mpcx_mpc_host* mpc1,
                             int32_t num_threads, int64_t** row_ptr_out,
                             int32_t** col_out, int64_t* nnz_out);
void mpcx_free_host(void* p);

#ifdef __cplusplus
}
#endif
#endif /* MPCX_H */
