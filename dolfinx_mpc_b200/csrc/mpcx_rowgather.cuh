// mpcx_rowgather.cuh -- "row gather" assembly for blocked (bs = gdim) P1 elasticity: owner computes, no atomics.
//
// The reference adds one dense block of bs*nd x bs*nd values per cell with MatSetValuesBlockedLocal
// (cpp/assemble_matrix.cpp:546); on the GPU that is 144 red.global.add.f64 per P1 tetrahedron (bs = 3), and the kernel
// sits exactly on the RED issue rate of the LSU (1.29 cycles per lane: 25.7 ms for the 40 M cells of BASELINE config 5,
// 8 % of the HBM roofline -- profiles/README.md).  Combining entries per cell tile in shared memory, the way the scalar
// P1 kernels do, needs 9 values per (cell, node pair): tiles would shrink to ~60 cells and combine almost nothing.
//
// Here the loop is turned inside out.  A warp owns one BLOCK ROW (one node I of the space):
//   * lanes = the bulk cells around I: each gathers its 4 vertices, evaluates the affine geometry (gradients of the
//     barycentric coordinates, volume) and leaves it in the warp's shared memory; the DIAGONAL block, to which every
//     one of those cells contributes, is reduced across the lanes with shuffles;
//   * lanes = the block columns of the row: each walks the (cell, i, j) contributions of its column -- a list built
//     once per pattern -- forms mu g_i[b] g_j[a] + lambda g_i[a] g_j[b] + delta_ab mu g_i.g_j from the staged geometry
//     and accumulates its 3 x 3 block in registers;
//   * the row is written ONCE, with plain coalesced stores: no atomics, no zero-fill of A beforehand (every entry of
//     every row is stored, zeros included), Dirichlet rows / columns zeroed on the way out.
// The price is the geometry of a cell being evaluated once per vertex (4 x ~80 flops) instead of once -- cheap next to
// 144 atomics.  Cells holding slave dofs are left out of the lists; the elimination kernel adds them afterwards.
#pragma once
#include <cub/cub.cuh>

namespace
{
struct RowPlan
{
  long long nrows_b = 0, n_inc = 0, n_con = 0, nnz_block = 0;
  int nd = 0, bs = 0, max_inc = 0;
  int* inc_off = nullptr;        // [nrows_b + 1] incidences (bulk cells around a node) of every block row
  unsigned* inc = nullptr;       // [n_inc] (position of the cell in the active list) * nd + local index of the node
  unsigned* con_off = nullptr;   // [nnz_block + 1] contributions of every block entry (block CSR order)
  unsigned short* con = nullptr; // [n_con] (incidence within the row) << 8 | local row index i << 4 | local column index j
  unsigned char* diag = nullptr; // [nrows_b] position of the diagonal block in its row
  // flat walk (k_rowgather_elast3): contributions of a row as ONE list, dealt evenly to the lanes
  unsigned char* ccol = nullptr; // [n_con] block column (position in its row) of every contribution
  unsigned* row_con = nullptr;   // [nrows_b + 1] first contribution of every block row
  unsigned char* rflag = nullptr;// [nrows_b] 1: the row has block entries without any contribution (stored as zeros)
  int flat_ok = 0;               // 0: some row has more than 256 block columns (ccol would not fit)
  double* geo = nullptr;         // [n_cells][14] scratch: staged geometry of every active cell (written by every assembly)
  long long n_cells = 0;
};
__device__ int g_rp_wide;

void row_plan_free(RowPlan* P)
{
  if (!P) return;
  cudaFree(P->inc_off); cudaFree(P->inc); cudaFree(P->con_off); cudaFree(P->con); cudaFree(P->diag);
  cudaFree(P->ccol); cudaFree(P->row_con); cudaFree(P->rflag); cudaFree(P->geo);
  delete P;
}

// block column k of block row I: col[rp[bs I] + bs k] / bs; returns the position of block column J or -1
__device__ __forceinline__ int blockcol_find(const CsrD& A, int bs, long long I, int J)
{
  const long long r0 = A.rp[bs * I];
  const int nb = (int)((A.rp[bs * I + 1] - r0) / bs);
  int lo = 0, hi = nb;
  while (lo < hi)
  {
    const int mid = (lo + hi) >> 1;
    if (__ldg(A.col + r0 + (long long)bs * mid) < bs * J) lo = mid + 1; else hi = mid;
  }
  return (lo < nb && __ldg(A.col + r0 + (long long)bs * lo) == bs * J) ? lo : -1;
}

__global__ void k_rp_keys(const int* __restrict__ dm, int nd, const int* __restrict__ cells, long long nc,
                          const int8_t* __restrict__ skip, unsigned long long* __restrict__ keys)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nc * nd) return;
  const long long idx = e / nd;
  const int i = (int)(e - idx * nd);
  if (skip && skip[idx]) { keys[e] = ~0ull; return; }
  const int cell = cells ? cells[idx] : (int)idx;
  keys[e] = ((unsigned long long)(unsigned)dm[(long long)cell * nd + i] << 32) | (unsigned long long)(idx * nd + i);
}

// first sorted key with (key >> shift) >= r, for r = 0 .. n
__global__ void k_rp_lower(const unsigned long long* __restrict__ keys, long long nkeys, int shift, long long n,
                           int* __restrict__ off32, unsigned* __restrict__ offu)
{
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > n) return;
  long long lo = 0, hi = nkeys;
  while (lo < hi)
  {
    const long long mid = (lo + hi) >> 1;
    if ((keys[mid] >> shift) < (unsigned long long)r) lo = mid + 1; else hi = mid;
  }
  if (off32) off32[r] = (int)lo;
  if (offu) offu[r] = (unsigned)lo;
}

__global__ void k_rp_low32(const unsigned long long* __restrict__ keys, long long n, unsigned* __restrict__ out)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) out[e] = (unsigned)(keys[e] & 0xffffffffull);
}

// one key per (incidence, local column j):
// (global block entry) << 24 | (block column in its row) << 16 | (incidence within the row) << 8 | i << 4 | j
__global__ void k_rp_contrib_keys(const unsigned long long* __restrict__ ikeys, long long n_inc, const int* __restrict__ inc_off,
                                  const int* __restrict__ dm, int nd, int bs, const int* __restrict__ cells, CsrD A,
                                  unsigned long long* __restrict__ ckeys, unsigned char* __restrict__ diag)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_inc) return;
  const long long I = (long long)(ikeys[e] >> 32);
  const unsigned low = (unsigned)(ikeys[e] & 0xffffffffull);
  const long long idx = low / nd;
  const int il = (int)(low - idx * nd);
  const int cell = cells ? cells[idx] : (int)idx;
  const int klocal = (int)(e - inc_off[I]);
  const long long blk0 = A.rp[bs * I] / ((long long)bs * bs);
  if (klocal > 255) g_dev_err = MPCX_ERR_UNSUPPORTED;
  for (int j = 0; j < nd; ++j)
  {
    const int J = dm[(long long)cell * nd + j];
    const int k = blockcol_find(A, bs, I, J);
    if (k < 0) { g_dev_err = MPCX_ERR_PATTERN; ckeys[e * nd + j] = ~0ull; continue; }
    if (J == I) diag[I] = (unsigned char)k;
    if (k > 255) g_rp_wide = 1;  // the flat walk stores k in 8 bits
    ckeys[e * nd + j] = ((unsigned long long)(blk0 + k) << 24) | ((unsigned long long)(k & 255) << 16)
                        | ((unsigned long long)(klocal & 255) << 8) | ((unsigned)il << 4) | (unsigned)j;
  }
}

__global__ void k_rp_con(const unsigned long long* __restrict__ ckeys, long long n, unsigned short* __restrict__ con,
                         unsigned char* __restrict__ ccol)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) { con[e] = (unsigned short)(ckeys[e] & 0xffffull); ccol[e] = (unsigned char)((ckeys[e] >> 16) & 0xffull); }
}

// per block row: its first contribution and whether one of its block entries has none
__global__ void k_rp_rowmeta(CsrD A, int bs, long long nrows_b, const unsigned* __restrict__ con_off,
                             unsigned* __restrict__ row_con, unsigned char* __restrict__ rflag)
{
  const long long I = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (I > nrows_b) return;
  const long long b0 = A.rp[bs * I] / ((long long)bs * bs);
  row_con[I] = con_off[b0];
  if (I == nrows_b) return;
  const long long b1 = A.rp[bs * (I + 1)] / ((long long)bs * bs);
  unsigned char f = 0;
  for (long long e = b0; e < b1; ++e) f |= con_off[e] == con_off[e + 1];
  rflag[I] = f;
}

// rows without any bulk cell still need their diagonal position (they are written as zeros; the value is unused)
__global__ void k_rp_diag_default(CsrD A, int bs, long long nrows_b, unsigned char* __restrict__ diag)
{
  const long long I = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (I >= nrows_b) return;
  const int k = blockcol_find(A, bs, I, (int)I);
  diag[I] = (unsigned char)(k < 0 ? 0 : k);
}

#define RP_CK(call)                                                      \
  do {                                                                   \
    cudaError_t e__ = (call);                                            \
    if (e__ != cudaSuccess) { rc = cuda_check(e__, #call); goto done; }  \
  } while (0)

int row_plan_build(const mpcx_dofmap* dm, const int32_t* cells, long long nc, const int8_t* skip, const mpcx_csr* Acsr,
                   cudaStream_t s, RowPlan** out)
{
  int rc = MPCX_OK;
  RowPlan* P = new RowPlan();
  const CsrD A{(const long long*)Acsr->row_ptr, Acsr->col, Acsr->val};
  const int nd = dm->nd, bs = dm->bs;
  const long long n0 = nc * nd;
  unsigned long long *k1 = nullptr, *k2 = nullptr;
  void* tmp = nullptr;
  size_t tb = 0;
  long long n_valid = 0;
  int last = 0;
  P->nd = nd; P->bs = bs; P->nrows_b = dm->num_dofs / bs; P->nnz_block = Acsr->nnz / ((long long)bs * bs);
  if (nd > 16 || bs < 1 || P->nrows_b >= (1ll << 31) || n0 * nd >= (1ll << 32) || P->nnz_block >= (1ll << 39))  // key: 39 + 8 + 16 bits
  { rc = fail(MPCX_ERR_UNSUPPORTED, "row plan: sizes outside the plan format"); goto done; }
  RP_CK(cudaMalloc(&P->inc_off, sizeof(int) * (size_t)(P->nrows_b + 1)));
  P->n_cells = nc;
  // scratch for the geometry-once-per-cell variant (MPCX_ROWGATHER_GEO=1 when the plan is created; measured: config 3
  // -1.4 %, config 5 +14 % because of the extra pass over the cells -- profiles/README.md r02_k -- so off by default)
  if (const char* ge = getenv("MPCX_ROWGATHER_GEO"))
    if (ge[0] == '1' && cudaMalloc(&P->geo, sizeof(double) * 14 * (size_t)(nc > 0 ? nc : 1)) != cudaSuccess)
    {
      (void)cudaGetLastError();  // no room for the scratch: the kernels evaluate the geometry per incidence instead
      P->geo = nullptr;
    }
  RP_CK(cudaMalloc(&P->diag, (size_t)P->nrows_b + 1));
  k_rp_diag_default<<<(unsigned)((P->nrows_b + 255) / 256), 256, 0, s>>>(A, bs, P->nrows_b, P->diag);
  if (n0 > 0)
  {
    RP_CK(cudaMalloc(&k1, sizeof(unsigned long long) * (size_t)n0));
    RP_CK(cudaMalloc(&k2, sizeof(unsigned long long) * (size_t)n0));
    k_rp_keys<<<(unsigned)((n0 + 255) / 256), 256, 0, s>>>(dm->map, nd, cells, nc, skip, k1);
    RP_CK(cub::DeviceRadixSort::SortKeys(nullptr, tb, k1, k2, (int)n0, 0, 64, s));
    RP_CK(cudaMalloc(&tmp, tb));
    RP_CK(cub::DeviceRadixSort::SortKeys(tmp, tb, k1, k2, (int)n0, 0, 64, s));
    cudaFree(tmp); tmp = nullptr;
  }
  // incidences of every block row (skipped cells sorted to the end: their "node" is 2^32 - 1 >= nrows_b)
  k_rp_lower<<<(unsigned)((P->nrows_b + 256) / 256), 256, 0, s>>>(k2, n0, 32, P->nrows_b, P->inc_off, nullptr);
  RP_CK(cudaMemcpyAsync(&last, P->inc_off + P->nrows_b, sizeof(int), cudaMemcpyDeviceToHost, s));
  RP_CK(cudaStreamSynchronize(s));
  n_valid = last;
  P->n_inc = n_valid;
  P->n_con = n_valid * nd;
  RP_CK(cudaMalloc(&P->inc, sizeof(unsigned) * (size_t)(n_valid > 0 ? n_valid : 1)));
  RP_CK(cudaMalloc(&P->con_off, sizeof(unsigned) * (size_t)(P->nnz_block + 1)));
  RP_CK(cudaMalloc(&P->con, sizeof(unsigned short) * (size_t)(P->n_con > 0 ? P->n_con : 1)));
  RP_CK(cudaMalloc(&P->ccol, (size_t)(P->n_con > 0 ? P->n_con : 1)));
  RP_CK(cudaMalloc(&P->row_con, sizeof(unsigned) * (size_t)(P->nrows_b + 1)));
  RP_CK(cudaMalloc(&P->rflag, (size_t)P->nrows_b + 1));
  {
    const int zero = 0;
    RP_CK(cudaMemcpyToSymbolAsync(g_rp_wide, &zero, sizeof(int), 0, cudaMemcpyHostToDevice, s));
  }
  if (n_valid > 0)
  {
    unsigned long long *c1 = nullptr, *c2 = nullptr;
    k_rp_low32<<<(unsigned)((n_valid + 255) / 256), 256, 0, s>>>(k2, n_valid, P->inc);
    RP_CK(cudaMalloc(&c1, sizeof(unsigned long long) * (size_t)P->n_con));
    cudaError_t e2 = cudaMalloc(&c2, sizeof(unsigned long long) * (size_t)P->n_con);
    if (e2 != cudaSuccess) { cudaFree(c1); rc = cuda_check(e2, "row plan alloc"); goto done; }
    k_rp_contrib_keys<<<(unsigned)((n_valid + 127) / 128), 128, 0, s>>>(k2, n_valid, P->inc_off, dm->map, nd, bs, cells, A, c1, P->diag);
    cudaFree(k1); k1 = nullptr;
    e2 = cub::DeviceRadixSort::SortKeys(nullptr, tb, c1, c2, (int)P->n_con, 0, 64, s);
    if (e2 == cudaSuccess) e2 = cudaMalloc(&tmp, tb);
    if (e2 == cudaSuccess) e2 = cub::DeviceRadixSort::SortKeys(tmp, tb, c1, c2, (int)P->n_con, 0, 64, s);
    if (e2 == cudaSuccess)
    {
      k_rp_lower<<<(unsigned)((P->nnz_block + 256) / 256), 256, 0, s>>>(c2, P->n_con, 24, P->nnz_block, nullptr, P->con_off);
      k_rp_con<<<(unsigned)((P->n_con + 255) / 256), 256, 0, s>>>(c2, P->n_con, P->con, P->ccol);
      e2 = cudaStreamSynchronize(s);
    }
    cudaFree(c1); cudaFree(c2);
    if (e2 != cudaSuccess) { rc = cuda_check(e2, "row plan contributions"); goto done; }
  }
  else
    RP_CK(cudaMemsetAsync(P->con_off, 0, sizeof(unsigned) * (size_t)(P->nnz_block + 1), s));
  k_rp_rowmeta<<<(unsigned)((P->nrows_b + 256) / 256), 256, 0, s>>>(A, bs, P->nrows_b, P->con_off, P->row_con, P->rflag);
  {
    int wide = 0;
    RP_CK(cudaMemcpyFromSymbolAsync(&wide, g_rp_wide, sizeof(int), 0, cudaMemcpyDeviceToHost, s));
    RP_CK(cudaStreamSynchronize(s));
    P->flat_ok = wide ? 0 : 1;
  }
done:
  cudaFree(k1); cudaFree(k2); cudaFree(tmp);
  if (rc != MPCX_OK) { row_plan_free(P); P = nullptr; }
  *out = P;
  return rc;
}

struct RowPlanD
{
  const int* inc_off;
  const unsigned* inc;
  const unsigned* con_off;
  const unsigned short* con;
  const unsigned char* diag;
  long long nrows_b;
  const unsigned char* ccol;
  const unsigned* row_con;
  const unsigned char* rflag;
  const double* geo;  // staged geometry per active cell (stride GSP), or nullptr: evaluate per incidence
};

// The geometry a cell lane stages, evaluated ONCE per cell instead of once per incident node (P2: 10 x): the row
// kernels then copy GS doubles per incidence -- one dependent load level and ~100 FP64 instructions less in their cell
// phase, which for P2 edge-node rows (5 cells on 16 lanes) was about as long as the contribution phase.
template <typename E>
__global__ void __launch_bounds__(256)
k_rg_geometry(IntD in, MeshD mesh, double* __restrict__ geo)
{
  constexpr int TD = E::TD, NG = TD + 1, GS = E::GS, GSP = (GS + 1) & ~1;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= in.ncells) return;
  const int cell = in.cells ? __ldg(in.cells + idx) : (int)idx;
  int xd[NG];
#pragma unroll
  for (int v = 0; v < NG; ++v) xd[v] = __ldg(mesh.xd + (long long)cell * NG + v);
  double X[NG][3];
  load_vertices<TD>(mesh, xd, X);
  P1Geom<TD> G;
  p1_geometry<TD>(X, G);
  double g[GSP];
  g[GSP - 1] = 0.0;
  E::stage(G, g);
  double2* dst = reinterpret_cast<double2*>(geo + idx * GSP);
#pragma unroll
  for (int k = 0; k < GSP / 2; ++k) dst[k] = make_double2(g[2 * k], g[2 * k + 1]);
}

template <typename E>
__device__ __forceinline__ void rg_copy_geometry(const double* __restrict__ geo, long long idx, double* g)
{
  constexpr int GS = E::GS, GSP = (GS + 1) & ~1;
  const double2* src = reinterpret_cast<const double2*>(geo + idx * GSP);
#pragma unroll
  for (int k = 0; k < GSP / 2; ++k)
  {
    const double2 v = __ldg(src + k);
    g[2 * k] = v.x;
    if (2 * k + 1 < GS) g[2 * k + 1] = v.y;
  }
}

// ---- element policies: what a cell lane stages, and how a (cell, i, j) contribution is formed from the staged data
// P1 simplex (TD = tdim = gdim = bs): the gradients of the barycentric coordinates and the volume (13 doubles in 3-D);
// block = vol (mu g_i[b] g_j[a] + lambda g_i[a] g_j[b] + delta_ab mu g_i.g_j)
template <int TD_>
struct RgP1
{
  static constexpr int TD = TD_, ND = TD_ + 1, BS = TD_, GS = ND * TD_ + 1, SMEM_TABLE = 0;
  __device__ static void init(const Tab&, double*) {}
  __device__ static void stage(const P1Geom<TD>& G, double* g)
  {
#pragma unroll
    for (int v = 0; v < ND; ++v)
#pragma unroll
      for (int k = 0; k < TD; ++k) g[v * TD + k] = G.g[v][k];
    g[ND * TD] = G.vol;
  }
  __device__ static void add(const double*, const double* g, int il, int j, double mu, double lmbda, double (*acc)[BS])
  {
    double gi[TD], gj[TD];
#pragma unroll
    for (int k = 0; k < TD; ++k) { gi[k] = g[il * TD + k]; gj[k] = g[j * TD + k]; }
    const double vol = g[ND * TD];
    double dot = 0.0;
#pragma unroll
    for (int k = 0; k < TD; ++k) dot += gi[k] * gj[k];
#pragma unroll
    for (int a = 0; a < BS; ++a)
#pragma unroll
      for (int b = 0; b < BS; ++b) acc[a][b] += vol * (mu * gi[b] * gj[a] + lmbda * gi[a] * gj[b] + (a == b ? mu * dot : 0.0));
  }
};

// Any Lagrange element on AFFINE tetrahedra, bs == 3 (P2: ND = 10; BASELINE config 3).  With a constant Jacobian the
// quadrature collapses into reference tables: for a node pair (i, j)
//   G = sum_q w_q grad phi_i (x) grad phi_j = K^T M_ij K,   M_ij[al][be] = sum_q w_q d_al phi_i(q) d_be phi_j(q),
// K = J^-1 (K[al][k] = d xi_al / d x_k), and the block is |det J| (mu G^T + lambda G + mu tr(G) I).  A cell lane stages
// K and |det J|; the M_ij of the element (ND^2 x 9 doubles, 7.2 KB for P2) sit in shared memory for the whole block: a
// contribution costs 54 + 15 FMAs instead of a quadrature loop.
template <int ND_>
struct RgAffineTet
{
  static constexpr int TD = 3, ND = ND_, BS = 3, GS = 11, SMEM_TABLE = ND_ * ND_ * 9;
  __device__ static void init(const Tab& t, double* M)
  {
    for (int e = threadIdx.x; e < ND * ND * 9; e += blockDim.x)
    {
      const int ij = e / 9, ab = e - ij * 9, i = ij / ND, j = ij - i * ND, al = ab / 3, be = ab - al * 3;
      double acc = 0.0;
      for (int q = 0; q < t.nq; ++q)
        acc += __ldg(t.w + q) * __ldg(t.dphi + (q * 3 + al) * ND + i) * __ldg(t.dphi + (q * 3 + be) * ND + j);
      M[e] = acc;
    }
  }
  __device__ static void stage(const P1Geom<3>& G, double* g)
  {
    // rows of J^-1 = gradients of the barycentric coordinates 1..3; |det J| = 6 vol
#pragma unroll
    for (int al = 0; al < 3; ++al)
#pragma unroll
      for (int k = 0; k < 3; ++k) g[al * 3 + k] = G.g[al + 1][k];
    g[9] = 6.0 * G.vol;
  }
  __device__ static void add(const double* M, const double* g, int i, int j, double mu, double lmbda, double (*acc)[BS])
  {
    const double* m = M + (i * ND + j) * 9;
    double T[3][3], Gm[3][3];
#pragma unroll
    for (int al = 0; al < 3; ++al)
#pragma unroll
      for (int l = 0; l < 3; ++l) T[al][l] = m[al * 3] * g[l] + m[al * 3 + 1] * g[3 + l] + m[al * 3 + 2] * g[6 + l];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int l = 0; l < 3; ++l) Gm[k][l] = g[k] * T[0][l] + g[3 + k] * T[1][l] + g[6 + k] * T[2][l];
    const double dj = g[9], tr = mu * (Gm[0][0] + Gm[1][1] + Gm[2][2]);
#pragma unroll
    for (int a = 0; a < BS; ++a)
#pragma unroll
      for (int b = 0; b < BS; ++b) acc[a][b] += dj * (mu * Gm[b][a] + lmbda * Gm[a][b] + (a == b ? tr : 0.0));
  }
};

// One warp per block row.  Global-load latency is what this kernel fights (three dependent levels: row header ->
// incidence / contribution lists -> dofmap -> coordinates): the header of the warp's NEXT row is requested while the
// current one is processed, the incidence word, the column's contribution words (8 at a time, in registers) and its
// Dirichlet flags are requested at the top of the row, before the cell phase.
template <typename E, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_rowgather_elast(RowPlanD P, Tab t, IntD in, MeshD mesh, const int8_t* __restrict__ bc, CsrD A)
{
  constexpr int TD = E::TD, ND = E::ND, BS = E::BS, GS = E::GS, NG = TD + 1, CW = 8;
  extern __shared__ double rg_smem[];
  double* M = rg_smem;                              // element tables (affine policy)
  double* geo = rg_smem + E::SMEM_TABLE + (threadIdx.x >> 5) * 32 * GS;  // [32 cells][GS] of this warp
  E::init(t, M);
  if (E::SMEM_TABLE) __syncthreads();
  const int lane = threadIdx.x & 31;
  const double mu = in.c[0], lmbda = in.c[1];
  const long long wstride = (long long)gridDim.x * 8;
  long long I = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (I >= P.nrows_b) return;
  int i0 = __ldg(P.inc_off + I), i1 = __ldg(P.inc_off + I + 1), kd = __ldg(P.diag + I);
  long long r0 = __ldg(A.rp + BS * I), r1 = __ldg(A.rp + BS * I + 1);
  for (;;)
  {
    const long long In = I + wstride;
    const bool more = In < P.nrows_b;
    int n_i0 = 0, n_i1 = 0, n_kd = 0;
    long long n_r0 = 0, n_r1 = 0;
    if (more)
    {
      n_i0 = __ldg(P.inc_off + In); n_i1 = __ldg(P.inc_off + In + 1); n_kd = __ldg(P.diag + In);
      n_r0 = __ldg(A.rp + BS * In); n_r1 = __ldg(A.rp + BS * In + 1);
    }
    const int ninc = i1 - i0, nb = (int)((r1 - r0) / BS);
    const long long blk0 = r0 / (BS * BS);
    bool bcr[BS];
#pragma unroll
    for (int a = 0; a < BS; ++a) bcr[a] = bc ? bc[BS * I + a] != 0 : false;
    for (int cc = 0; cc < nb; cc += 32)  // block columns of the row, 32 at a time (one pass on P1 meshes)
    {
      const int kc = cc + lane;
      const bool col_ok = kc < nb;
      double acc[BS][BS];
#pragma unroll
      for (int a = 0; a < BS; ++a)
#pragma unroll
        for (int b = 0; b < BS; ++b) acc[a][b] = 0.0;
      // requests of this pass: contribution range, column node (for its Dirichlet flags), first incidence word
      unsigned c_lo = 0, c_hi = 0;
      int J = 0;
      if (col_ok)
      {
        c_lo = __ldg(P.con_off + blk0 + kc); c_hi = __ldg(P.con_off + blk0 + kc + 1);
        J = __ldg(A.col + r0 + (long long)BS * kc) / BS;
      }
      unsigned w0 = lane < ninc ? __ldg(P.inc + i0 + lane) : 0u;
      unsigned short cw[CW];
#pragma unroll
      for (int u = 0; u < CW; ++u) cw[u] = (col_ok && c_lo + u < c_hi) ? __ldg(P.con + c_lo + u) : (unsigned short)0xffffu;
      bool bcc[BS];
#pragma unroll
      for (int b = 0; b < BS; ++b) bcc[b] = (bc && col_ok) ? bc[BS * J + b] != 0 : false;
      const bool diag_here = kd >= cc && kd < cc + 32;
      for (int ch = 0; ch < ninc || ch == 0; ch += 32)  // the cells around the node, 32 at a time
      {
        // ---- lanes = cells: affine geometry -> shared memory; diagonal block reduced with shuffles
        double dg[BS][BS];
#pragma unroll
        for (int a = 0; a < BS; ++a)
#pragma unroll
          for (int b = 0; b < BS; ++b) dg[a][b] = 0.0;
        __syncwarp();
        if (ch + lane < ninc)
        {
          const unsigned w = ch == 0 ? w0 : __ldg(P.inc + i0 + ch + lane);
          const long long idx = w / ND;
          const int il = (int)(w - idx * ND);
          const int cell = in.cells ? __ldg(in.cells + idx) : (int)idx;
          int xd[NG];
#pragma unroll
          for (int v = 0; v < NG; ++v) xd[v] = __ldg(mesh.xd + (long long)cell * NG + v);
          double X[NG][3];
          load_vertices<TD>(mesh, xd, X);
          P1Geom<TD> G;
          p1_geometry<TD>(X, G);
          double* g = geo + lane * GS;
          E::stage(G, g);
          if (diag_here) E::add(M, g, il, il, mu, lmbda, dg);  // reads back this lane's own stores
        }
        if (diag_here)
        {
#pragma unroll
          for (int a = 0; a < BS; ++a)
#pragma unroll
            for (int b = 0; b < BS; ++b)
            {
              double v = dg[a][b];
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
              if (kc == kd) acc[a][b] += v;
            }
        }
        __syncwarp();
        // ---- lanes = block columns: the contributions of the column that come from this chunk of cells
        if (col_ok && kc != kd)
        {
#pragma unroll
          for (int u = 0; u < CW; ++u)
          {
            const unsigned c = cw[u];
            const int kl = (int)(c >> 8) - ch;
            if (c != 0xffffu && kl >= 0 && kl < 32) E::add(M, geo + kl * GS, (int)((c >> 4) & 15u), (int)(c & 15u), mu, lmbda, acc);
          }
          for (unsigned q = c_lo + CW; q < c_hi; ++q)  // columns fed by more than CW cells (rare)
          {
            const unsigned c = __ldg(P.con + q);
            const int kl = (int)(c >> 8) - ch;
            if (kl >= 0 && kl < 32) E::add(M, geo + kl * GS, (int)((c >> 4) & 15u), (int)(c & 15u), mu, lmbda, acc);
          }
        }
      }
      // ---- the row is written once (Dirichlet rows / columns zeroed, cpp/assemble_matrix.cpp:513-533); scalar row
      // BS I + a starts at r0 + a BS nb
      if (col_ok)
      {
#pragma unroll
        for (int a = 0; a < BS; ++a)
        {
          double* dst = A.val + r0 + (long long)a * BS * nb + (long long)BS * kc;
#pragma unroll
          for (int b = 0; b < BS; ++b) dst[b] = (bcr[a] || bcc[b]) ? 0.0 : acc[a][b];
        }
      }
    }
    if (!more) break;
    I = In; i0 = n_i0; i1 = n_i1; kd = n_kd; r0 = n_r0; r1 = n_r1;
  }
}
// Two block rows per warp (16 lanes each): P1 rows have ~15 block columns and ~24 cells, so a whole warp per row left
// half of the lanes idle in the contribution loop and one row's chain of dependent global loads in flight per warp.
// All cells of a row (up to 32 per pass) are staged first, 16 at a time, then the columns are walked 16 at a time
// without re-staging.  Rows with more than 32 cells take further passes that add onto the stored row (same lanes own
// the same entries: plain read-modify-write, no atomics).
template <typename E, int MINB, bool PRE>
__global__ void __launch_bounds__(256, MINB)
k_rowgather_elast2(RowPlanD P, Tab t, IntD in, MeshD mesh, const int8_t* __restrict__ bc, CsrD A)
{
  constexpr int TD = E::TD, ND = E::ND, BS = E::BS, GS = E::GS, NG = TD + 1, CW = 8;
  extern __shared__ double rg_smem[];
  double* M = rg_smem;
  const int lane = threadIdx.x & 31, sub = lane >> 4, sl = lane & 15;
  double* geo = rg_smem + E::SMEM_TABLE + ((threadIdx.x >> 5) * 2 + sub) * 32 * GS;  // [32 cells][GS] of this half-warp
  E::init(t, M);
  if (E::SMEM_TABLE) __syncthreads();
  const double mu = in.c[0], lmbda = in.c[1];
  const long long wstride = (long long)gridDim.x * 16;  // rows per sweep of the grid (8 warps x 2 rows per block)
  long long I = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * 2 + sub;
  if (I - sub >= P.nrows_b) return;  // both rows of the warp are past the end
  auto hdr = [&](long long row, int& i0, int& i1, int& kd, long long& r0, long long& r1) {
    i0 = i1 = kd = 0; r0 = r1 = 0;
    if (row < P.nrows_b)
    {
      i0 = __ldg(P.inc_off + row); i1 = __ldg(P.inc_off + row + 1); kd = __ldg(P.diag + row);
      r0 = __ldg(A.rp + BS * row); r1 = __ldg(A.rp + BS * row + 1);
    }
  };
  int i0, i1, kd;
  long long r0, r1;
  hdr(I, i0, i1, kd, r0, r1);
  for (;;)
  {
    const long long In = I + wstride;
    const bool more = In - sub < P.nrows_b;  // warp-uniform
    int n_i0, n_i1, n_kd;
    long long n_r0, n_r1;
    hdr(In, n_i0, n_i1, n_kd, n_r0, n_r1);
    const int ninc = i1 - i0, nb = (int)((r1 - r0) / BS);
    const long long blk0 = r0 / (BS * BS);
    const int ninc_w = max(__shfl_sync(0xffffffffu, ninc, 0), __shfl_sync(0xffffffffu, ninc, 16));
    const int nb_w = max(__shfl_sync(0xffffffffu, nb, 0), __shfl_sync(0xffffffffu, nb, 16));
    bool bcr[BS];
#pragma unroll
    for (int a = 0; a < BS; ++a) bcr[a] = (bc && nb > 0) ? bc[BS * I + a] != 0 : false;
    // requests for the first 16 columns, issued before the cell phase
    unsigned c_lo = 0, c_hi = 0;
    int J = 0;
    if (sl < nb)
    {
      c_lo = __ldg(P.con_off + blk0 + sl); c_hi = __ldg(P.con_off + blk0 + sl + 1);
      J = __ldg(A.col + r0 + (long long)BS * sl) / BS;
    }
    for (int sc = 0; sc < ninc_w || sc == 0; sc += 32)
    {
      // ---- lanes = cells (two steps of 16): affine geometry -> shared memory; diagonal block reduced over the half-warp
      double dg[BS][BS];
#pragma unroll
      for (int a = 0; a < BS; ++a)
#pragma unroll
        for (int b = 0; b < BS; ++b) dg[a][b] = 0.0;
      __syncwarp();
#pragma unroll
      for (int ss = 0; ss < 32; ss += 16)
      {
        const int k = sc + ss + sl;
        if (k < ninc)
        {
          const unsigned w = __ldg(P.inc + i0 + k);
          const long long idx = w / ND;
          const int il = (int)(w - idx * ND);
          double* g = geo + (ss + sl) * GS;
          if (PRE) rg_copy_geometry<E>(P.geo, idx, g);
          else
          {
            const int cell = in.cells ? __ldg(in.cells + idx) : (int)idx;
            int xd[NG];
#pragma unroll
            for (int v = 0; v < NG; ++v) xd[v] = __ldg(mesh.xd + (long long)cell * NG + v);
            double X[NG][3];
            load_vertices<TD>(mesh, xd, X);
            P1Geom<TD> G;
            p1_geometry<TD>(X, G);
            E::stage(G, g);
          }
          E::add(M, g, il, il, mu, lmbda, dg);  // reads back this lane's own stores
        }
      }
#pragma unroll
      for (int a = 0; a < BS; ++a)
#pragma unroll
        for (int b = 0; b < BS; ++b)
        {
          double v = dg[a][b];
#pragma unroll
          for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          dg[a][b] = v;
        }
      __syncwarp();
      // ---- lanes = block columns, 16 at a time
      for (int cc = 0; cc < nb_w; cc += 16)
      {
        const int kc = cc + sl;
        const bool col_ok = kc < nb;
        if (cc > 0 || sc > 0)
        {
          c_lo = c_hi = 0; J = 0;
          if (col_ok)
          {
            c_lo = __ldg(P.con_off + blk0 + kc); c_hi = __ldg(P.con_off + blk0 + kc + 1);
            J = __ldg(A.col + r0 + (long long)BS * kc) / BS;
          }
        }
        unsigned short cw[CW];
#pragma unroll
        for (int u = 0; u < CW; ++u) cw[u] = (col_ok && c_lo + u < c_hi) ? __ldg(P.con + c_lo + u) : (unsigned short)0xffffu;
        bool bcc[BS];
#pragma unroll
        for (int b = 0; b < BS; ++b) bcc[b] = (bc && col_ok) ? bc[BS * J + b] != 0 : false;
        double acc[BS][BS];
        const bool is_diag = col_ok && kc == kd;
#pragma unroll
        for (int a = 0; a < BS; ++a)
#pragma unroll
          for (int b = 0; b < BS; ++b) acc[a][b] = is_diag ? dg[a][b] : 0.0;
        if (col_ok && !is_diag)
        {
#pragma unroll
          for (int u = 0; u < CW; ++u)
          {
            const unsigned c = cw[u];
            const int kl = (int)(c >> 8) - sc;
            if (c != 0xffffu && kl >= 0 && kl < 32) E::add(M, geo + kl * GS, (int)((c >> 4) & 15u), (int)(c & 15u), mu, lmbda, acc);
          }
          for (unsigned q = c_lo + CW; q < c_hi; ++q)  // columns fed by more than CW cells (rare)
          {
            const unsigned c = __ldg(P.con + q);
            const int kl = (int)(c >> 8) - sc;
            if (kl >= 0 && kl < 32) E::add(M, geo + kl * GS, (int)((c >> 4) & 15u), (int)(c & 15u), mu, lmbda, acc);
          }
        }
        // ---- the row is written once (Dirichlet rows / columns zeroed, cpp/assemble_matrix.cpp:513-533); scalar row
        // BS I + a starts at r0 + a BS nb.  Passes after the first (rows with more than 32 cells) add onto it.
        if (col_ok)
        {
#pragma unroll
          for (int a = 0; a < BS; ++a)
          {
            double* dst = A.val + r0 + (long long)a * BS * nb + (long long)BS * kc;
#pragma unroll
            for (int b = 0; b < BS; ++b)
            {
              const double v = (bcr[a] || bcc[b]) ? 0.0 : acc[a][b];
              if (sc == 0) dst[b] = v; else dst[b] += v;
            }
          }
        }
      }
    }
    if (!more) break;
    I = In; i0 = n_i0; i1 = n_i1; kd = n_kd; r0 = n_r0; r1 = n_r1;
  }
}

// Flat walk: the contributions of a row are ONE list (sorted by block column); the 16 lanes of the row's half-warp take
// equal contiguous shares of it instead of a column each.  With a lane per column the lanes of a pass ran as long as
// the column with the most contributions (P2: 1 .. 24 per column) and 13 of 32 lanes were active in the FP64
// instructions of the kernel (ncu, profiles/r02_i); here every lane forms the same number of 3 x 3 contributions
// (+- 1).  A lane accumulates in registers while the column stays the same and stores a column when it leaves it; the
// column a share STARTS in the middle of (head) goes to the lane that holds the column's beginning through a segmented
// suffix sum over the half-warp (shuffles, 9 values, once per row), so every entry is still written exactly once,
// without atomics.  The diagonal block is just another column.  Block entries without any contribution (pairs whose
// cells all hold slaves) are stored as zeros by the rows flagged at plan time.
template <typename E, int MINB, bool PRE>
__global__ void __launch_bounds__(256, MINB)
k_rowgather_elast3(RowPlanD P, Tab t, IntD in, MeshD mesh, const int8_t* __restrict__ bc, CsrD A)
{
  constexpr int TD = E::TD, ND = E::ND, BS = E::BS, GS = E::GS, NG = TD + 1;
  extern __shared__ double rg_smem[];
  double* M = rg_smem;
  const int lane = threadIdx.x & 31, sub = lane >> 4, sl = lane & 15, hw = (threadIdx.x >> 5) * 2 + sub;
  double* geo = rg_smem + E::SMEM_TABLE + hw * 32 * GS;  // [32 cells][GS] of this half-warp
  unsigned char* cmask = reinterpret_cast<unsigned char*>(rg_smem + E::SMEM_TABLE + 16 * 32 * GS) + hw * 256;  // Dirichlet bits per column
  E::init(t, M);
  if (E::SMEM_TABLE) __syncthreads();
  const double mu = in.c[0], lmbda = in.c[1];
  const long long wstride = (long long)gridDim.x * 16;
  long long I = ((long long)blockIdx.x * 8 + (threadIdx.x >> 5)) * 2 + sub;
  if (I - sub >= P.nrows_b) return;
  auto hdr = [&](long long row, int& i0, int& i1, unsigned& t0, unsigned& t1, long long& r0, long long& r1, int& fl) {
    i0 = i1 = fl = 0; t0 = t1 = 0u; r0 = r1 = 0;
    if (row < P.nrows_b)
    {
      i0 = __ldg(P.inc_off + row); i1 = __ldg(P.inc_off + row + 1); fl = __ldg(P.rflag + row);
      t0 = __ldg(P.row_con + row); t1 = __ldg(P.row_con + row + 1);
      r0 = __ldg(A.rp + BS * row); r1 = __ldg(A.rp + BS * row + 1);
    }
  };
  int i0, i1, fl;
  unsigned tb0, tb1;
  long long r0, r1;
  hdr(I, i0, i1, tb0, tb1, r0, r1, fl);
  for (;;)
  {
    const long long In = I + wstride;
    const bool more = In - sub < P.nrows_b;  // warp-uniform
    int n_i0, n_i1, n_fl;
    unsigned n_t0, n_t1;
    long long n_r0, n_r1;
    hdr(In, n_i0, n_i1, n_t0, n_t1, n_r0, n_r1, n_fl);
    const int ninc = i1 - i0, nb = (int)((r1 - r0) / BS);
    const int T = (int)(tb1 - tb0);
    const int ninc_w = max(__shfl_sync(0xffffffffu, ninc, 0), __shfl_sync(0xffffffffu, ninc, 16));
    const int T_w = max(__shfl_sync(0xffffffffu, T, 0), __shfl_sync(0xffffffffu, T, 16));
    const int nb_w = max(__shfl_sync(0xffffffffu, nb, 0), __shfl_sync(0xffffffffu, nb, 16));
    bool bcr[BS];
#pragma unroll
    for (int a = 0; a < BS; ++a) bcr[a] = (bc && nb > 0) ? bc[BS * I + a] != 0 : false;
    // this lane's share of the list and its first words, requested before the cell phase
    const unsigned ta = tb0 + (unsigned)(((long long)T * sl) >> 4), te = tb0 + (unsigned)(((long long)T * (sl + 1)) >> 4);
    const int prevcol = ta > tb0 ? (int)__ldg(P.ccol + ta - 1) : -1;  // column of the contribution before the share
    constexpr int CW = 8;
    unsigned pw[CW];  // (column << 16 | contribution word) of the first CW contributions of the share
#pragma unroll
    for (int u = 0; u < CW; ++u)
      pw[u] = ta + u < te ? ((unsigned)__ldg(P.ccol + ta + u) << 16) | (unsigned)__ldg(P.con + ta + u) : 0xffffffffu;
    // Dirichlet bits of the columns -> shared memory (lanes = columns; two dependent loads, hidden behind the cell phase)
    __syncwarp();
    if (bc)
      for (int kc = sl; kc < nb; kc += 16)
      {
        const int J = __ldg(A.col + r0 + (long long)BS * kc) / BS;
        unsigned m = 0u;
#pragma unroll
        for (int b = 0; b < BS; ++b) m |= (bc[BS * J + b] != 0 ? 1u : 0u) << b;
        cmask[kc] = (unsigned char)m;
      }
    auto store = [&](int col, const double (*acc)[BS], bool first_pass) {
      const unsigned m = bc ? cmask[col] : 0u;
#pragma unroll
      for (int a = 0; a < BS; ++a)
      {
        double* dst = A.val + r0 + (long long)a * BS * nb + (long long)BS * col;
#pragma unroll
        for (int b = 0; b < BS; ++b)
        {
          const double v = (bcr[a] || ((m >> b) & 1u)) ? 0.0 : acc[a][b];
          if (first_pass) dst[b] = v; else dst[b] += v;
        }
      }
    };
    for (int sc = 0; sc < ninc_w || sc == 0; sc += 32)
    {
      // ---- lanes = cells (two steps of 16): affine geometry -> shared memory
      __syncwarp();
#pragma unroll
      for (int ss = 0; ss < 32; ss += 16)
      {
        const int k = sc + ss + sl;
        if (k < ninc)
        {
          const unsigned w = __ldg(P.inc + i0 + k);
          const long long idx = w / ND;
          if (PRE) rg_copy_geometry<E>(P.geo, idx, geo + (ss + sl) * GS);
          else
          {
            const int cell = in.cells ? __ldg(in.cells + idx) : (int)idx;
            int xd[NG];
#pragma unroll
            for (int v = 0; v < NG; ++v) xd[v] = __ldg(mesh.xd + (long long)cell * NG + v);
            double X[NG][3];
            load_vertices<TD>(mesh, xd, X);
            P1Geom<TD> G;
            p1_geometry<TD>(X, G);
            E::stage(G, geo + (ss + sl) * GS);
          }
        }
      }
      __syncwarp();
      // ---- lanes = equal shares of the row's contributions
      double acc[BS][BS], H[BS][BS];
#pragma unroll
      for (int a = 0; a < BS; ++a)
#pragma unroll
        for (int b = 0; b < BS; ++b) { acc[a][b] = 0.0; H[a][b] = 0.0; }
      const bool nonempty = ta < te;
      const int firstcol = nonempty ? (int)(pw[0] >> 16) : -2;
      const bool headful = nonempty && prevcol == firstcol;  // the share starts inside the column of the lane before it
      int cur = firstcol;
      bool in_head = headful;
      const int steps = (T_w + 15) >> 4;  // >= the longest share of the warp
      auto step = [&](unsigned word) {
        const unsigned c16 = word & 0xffffu;
        const int col = (int)(word >> 16);
        if (col != cur)
        {
          if (in_head)
          {
#pragma unroll
            for (int a = 0; a < BS; ++a)
#pragma unroll
              for (int b = 0; b < BS; ++b) H[a][b] = acc[a][b];
            in_head = false;
          }
          else store(cur, acc, sc == 0);
          cur = col;
#pragma unroll
          for (int a = 0; a < BS; ++a)
#pragma unroll
            for (int b = 0; b < BS; ++b) acc[a][b] = 0.0;
        }
        const int kl = (int)(c16 >> 8) - sc;
        if (kl >= 0 && kl < 32) E::add(M, geo + kl * GS, (int)((c16 >> 4) & 15u), (int)(c16 & 15u), mu, lmbda, acc);
      };
#pragma unroll
      for (int u = 0; u < CW; ++u)
        if (u < steps && pw[u] != 0xffffffffu) step(pw[u]);
      for (int m = CW; m < steps; ++m)  // shares longer than CW (rows with more than 16 CW contributions)
      {
        const unsigned tt = ta + (unsigned)m;
        if (tt < te) step(((unsigned)__ldg(P.ccol + tt) << 16) | (unsigned)__ldg(P.con + tt));
      }
      // What is left in acc: the LAST column of the share, whose beginning this lane holds (tail; the lanes after it may
      // hold more of it as their heads) -- or, when the whole share lies inside the previous lane's column, a head.  An
      // empty share carries a zero head of the column before it, which keeps that column's lanes consecutive.
      const bool has_tail = nonempty && !in_head;
      if (in_head)
      {
#pragma unroll
        for (int a = 0; a < BS; ++a)
#pragma unroll
          for (int b = 0; b < BS; ++b) H[a][b] = acc[a][b];
      }
      const int hkey = headful ? firstcol : ((!nonempty && prevcol >= 0) ? prevcol : -1 - sl);
      // segmented suffix sum of the heads over the half-warp (equal keys sit in consecutive lanes)
#pragma unroll
      for (int o = 1; o < 16; o <<= 1)
      {
        const int k2 = __shfl_down_sync(0xffffffffu, hkey, o, 16);
        const bool take = sl + o < 16 && k2 == hkey;
#pragma unroll
        for (int a = 0; a < BS; ++a)
#pragma unroll
          for (int b = 0; b < BS; ++b)
          {
            const double v2 = __shfl_down_sync(0xffffffffu, H[a][b], o, 16);
            if (take) H[a][b] += v2;
          }
      }
      {
        const int k2 = __shfl_down_sync(0xffffffffu, hkey, 1, 16);
        const bool take = has_tail && sl + 1 < 16 && k2 == cur;
#pragma unroll
        for (int a = 0; a < BS; ++a)
#pragma unroll
          for (int b = 0; b < BS; ++b)
          {
            const double v2 = __shfl_down_sync(0xffffffffu, H[a][b], 1, 16);
            if (take) acc[a][b] += v2;
          }
        if (has_tail) store(cur, acc, sc == 0);
      }
    }
    // block entries without contributions (and rows without any bulk cell): zeros
    if (fl || T == 0)
    {
      const long long blk0 = r0 / (BS * BS);
      for (int kc = sl; kc < nb; kc += 16)
        if (__ldg(P.con_off + blk0 + kc) == __ldg(P.con_off + blk0 + kc + 1))
        {
#pragma unroll
          for (int a = 0; a < BS; ++a)
#pragma unroll
            for (int b = 0; b < BS; ++b) A.val[r0 + (long long)a * BS * nb + (long long)BS * kc + b] = 0.0;
        }
    }
    (void)nb_w;
    if (!more) break;
    I = In; i0 = n_i0; i1 = n_i1; tb0 = n_t0; tb1 = n_t1; r0 = n_r0; r1 = n_r1; fl = n_fl;
  }
}
}  // namespace
