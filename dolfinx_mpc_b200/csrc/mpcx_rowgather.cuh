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
};

void row_plan_free(RowPlan* P)
{
  if (!P) return;
  cudaFree(P->inc_off); cudaFree(P->inc); cudaFree(P->con_off); cudaFree(P->con); cudaFree(P->diag);
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

// one key per (incidence, local column j): (global block entry) << 16 | (incidence within the row) << 8 | i << 4 | j
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
    ckeys[e * nd + j] = ((unsigned long long)(blk0 + k) << 16) | ((unsigned long long)(klocal & 255) << 8) | ((unsigned)il << 4) | (unsigned)j;
  }
}

__global__ void k_rp_con(const unsigned long long* __restrict__ ckeys, long long n, unsigned short* __restrict__ con)
{
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) con[e] = (unsigned short)(ckeys[e] & 0xffffull);
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
  if (nd > 16 || bs < 1 || P->nrows_b >= (1ll << 31) || n0 * nd >= (1ll << 32) || P->nnz_block >= (1ll << 39))
  { rc = fail(MPCX_ERR_UNSUPPORTED, "row plan: sizes outside the plan format"); goto done; }
  RP_CK(cudaMalloc(&P->inc_off, sizeof(int) * (size_t)(P->nrows_b + 1)));
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
  if (n_valid > 0)
  {
    unsigned long long *c1 = nullptr, *c2 = nullptr;
    k_rp_low32<<<(unsigned)((n_valid + 255) / 256), 256, 0, s>>>(k2, n_valid, P->inc);
    RP_CK(cudaMalloc(&c1, sizeof(unsigned long long) * (size_t)P->n_con));
    cudaError_t e2 = cudaMalloc(&c2, sizeof(unsigned long long) * (size_t)P->n_con);
    if (e2 != cudaSuccess) { cudaFree(c1); rc = cuda_check(e2, "row plan alloc"); goto done; }
    k_rp_contrib_keys<<<(unsigned)((n_valid + 127) / 128), 128, 0, s>>>(k2, n_valid, P->inc_off, dm->map, nd, bs, cells, A, c1, P->diag);
    cudaFree(k1); k1 = nullptr;
    e2 = cub::DeviceRadixSort::SortKeys(nullptr, tb, c1, c2, (int)P->n_con, 0, 56, s);
    if (e2 == cudaSuccess) e2 = cudaMalloc(&tmp, tb);
    if (e2 == cudaSuccess) e2 = cub::DeviceRadixSort::SortKeys(tmp, tb, c1, c2, (int)P->n_con, 0, 56, s);
    if (e2 == cudaSuccess)
    {
      k_rp_lower<<<(unsigned)((P->nnz_block + 256) / 256), 256, 0, s>>>(c2, P->n_con, 16, P->nnz_block, nullptr, P->con_off);
      k_rp_con<<<(unsigned)((P->n_con + 255) / 256), 256, 0, s>>>(c2, P->n_con, P->con);
      e2 = cudaStreamSynchronize(s);
    }
    cudaFree(c1); cudaFree(c2);
    if (e2 != cudaSuccess) { rc = cuda_check(e2, "row plan contributions"); goto done; }
  }
  else
    RP_CK(cudaMemsetAsync(P->con_off, 0, sizeof(unsigned) * (size_t)(P->nnz_block + 1), s));
  RP_CK(cudaStreamSynchronize(s));
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
};

// One warp per block row.  TD = tdim = gdim = bs, P1 (nd = TD + 1).
template <int TD>
__global__ void __launch_bounds__(256)
k_rowgather_elast_p1(RowPlanD P, IntD in, MeshD mesh, const int* __restrict__ dm, const int8_t* __restrict__ bc, CsrD A)
{
  constexpr int NV = TD + 1, BS = TD, GS = NV * TD + 1;  // doubles per staged cell: gradients + volume
  __shared__ double geo_all[8][32 * GS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* geo = geo_all[warp];
  const double mu = in.c[0], lmbda = in.c[1];
  const long long wstride = (long long)gridDim.x * 8;
  for (long long I = (long long)blockIdx.x * 8 + warp; I < P.nrows_b; I += wstride)
  {
    const int i0 = __ldg(P.inc_off + I), ninc = __ldg(P.inc_off + I + 1) - i0;
    const long long r0 = __ldg(A.rp + BS * I);
    const int nb = (int)((__ldg(A.rp + BS * I + 1) - r0) / BS);
    const long long blk0 = r0 / (BS * BS);
    const int kd = __ldg(P.diag + I);
    bool bcr[BS];
#pragma unroll
    for (int a = 0; a < BS; ++a) bcr[a] = bc ? bc[BS * I + a] != 0 : false;
    for (int cc = 0; cc < nb; cc += 32)  // block columns of the row, 32 at a time (one pass for P1 meshes)
    {
      const int kc = cc + lane;
      double acc[BS][BS];
#pragma unroll
      for (int a = 0; a < BS; ++a)
#pragma unroll
        for (int b = 0; b < BS; ++b) acc[a][b] = 0.0;
      unsigned c_lo = 0, c_hi = 0;
      if (kc < nb) { c_lo = __ldg(P.con_off + blk0 + kc); c_hi = __ldg(P.con_off + blk0 + kc + 1); }
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
          const unsigned w = __ldg(P.inc + i0 + ch + lane);
          const long long idx = w / NV;
          const int il = (int)(w - idx * NV);
          const int cell = in.cells ? __ldg(in.cells + idx) : (int)idx;
          int xd[NV];
#pragma unroll
          for (int v = 0; v < NV; ++v) xd[v] = __ldg(mesh.xd + (long long)cell * NV + v);
          double X[NV][3];
          load_vertices<TD>(mesh, xd, X);
          P1Geom<TD> G;
          p1_geometry<TD>(X, G);
          double* g = geo + lane * GS;
#pragma unroll
          for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int k = 0; k < TD; ++k) g[v * TD + k] = G.g[v][k];
          g[NV * TD] = G.vol;
          double gi[TD];
#pragma unroll
          for (int v = 0; v < NV; ++v)
            if (v == il)
            {
#pragma unroll
              for (int k = 0; k < TD; ++k) gi[k] = G.g[v][k];
            }
          double dot = 0.0;
#pragma unroll
          for (int k = 0; k < TD; ++k) dot += gi[k] * gi[k];
#pragma unroll
          for (int a = 0; a < BS; ++a)
#pragma unroll
            for (int b = 0; b < BS; ++b) dg[a][b] = G.vol * ((mu + lmbda) * gi[a] * gi[b] + (a == b ? mu * dot : 0.0));
        }
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
        __syncwarp();
        // ---- lanes = block columns: the contributions of the column that come from this chunk of cells
        if (kc < nb && kc != kd)
          for (unsigned q = c_lo; q < c_hi; ++q)
          {
            const unsigned cw = __ldg(P.con + q);
            const int kl = (int)(cw >> 8) - ch, il = (int)((cw >> 4) & 15u), j = (int)(cw & 15u);
            if (kl < 0 || kl >= 32) continue;  // a cell of another chunk (rows with more than 32 cells only)
            const double* g = geo + kl * GS;
            double gi[TD], gj[TD];
#pragma unroll
            for (int k = 0; k < TD; ++k) { gi[k] = g[il * TD + k]; gj[k] = g[j * TD + k]; }
            const double vol = g[NV * TD];
            double dot = 0.0;
#pragma unroll
            for (int k = 0; k < TD; ++k) dot += gi[k] * gj[k];
#pragma unroll
            for (int a = 0; a < BS; ++a)
#pragma unroll
              for (int b = 0; b < BS; ++b)
                acc[a][b] += vol * (mu * gi[b] * gj[a] + lmbda * gi[a] * gj[b] + (a == b ? mu * dot : 0.0));
          }
      }
      // ---- the row is written once (Dirichlet rows / columns zeroed, cpp/assemble_matrix.cpp:513-533)
      if (kc < nb)
      {
        const int J = __ldg(A.col + r0 + (long long)BS * kc) / BS;
#pragma unroll
        for (int a = 0; a < BS; ++a)
        {
          double* dst = A.val + __ldg(A.rp + BS * I + a) + (long long)BS * kc;
#pragma unroll
          for (int b = 0; b < BS; ++b)
          {
            const bool z = bcr[a] || (bc && bc[BS * J + b] != 0);
            dst[b] = z ? 0.0 : acc[a][b];
          }
        }
      }
    }
  }
}
// The same row-gather scheme for ANY Lagrange element on affine tetrahedra with bs == 3 (P2: ND = 10; BASELINE config 3).
// With a constant Jacobian the quadrature collapses into reference tables: for a node pair (i, j)
//   G = sum_q w_q grad phi_i (x) grad phi_j = K^T M_ij K,   M_ij[al][be] = sum_q w_q d_al phi_i(q) d_be phi_j(q),
// K = J^-1 (K[al][k] = d xi_al / d x_k), and the 3 x 3 block is |det J| (mu G^T + lambda G + mu tr(G) I).  The cell lanes
// stage K and |det J| (10 doubles per cell), the M_ij of the element (ND^2 x 9 doubles, 7.2 KB for P2) sit in shared
// memory for the whole block; a contribution costs 54 + 15 FMAs instead of a quadrature loop.
template <int ND>
__global__ void __launch_bounds__(256)
k_rowgather_elast_affine3d(RowPlanD P, Tab t, IntD in, MeshD mesh, const int* __restrict__ dm, const int8_t* __restrict__ bc, CsrD A)
{
  constexpr int BS = 3, GS = 11;  // doubles per staged cell: K (9) + |det J| (+ 1 of padding: odd stride)
  extern __shared__ double rg_smem[];
  double* M = rg_smem;                       // [ND][ND][3][3]
  double* geo_all = rg_smem + ND * ND * 9;   // [8 warps][32 cells][GS]
  for (int e = threadIdx.x; e < ND * ND * 9; e += blockDim.x)
  {
    const int ij = e / 9, ab = e - ij * 9, i = ij / ND, j = ij - i * ND, al = ab / 3, be = ab - al * 3;
    double acc = 0.0;
    for (int q = 0; q < t.nq; ++q)
      acc += __ldg(t.w + q) * __ldg(t.dphi + (q * 3 + al) * ND + i) * __ldg(t.dphi + (q * 3 + be) * ND + j);
    M[e] = acc;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* geo = geo_all + warp * 32 * GS;
  const double mu = in.c[0], lmbda = in.c[1];
  // block (i, j) of one cell from its staged K, |det J|: acc += |det J| (mu G^T + lambda G + mu tr(G) I), G = K^T M_ij K
  auto add_block = [&](const double* g, int i, int j, double (*acc)[BS]) {
    const double* m = M + (i * ND + j) * 9;
    double T[3][3], G[3][3];
#pragma unroll
    for (int al = 0; al < 3; ++al)
#pragma unroll
      for (int l = 0; l < 3; ++l) T[al][l] = m[al * 3] * g[l] + m[al * 3 + 1] * g[3 + l] + m[al * 3 + 2] * g[6 + l];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int l = 0; l < 3; ++l) G[k][l] = g[k] * T[0][l] + g[3 + k] * T[1][l] + g[6 + k] * T[2][l];
    const double dj = g[9], tr = mu * (G[0][0] + G[1][1] + G[2][2]);
#pragma unroll
    for (int a = 0; a < BS; ++a)
#pragma unroll
      for (int b = 0; b < BS; ++b) acc[a][b] += dj * (mu * G[b][a] + lmbda * G[a][b] + (a == b ? tr : 0.0));
  };
  const long long wstride = (long long)gridDim.x * 8;
  for (long long I = (long long)blockIdx.x * 8 + warp; I < P.nrows_b; I += wstride)
  {
    const int i0 = __ldg(P.inc_off + I), ninc = __ldg(P.inc_off + I + 1) - i0;
    const long long r0 = __ldg(A.rp + BS * I);
    const int nb = (int)((__ldg(A.rp + BS * I + 1) - r0) / BS);
    const long long blk0 = r0 / (BS * BS);
    const int kd = __ldg(P.diag + I);
    bool bcr[BS];
#pragma unroll
    for (int a = 0; a < BS; ++a) bcr[a] = bc ? bc[BS * I + a] != 0 : false;
    for (int cc = 0; cc < nb; cc += 32)
    {
      const int kc = cc + lane;
      double acc[BS][BS];
#pragma unroll
      for (int a = 0; a < BS; ++a)
#pragma unroll
        for (int b = 0; b < BS; ++b) acc[a][b] = 0.0;
      unsigned c_lo = 0, c_hi = 0;
      if (kc < nb) { c_lo = __ldg(P.con_off + blk0 + kc); c_hi = __ldg(P.con_off + blk0 + kc + 1); }
      const bool diag_here = kd >= cc && kd < cc + 32;
      for (int ch = 0; ch < ninc || ch == 0; ch += 32)
      {
        double dg[BS][BS];
#pragma unroll
        for (int a = 0; a < BS; ++a)
#pragma unroll
          for (int b = 0; b < BS; ++b) dg[a][b] = 0.0;
        __syncwarp();
        if (ch + lane < ninc)
        {
          const unsigned w = __ldg(P.inc + i0 + ch + lane);
          const long long idx = w / ND;
          const int il = (int)(w - idx * ND);
          const int cell = in.cells ? __ldg(in.cells + idx) : (int)idx;
          int xd[4];
#pragma unroll
          for (int v = 0; v < 4; ++v) xd[v] = __ldg(mesh.xd + (long long)cell * 4 + v);
          double X[4][3];
          load_vertices<3>(mesh, xd, X);
          P1Geom<3> Gm;
          p1_geometry<3>(X, Gm);  // rows of J^-1 = gradients of the barycentric coordinates 1..3; |det J| = 6 vol
          double* g = geo + lane * GS;
#pragma unroll
          for (int al = 0; al < 3; ++al)
#pragma unroll
            for (int k = 0; k < 3; ++k) g[al * 3 + k] = Gm.g[al + 1][k];
          g[9] = 6.0 * Gm.vol;
          if (diag_here) add_block(g, il, il, dg);  // reads back this lane's own stores
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
        if (kc < nb && kc != kd)
          for (unsigned q = c_lo; q < c_hi; ++q)
          {
            const unsigned cw = __ldg(P.con + q);
            const int kl = (int)(cw >> 8) - ch, il = (int)((cw >> 4) & 15u), j = (int)(cw & 15u);
            if (kl < 0 || kl >= 32) continue;
            add_block(geo + kl * GS, il, j, acc);
          }
      }
      if (kc < nb)
      {
        const int J = __ldg(A.col + r0 + (long long)BS * kc) / BS;
#pragma unroll
        for (int a = 0; a < BS; ++a)
        {
          double* dst = A.val + __ldg(A.rp + BS * I + a) + (long long)BS * kc;
#pragma unroll
          for (int b = 0; b < BS; ++b)
          {
            const bool z = bcr[a] || (bc && bc[BS * J + b] != 0);
            dst[b] = z ? 0.0 : acc[a][b];
          }
        }
      }
    }
  }
}
}  // namespace
