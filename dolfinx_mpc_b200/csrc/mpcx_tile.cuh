// mpcx_tile.cuh -- atomic-free "tile" assembly of the bulk (constraint-free) cells.
//
// The reference inserts every element matrix with MatSetValuesBlockedLocal (cpp/assemble_matrix.cpp:546), i.e.
// a search + add per entry.  On a GPU the equivalent scatter needs one red.global.add.f64 per entry, and at 16
// entries per P1 tetrahedron the kernel is bound by the ~1.3 cycles/lane issue rate of RED, not by HBM
// (profiles/r01_a: 7.3 ms, 15 % of the HBM roofline).  This file replaces the scatter by a gather:
//
//   setup (once per pattern / dofmap / bc set, mpcx_tile_plan_create):
//     rows are ordered along a Morton curve through the mesh and cut into tiles of R consecutive rows; a tile
//     owns the CSR entries ("dests") of its rows and lists the cells touching them, their vertices, and for
//     every dest the element-matrix entries ("sources") that sum into it;
//   assembly (k_tile_matrix_p1, one CTA per tile):
//     phase 0  vertex coordinates of the tile (gathered once per tile, not once per cell) and the source
//              lists go to shared memory,
//     phase 1  one thread per tile cell evaluates the element matrix into a shared-memory element buffer,
//     phase 2  one thread per dest sums its sources from the buffer and stores the CSR value -- plain
//              stores, no atomics, no zero-fill of the matrix, and a fixed summation order (bit-reproducible).
//
// Cells on tile borders are evaluated by every tile they touch (about 1.3-1.5x the cells); that is the price
// of not communicating between CTAs.  Cells holding slave dofs are excluded here (the `skip` flags) and
// handled by the elimination kernel afterwards, which adds into the stored values.
#pragma once
#include <cub/cub.cuh>

namespace
{
struct TilePlan
{
  int nt = 0, R = 0, cap = 0;
  int max_nodes = 0, max_dests = 0, max_src = 0, max_cells = 0;
  int ne = 0, n0 = 0, n1 = 0, ng = 0;
  long long nrows = 0, total_cells = 0, total_nodes = 0, total_dests = 0, total_src = 0, bytes = 0;
  int *tile_cell_off = nullptr, *tile_node_off = nullptr;
  long long *tile_dest_off = nullptr, *tile_src_off = nullptr;
  int *node_ids = nullptr, *cell_pos = nullptr, *row_of_rank = nullptr;
  uint16_t *cell_nodes = nullptr, *dest_src_end = nullptr, *dest_row = nullptr, *dest_pos = nullptr, *src = nullptr;
};

struct TilePlanD  // what the kernel sees
{
  int R, cap, max_nodes;
  const int *tile_cell_off, *tile_node_off;
  const long long *tile_dest_off, *tile_src_off;
  const int *node_ids, *cell_pos, *row_of_rank;
  const uint16_t *cell_nodes, *dest_src_end, *dest_row, *dest_pos, *src;
};

// ------------------------------------------------------------------ setup kernels (cold path)
__device__ __forceinline__ unsigned long long enc_f64(double d)
{
  const unsigned long long u = (unsigned long long)__double_as_longlong(d);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
inline double dec_f64(unsigned long long e)
{
  const unsigned long long u = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
  double d;
  memcpy(&d, &u, 8);
  return d;
}

__global__ void k_tp_bbox(const double* __restrict__ x, int xs, long long nn, unsigned long long* __restrict__ mm)
{
  unsigned long long lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0, 0, 0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += (long long)gridDim.x * blockDim.x)
    for (int k = 0; k < 3; ++k)
    {
      const unsigned long long e = enc_f64(x[i * xs + k]);
      lo[k] = e < lo[k] ? e : lo[k];
      hi[k] = e > hi[k] ? e : hi[k];
    }
  for (int k = 0; k < 3; ++k)
  {
    atomicMin(mm + k, lo[k]);
    atomicMax(mm + 3 + k, hi[k]);
  }
}

__device__ __forceinline__ unsigned long long spread21(unsigned long long v)
{
  v &= 0x1fffffull;
  v = (v | v << 32) & 0x1f00000000ffffull;
  v = (v | v << 16) & 0x1f0000ff0000ffull;
  v = (v | v << 8) & 0x100f00f00f00f00full;
  v = (v | v << 4) & 0x10c30c30c30c30c3ull;
  v = (v | v << 2) & 0x1249249249249249ull;
  return v;
}

struct BBox
{
  double lo[3], inv[3];
};

// row key = smallest Morton code (cell centroid) over the non-skipped cells touching the block row
__global__ void k_tp_rowkeys(MeshD mesh, BBox bb, const int* __restrict__ dm0, int nd0, const int* __restrict__ cells,
                             long long nc, const int8_t* __restrict__ skip, unsigned long long* __restrict__ rowkey)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc || (skip && skip[i])) return;
  const int cell = cells ? cells[i] : (int)i;
  double c[3] = {0, 0, 0};
  for (int g = 0; g < mesh.ng; ++g)
  {
    const double* p = mesh.x + (long long)mesh.xd[(long long)cell * mesh.ng + g] * mesh.xs;
    for (int k = 0; k < 3; ++k) c[k] += p[k];
  }
  unsigned long long code = 0;
  for (int k = 0; k < 3; ++k)
  {
    double u = (c[k] / mesh.ng - bb.lo[k]) * bb.inv[k];
    u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
    code |= spread21((unsigned long long)(u * 2097151.0)) << k;
  }
  for (int a = 0; a < nd0; ++a) atomicMin(rowkey + dm0[(long long)cell * nd0 + a], code);
}

__global__ void k_tp_iota(int* __restrict__ a, long long n)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = (int)i;
}

// rank of every block row and the scalar row of every scalar rank
__global__ void k_tp_ranks(const int* __restrict__ order, long long nbr, int bs0, int* __restrict__ rank,
                           int* __restrict__ row_of_rank)
{
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nbr) return;
  const int rb = order[k];
  rank[rb] = (int)k;
  for (int a = 0; a < bs0; ++a) row_of_rank[k * bs0 + a] = rb * bs0 + a;
}

#define MPCX_TP_MAXT 64
// distinct tiles touched by the rows of one cell
__device__ __forceinline__ int cell_tiles(const int* __restrict__ d0, int nd0, int bs0, const int* __restrict__ rank,
                                          int R, int* out)
{
  int n = 0;
  for (int a = 0; a < nd0; ++a)
  {
    const long long s = (long long)rank[d0[a]] * bs0;
    const int t0 = (int)(s / R), t1 = (int)((s + bs0 - 1) / R);
    for (int t = t0; t <= t1; ++t)
    {
      bool seen = false;
      for (int k = 0; k < n; ++k) seen |= out[k] == t;
      if (!seen && n < MPCX_TP_MAXT) out[n++] = t;
    }
  }
  return n;
}

__global__ void k_tp_count_pairs(const int* __restrict__ dm0, int nd0, int bs0, const int* __restrict__ cells,
                                 long long nc, const int8_t* __restrict__ skip, const int* __restrict__ rank, int R,
                                 long long* __restrict__ count)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  int tl[MPCX_TP_MAXT];
  const int cell = cells ? cells[i] : (int)i;
  count[i] = (skip && skip[i]) ? 0 : cell_tiles(dm0 + (long long)cell * nd0, nd0, bs0, rank, R, tl);
}

__global__ void k_tp_fill_pairs(const int* __restrict__ dm0, int nd0, int bs0, const int* __restrict__ cells,
                                long long nc, const int8_t* __restrict__ skip, const int* __restrict__ rank, int R,
                                const long long* __restrict__ off, unsigned long long* __restrict__ keys)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc || (skip && skip[i])) return;
  int tl[MPCX_TP_MAXT];
  const int cell = cells ? cells[i] : (int)i;
  const int n = cell_tiles(dm0 + (long long)cell * nd0, nd0, bs0, rank, R, tl);
  for (int k = 0; k < n; ++k) keys[off[i] + k] = ((unsigned long long)tl[k] << 32) | (unsigned long long)i;
}

// off[t] = first index with (key >> 32) >= t, for t = 0..nt
__global__ void k_tp_tile_offsets(const unsigned long long* __restrict__ keys, long long n, int nt, int* __restrict__ off)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > nt) return;
  const unsigned long long target = (unsigned long long)t << 32;
  long long lo = 0, hi = n;
  while (lo < hi)
  {
    const long long mid = (lo + hi) >> 1;
    if (keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  off[t] = (int)lo;
}

__global__ void k_tp_node_pairs(const unsigned long long* __restrict__ ckeys, long long ntc, MeshD mesh,
                                const int* __restrict__ cells, unsigned long long* __restrict__ nkeys,
                                int* __restrict__ cell_pos)
{
  const long long tc = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tc >= ntc) return;
  const unsigned long long key = ckeys[tc];
  const int i = (int)(key & 0xffffffffull);
  cell_pos[tc] = i;
  const int cell = cells ? cells[i] : i;
  for (int g = 0; g < mesh.ng; ++g)
    nkeys[tc * mesh.ng + g] = (key & 0xffffffff00000000ull) | (unsigned long long)(unsigned)mesh.xd[(long long)cell * mesh.ng + g];
}

__global__ void k_tp_cell_nodes(const unsigned long long* __restrict__ ckeys, long long ntc, MeshD mesh,
                                const int* __restrict__ cells, const unsigned long long* __restrict__ ukeys,
                                const int* __restrict__ tile_node_off, int* __restrict__ node_ids, long long nun,
                                uint16_t* __restrict__ cell_nodes)
{
  const long long tc = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tc < nun) node_ids[tc] = (int)(ukeys[tc] & 0xffffffffull);
  if (tc >= ntc) return;
  const unsigned long long key = ckeys[tc];
  const int t = (int)(key >> 32), i = (int)(key & 0xffffffffull);
  const int cell = cells ? cells[i] : i;
  const int b = tile_node_off[t], e = tile_node_off[t + 1];
  for (int g = 0; g < mesh.ng; ++g)
  {
    const unsigned long long target = (key & 0xffffffff00000000ull) | (unsigned long long)(unsigned)mesh.xd[(long long)cell * mesh.ng + g];
    int lo = b, hi = e;
    while (lo < hi)
    {
      const int mid = (lo + hi) >> 1;
      if (ukeys[mid] < target) lo = mid + 1; else hi = mid;
    }
    cell_nodes[tc * mesh.ng + g] = (uint16_t)(lo - b);
  }
}

__global__ void k_tp_row_len(const int* __restrict__ row_of_rank, long long nrows, const long long* __restrict__ rp,
                             long long* __restrict__ len)
{
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nrows) len[k] = rp[row_of_rank[k] + 1] - rp[row_of_rank[k]];
}

__global__ void k_tp_tile_dest_off(const long long* __restrict__ dstart, long long nrows, long long nnz, int nt, int R,
                                   long long* __restrict__ off)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > nt) return;
  const long long k = (long long)t * R;
  off[t] = k >= nrows ? nnz : dstart[k];
}

// visits every (tile cell, p, q) whose row belongs to the tile and that survives the bc zeroing
// (cpp/assemble_matrix.cpp:513-533); F(dest index in rank order, element entry e)
template <typename F>
__device__ __forceinline__ void for_each_source(const unsigned long long key, const int* __restrict__ dm0,
                                                const int* __restrict__ dm1, int nd0, int nd1, int bs0, int bs1,
                                                const int* __restrict__ cells, const int* __restrict__ rank, int R,
                                                const long long* __restrict__ dstart, const int8_t* __restrict__ bc0,
                                                const int8_t* __restrict__ bc1, const CsrD& A, F&& f)
{
  const int t = (int)(key >> 32), i = (int)(key & 0xffffffffull);
  const int cell = cells ? cells[i] : i;
  const int n1 = nd1 * bs1;
  for (int ib = 0; ib < nd0; ++ib)
  {
    const int rb = dm0[(long long)cell * nd0 + ib];
    for (int ia = 0; ia < bs0; ++ia)
    {
      const long long ks = (long long)rank[rb] * bs0 + ia;
      if ((int)(ks / R) != t) continue;
      const int r = rb * bs0 + ia;
      if (bc0 && bc0[r]) continue;
      for (int jb = 0; jb < nd1; ++jb)
        for (int ja = 0; ja < bs1; ++ja)
        {
          const int c = dm1[(long long)cell * nd1 + jb] * bs1 + ja;
          if (bc1 && bc1[c]) continue;
          const long long p = csr_find(A, r, c);
          if (p < 0) { g_dev_err = MPCX_ERR_PATTERN; continue; }
          f(dstart[ks] + (p - A.rp[r]), (ib * bs0 + ia) * n1 + jb * bs1 + ja);
        }
    }
  }
}

__global__ void k_tp_count_sources(const unsigned long long* __restrict__ ckeys, long long ntc, const int* __restrict__ dm0,
                                   const int* __restrict__ dm1, int nd0, int nd1, int bs0, int bs1,
                                   const int* __restrict__ cells, const int* __restrict__ rank, int R,
                                   const long long* __restrict__ dstart, const int8_t* __restrict__ bc0,
                                   const int8_t* __restrict__ bc1, CsrD A, unsigned* __restrict__ cnt)
{
  const long long tc = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tc >= ntc) return;
  for_each_source(ckeys[tc], dm0, dm1, nd0, nd1, bs0, bs1, cells, rank, R, dstart, bc0, bc1, A,
                  [&](long long d, int) { atomicAdd(cnt + d, 1u); });
}

// sort key of a dest: tile | (0xffff - count) | local index  -> per tile: descending count, then row order
__global__ void k_tp_dest_keys(const long long* __restrict__ dstart, long long nrows, long long nnz, int R,
                               const long long* __restrict__ tile_dest_off, const unsigned* __restrict__ cnt,
                               unsigned long long* __restrict__ keys)
{
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nrows) return;
  const int t = (int)(k / R);
  const long long b = dstart[k], e = (k + 1 < nrows) ? dstart[k + 1] : nnz, toff = tile_dest_off[t];
  for (long long d = b; d < e; ++d)
  {
    const unsigned c = cnt[d] > 0xffffu ? 0xffffu : cnt[d];
    keys[d] = ((unsigned long long)t << 32) | ((unsigned long long)(0xffffu - c) << 16) | (unsigned long long)(d - toff);
  }
}

// after the sort: position j holds the dest with local index (key & 0xffff) of tile (key >> 32)
__global__ void k_tp_unpack_dests(const unsigned long long* __restrict__ keys, long long nnz, int R, long long nrows,
                                  const long long* __restrict__ dstart, const long long* __restrict__ tile_dest_off,
                                  const unsigned* __restrict__ cnt, uint16_t* __restrict__ newidx,
                                  uint16_t* __restrict__ dest_row, uint16_t* __restrict__ dest_pos,
                                  long long* __restrict__ cnt_sorted)
{
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  const unsigned long long key = keys[j];
  const int t = (int)(key >> 32);
  const long long toff = tile_dest_off[t];
  const long long d = toff + (long long)(key & 0xffffull);
  // row of dest d: last rank k in [tR, tR + R) with dstart[k] <= d
  long long lo = (long long)t * R, hi = lo + R < nrows ? lo + R : nrows;
  while (hi - lo > 1)
  {
    const long long mid = (lo + hi) >> 1;
    if (dstart[mid] <= d) lo = mid; else hi = mid;
  }
  dest_row[j] = (uint16_t)(lo - (long long)t * R);
  dest_pos[j] = (uint16_t)(d - dstart[lo]);
  newidx[d] = (uint16_t)(j - toff);
  cnt_sorted[j] = cnt[d];
}

// per tile: end offsets of the dest segments (relative to the tile's source segment), padded totals
__global__ void k_tp_tile_src_totals(const long long* __restrict__ incl, const long long* __restrict__ tile_dest_off, int nt,
                                     long long* __restrict__ totals)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nt) return;
  const long long b = tile_dest_off[t], e = tile_dest_off[t + 1];
  const long long tot = e > b ? incl[e - 1] - (b > 0 ? incl[b - 1] : 0) : 0;
  totals[t] = (tot + 7) & ~7ll;  // 16-byte aligned source segments
}

__global__ void k_tp_dest_src_end(const long long* __restrict__ incl, const unsigned long long* __restrict__ keys,
                                  const long long* __restrict__ tile_dest_off, long long nnz,
                                  uint16_t* __restrict__ dest_src_end)
{
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  const long long b = tile_dest_off[(int)(keys[j] >> 32)];
  dest_src_end[j] = (uint16_t)(incl[j] - (b > 0 ? incl[b - 1] : 0));
}

__global__ void k_tp_fill_sources(const unsigned long long* __restrict__ ckeys, long long ntc, const int* __restrict__ dm0,
                                  const int* __restrict__ dm1, int nd0, int nd1, int bs0, int bs1,
                                  const int* __restrict__ cells, const int* __restrict__ rank, int R,
                                  const long long* __restrict__ dstart, const int8_t* __restrict__ bc0,
                                  const int8_t* __restrict__ bc1, CsrD A, unsigned* __restrict__ cnt,
                                  const int* __restrict__ tile_cell_off, const long long* __restrict__ tile_dest_off,
                                  const long long* __restrict__ tile_src_off, const uint16_t* __restrict__ newidx,
                                  const uint16_t* __restrict__ dest_src_end, int cap, uint16_t* __restrict__ src)
{
  const long long tc = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (tc >= ntc) return;
  const unsigned long long key = ckeys[tc];
  const int t = (int)(key >> 32);
  const int cl = (int)(tc - tile_cell_off[t]);
  const long long toff = tile_dest_off[t], soff = tile_src_off[t];
  for_each_source(key, dm0, dm1, nd0, nd1, bs0, bs1, cells, rank, R, dstart, bc0, bc1, A, [&](long long d, int e) {
    const int slot = (int)atomicSub(cnt + d, 1u) - 1;
    const int nj = newidx[d];
    const int beg = nj ? dest_src_end[toff + nj - 1] : 0;
    src[soff + beg + slot] = (uint16_t)(e * cap + cl);
  });
}

// fixed summation order: ascending source id inside every dest
__global__ void k_tp_sort_sources(const unsigned long long* __restrict__ keys, long long nnz,
                                  const long long* __restrict__ tile_dest_off, const long long* __restrict__ tile_src_off,
                                  const uint16_t* __restrict__ dest_src_end, uint16_t* __restrict__ src)
{
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nnz) return;
  const int t = (int)(keys[j] >> 32);
  const long long toff = tile_dest_off[t];
  const int beg = j > toff ? dest_src_end[j - 1] : 0, end = dest_src_end[j];
  uint16_t* s = src + tile_src_off[t];
  for (int a = beg + 1; a < end; ++a)
  {
    const uint16_t v = s[a];
    int b = a - 1;
    while (b >= beg && s[b] > v) { s[b + 1] = s[b]; --b; }
    s[b + 1] = v;
  }
}

// ------------------------------------------------------------------ the assembly kernel
#define MPCX_TILE_THREADS 256

template <int TD>
__global__ void __launch_bounds__(MPCX_TILE_THREADS)
k_tile_matrix_p1(TilePlanD P, IntD in, MeshD mesh, CsrD A, int accumulate)
{
  constexpr int NV = TD + 1, NE = NV * NV;
  extern __shared__ __align__(16) unsigned char tile_smem[];
  double* Xs = reinterpret_cast<double*>(tile_smem);      // [max_nodes][3]
  double* ebuf = Xs + 3 * P.max_nodes;                    // [NE][cap]
  uint16_t* ssrc = reinterpret_cast<uint16_t*>(ebuf + NE * P.cap);
  const int t = blockIdx.x, tid = threadIdx.x;
  const int c0 = P.tile_cell_off[t], nc_t = P.tile_cell_off[t + 1] - c0;
  const int n0 = P.tile_node_off[t], nn_t = P.tile_node_off[t + 1] - n0;
  const long long d0 = P.tile_dest_off[t];
  const int nd_t = (int)(P.tile_dest_off[t + 1] - d0);
  const long long s0 = P.tile_src_off[t];
  const int ns_t = nd_t ? P.dest_src_end[d0 + nd_t - 1] : 0;

  // phase 0: vertex coordinates and source lists of the tile -> shared memory
  for (int i = tid; i < nn_t; i += MPCX_TILE_THREADS)
  {
    const double* p = mesh.x + (long long)__ldg(P.node_ids + n0 + i) * mesh.xs;
    Xs[3 * i] = __ldg(p);
    Xs[3 * i + 1] = __ldg(p + 1);
    Xs[3 * i + 2] = __ldg(p + 2);
  }
  {
    const uint4* g = reinterpret_cast<const uint4*>(P.src + s0);
    uint4* s = reinterpret_cast<uint4*>(ssrc);
    for (int i = tid; i < (ns_t + 7) / 8; i += MPCX_TILE_THREADS) s[i] = __ldg(g + i);
  }
  __syncthreads();

  // phase 1: element matrices of the tile's cells -> element buffer, entry-major (conflict-free stores)
  for (int cl = tid; cl < nc_t; cl += MPCX_TILE_THREADS)
  {
    int lnode[NV];
    if (NV == 4)
    {
      const uint2 pk = __ldg(reinterpret_cast<const uint2*>(P.cell_nodes) + (c0 + cl));
      lnode[0] = pk.x & 0xffff; lnode[1] = pk.x >> 16; lnode[2] = pk.y & 0xffff; lnode[NV - 1] = pk.y >> 16;
    }
    else
    {
#pragma unroll
      for (int v = 0; v < NV; ++v) lnode[v] = __ldg(P.cell_nodes + (long long)(c0 + cl) * NV + v);
    }
    double X[NV][3];
#pragma unroll
    for (int v = 0; v < NV; ++v)
    {
      const int l = lnode[v];
      X[v][0] = Xs[3 * l];
      X[v][1] = Xs[3 * l + 1];
      X[v][2] = TD == 3 ? Xs[3 * l + 2] : 0.0;
    }
    P1Geom<TD> G;
    p1_geometry<TD>(X, G);
    double w[NV], Ae[NV][NV];
    if (in.kernel == MPCX_KERNEL_LAPLACE_VARCOEF)
    {
      const long long index = __ldg(P.cell_pos + c0 + cl);
      p1_load_w<TD>(in, index, in.cells ? __ldg(in.cells + index) : (int)index, w);
    }
    p1_element<TD>(in.kernel, G, in.c, w, Ae);
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < NV; ++j) ebuf[(i * NV + j) * P.cap + cl] = Ae[i][j];
  }
  __syncthreads();

  // phase 2: one thread per CSR entry of the tile sums its sources; plain store
  for (int k = tid; k < nd_t; k += MPCX_TILE_THREADS)
  {
    const int beg = k ? __ldg(P.dest_src_end + d0 + k - 1) : 0, end = __ldg(P.dest_src_end + d0 + k);
    double sum = 0.0;
    for (int p = beg; p < end; ++p) sum += ebuf[ssrc[p]];
    const int row = __ldg(P.row_of_rank + (long long)t * P.R + __ldg(P.dest_row + d0 + k));
    const long long out = __ldg(A.rp + row) + __ldg(P.dest_pos + d0 + k);
    if (accumulate) A.val[out] += sum; else A.val[out] = sum;
  }
}

// ------------------------------------------------------------------ host side of the setup
#define TP_CK(call)                                                      \
  do {                                                                   \
    cudaError_t e__ = (call);                                            \
    if (e__ != cudaSuccess) { rc = cuda_check(e__, #call); goto done; }  \
  } while (0)

template <typename T>
cudaError_t tp_alloc(T** p, long long n)
{
  return cudaMalloc((void**)p, sizeof(T) * (size_t)(n > 0 ? n : 1));
}

void tile_plan_free(TilePlan* P)
{
  if (!P) return;
  cudaFree(P->tile_cell_off); cudaFree(P->tile_node_off); cudaFree(P->tile_dest_off); cudaFree(P->tile_src_off);
  cudaFree(P->node_ids); cudaFree(P->cell_pos); cudaFree(P->row_of_rank); cudaFree(P->cell_nodes);
  cudaFree(P->dest_src_end); cudaFree(P->dest_row); cudaFree(P->dest_pos); cudaFree(P->src);
  delete P;
}

inline unsigned tp_grid(long long n, int b = 256) { return (unsigned)((n + b - 1) / b > 0 ? (n + b - 1) / b : 1); }

int tile_plan_build(const mpcx_mesh* mesh, const mpcx_dofmap* dm0, const mpcx_dofmap* dm1, const int32_t* cells,
                    long long nc, const int8_t* skip, const int8_t* bc0, const int8_t* bc1, const mpcx_csr* Acsr,
                    int max_cells, int max_rows, cudaStream_t s, TilePlan** out)
{
  int rc = MPCX_OK;
  TilePlan* P = new TilePlan();
  const MeshD md{mesh->x, mesh->x_dofmap, mesh->ng, mesh->x_stride};
  const CsrD A{(const long long*)Acsr->row_ptr, Acsr->col, Acsr->val};
  const int nd0 = dm0->nd, nd1 = dm1->nd, bs0 = dm0->bs, bs1 = dm1->bs;
  const long long nrows = Acsr->num_rows, nnz = Acsr->nnz, nbr = nrows / bs0;
  P->n0 = nd0 * bs0; P->n1 = nd1 * bs1; P->ne = P->n0 * P->n1; P->ng = mesh->ng; P->nrows = nrows;
  // scratch
  unsigned long long *mm = nullptr, *rowkey = nullptr, *rowkey2 = nullptr, *ckeys = nullptr, *ckeys2 = nullptr,
                     *nkeys = nullptr, *nkeys2 = nullptr, *dkeys = nullptr, *dkeys2 = nullptr;
  int *iota = nullptr, *order = nullptr, *rank = nullptr;
  long long *pcount = nullptr, *poff = nullptr, *dstart = nullptr, *cnt_sorted = nullptr, *totals = nullptr, *nsel = nullptr;
  unsigned* cnt = nullptr;
  uint16_t* newidx = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  long long npairs = 0, nun = 0, ntc = 0;
  std::vector<int> h_off;
  std::vector<long long> h_doff;
  BBox bb;
  auto need_tmp = [&](size_t bytes) -> cudaError_t {
    if (bytes <= tmp_bytes) return cudaSuccess;
    cudaFree(tmp);
    tmp = nullptr;
    tmp_bytes = bytes + bytes / 8;
    return cudaMalloc(&tmp, tmp_bytes);
  };
  size_t tb = 0;

  if (P->ne * 32 > 65536 || mesh->ng > 16 || nd0 > 32) { rc = fail(MPCX_ERR_UNSUPPORTED, "element too large for the tile plan"); goto done; }
  if (max_cells <= 0) max_cells = 640;
  if (max_cells * P->ne > 65535) max_cells = 65535 / P->ne;
  if (max_rows <= 0) max_rows = 4096;
  if (max_rows > 65535) max_rows = 65535;

  // 1. bounding box -> Morton quantisation
  {
    unsigned long long h[6] = {~0ull, ~0ull, ~0ull, 0, 0, 0};
    TP_CK(tp_alloc(&mm, 6));
    TP_CK(cudaMemcpyAsync(mm, h, sizeof(h), cudaMemcpyHostToDevice, s));
    k_tp_bbox<<<148 * 8, 256, 0, s>>>(mesh->x, mesh->x_stride, mesh->num_nodes, mm);
    TP_CK(cudaMemcpyAsync(h, mm, sizeof(h), cudaMemcpyDeviceToHost, s));
    TP_CK(cudaStreamSynchronize(s));
    for (int k = 0; k < 3; ++k)
    {
      const double lo = dec_f64(h[k]), hi = dec_f64(h[3 + k]);
      bb.lo[k] = lo;
      bb.inv[k] = hi > lo ? 1.0 / (hi - lo) : 0.0;
    }
  }
  // 2. row keys, 3. rank of every row along the curve
  TP_CK(tp_alloc(&rowkey, nbr)); TP_CK(tp_alloc(&rowkey2, nbr));
  TP_CK(tp_alloc(&iota, nbr)); TP_CK(tp_alloc(&order, nbr)); TP_CK(tp_alloc(&rank, nbr));
  TP_CK(tp_alloc(&P->row_of_rank, nrows));
  TP_CK(cudaMemsetAsync(rowkey, 0xff, sizeof(unsigned long long) * (size_t)nbr, s));
  k_tp_rowkeys<<<tp_grid(nc), 256, 0, s>>>(md, bb, dm0->map, nd0, cells, nc, skip, rowkey);
  k_tp_iota<<<tp_grid(nbr), 256, 0, s>>>(iota, nbr);
  TP_CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, rowkey, rowkey2, iota, order, (int)nbr, 0, 64, s));
  TP_CK(need_tmp(tb));
  TP_CK(cub::DeviceRadixSort::SortPairs(tmp, tb, rowkey, rowkey2, iota, order, (int)nbr, 0, 64, s));
  k_tp_ranks<<<tp_grid(nbr), 256, 0, s>>>(order, nbr, bs0, rank, P->row_of_rank);
  cudaFree(rowkey); rowkey = nullptr; cudaFree(rowkey2); rowkey2 = nullptr; cudaFree(iota); iota = nullptr;

  // 4./5. rows per tile: start from an estimate, shrink until every tile fits the element buffer
  TP_CK(tp_alloc(&pcount, nc + 1)); TP_CK(tp_alloc(&poff, nc + 1));
  TP_CK(tp_alloc(&dstart, nrows + 1));
  k_tp_row_len<<<tp_grid(nrows), 256, 0, s>>>(P->row_of_rank, nrows, A.rp, dstart);
  TP_CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, dstart, dstart, (int)nrows, s));
  TP_CK(need_tmp(tb));
  TP_CK(cub::DeviceScan::ExclusiveSum(tmp, tb, dstart, dstart, (int)nrows, s));
  {
    // cells per row ~ nc * nd0 / nbr incident, each shared by nd0 rows: a compact tile of R rows holds about
    // R * nc / nbr cells plus a halo; start optimistic and let the loop below correct it
    double cells_per_row = (double)nc * bs0 > 0 ? (double)nc / (double)(nrows > 0 ? nrows : 1) : 1.0;
    long long R = (long long)(0.6 * max_cells / (cells_per_row > 1e-9 ? cells_per_row : 1.0));
    if (R < 1) R = 1;
    if (R > max_rows) R = max_rows;
    for (int iter = 0; iter < 12; ++iter)
    {
      P->R = (int)R;
      P->nt = (int)((nrows + R - 1) / R);
      if (P->nt < 1) P->nt = 1;
      TP_CK(cudaMemsetAsync(pcount, 0, sizeof(long long) * (size_t)(nc + 1), s));
      k_tp_count_pairs<<<tp_grid(nc), 256, 0, s>>>(dm0->map, nd0, bs0, cells, nc, skip, rank, P->R, pcount);
      TP_CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, pcount, poff, (int)(nc + 1), s));
      TP_CK(need_tmp(tb));
      TP_CK(cub::DeviceScan::ExclusiveSum(tmp, tb, pcount, poff, (int)(nc + 1), s));
      TP_CK(cudaMemcpyAsync(&npairs, poff + nc, sizeof(long long), cudaMemcpyDeviceToHost, s));
      TP_CK(cudaStreamSynchronize(s));
      if (npairs >= (1ll << 31)) { rc = fail(MPCX_ERR_UNSUPPORTED, "tile plan: more than 2^31 tile cells"); goto done; }
      cudaFree(ckeys); cudaFree(ckeys2); ckeys = ckeys2 = nullptr;
      TP_CK(tp_alloc(&ckeys, npairs)); TP_CK(tp_alloc(&ckeys2, npairs));
      k_tp_fill_pairs<<<tp_grid(nc), 256, 0, s>>>(dm0->map, nd0, bs0, cells, nc, skip, rank, P->R, poff, ckeys);
      TP_CK(cub::DeviceRadixSort::SortKeys(nullptr, tb, ckeys, ckeys2, (int)npairs, 0, 64, s));
      TP_CK(need_tmp(tb));
      TP_CK(cub::DeviceRadixSort::SortKeys(tmp, tb, ckeys, ckeys2, (int)npairs, 0, 64, s));
      cudaFree(P->tile_cell_off); P->tile_cell_off = nullptr;
      TP_CK(tp_alloc(&P->tile_cell_off, P->nt + 1));
      k_tp_tile_offsets<<<tp_grid(P->nt + 1), 256, 0, s>>>(ckeys2, npairs, P->nt, P->tile_cell_off);
      cudaFree(P->tile_dest_off); P->tile_dest_off = nullptr;
      TP_CK(tp_alloc(&P->tile_dest_off, P->nt + 1));
      k_tp_tile_dest_off<<<tp_grid(P->nt + 1), 256, 0, s>>>(dstart, nrows, nnz, P->nt, P->R, P->tile_dest_off);
      h_off.resize(P->nt + 1); h_doff.resize(P->nt + 1);
      TP_CK(cudaMemcpyAsync(h_off.data(), P->tile_cell_off, sizeof(int) * (P->nt + 1), cudaMemcpyDeviceToHost, s));
      TP_CK(cudaMemcpyAsync(h_doff.data(), P->tile_dest_off, sizeof(long long) * (P->nt + 1), cudaMemcpyDeviceToHost, s));
      TP_CK(cudaStreamSynchronize(s));
      int mc = 0;
      long long mdst = 0;
      for (int t = 0; t < P->nt; ++t)
      {
        mc = std::max(mc, h_off[t + 1] - h_off[t]);
        mdst = std::max(mdst, h_doff[t + 1] - h_doff[t]);
      }
      P->max_cells = mc; P->max_dests = (int)mdst;
      if ((mc <= max_cells && mdst <= 65535) || R == 1) break;
      double shrink = std::min((double)max_cells / (double)std::max(mc, 1), 65535.0 / (double)std::max<long long>(mdst, 1));
      long long Rn = (long long)(R * shrink * 0.95);
      if (Rn >= R) Rn = R - 1;
      R = Rn < 1 ? 1 : Rn;
    }
    if (P->max_cells > 65535 / P->ne || P->max_dests > 65535)
    { rc = fail(MPCX_ERR_UNSUPPORTED, "tile plan: a single row touches more cells than the element buffer holds"); goto done; }
  }
  ntc = npairs;
  P->total_cells = ntc;
  P->cap = (P->max_cells + 31) & ~31;
  if (P->cap * P->ne > 65536) P->cap = 65536 / P->ne;
  cudaFree(pcount); pcount = nullptr; cudaFree(poff); poff = nullptr; cudaFree(ckeys); ckeys = nullptr;

  // 6. vertices of every tile and tile-local vertex ids of every tile cell
  TP_CK(tp_alloc(&nkeys, ntc * md.ng)); TP_CK(tp_alloc(&nkeys2, ntc * md.ng)); TP_CK(tp_alloc(&P->cell_pos, ntc));
  k_tp_node_pairs<<<tp_grid(ntc), 256, 0, s>>>(ckeys2, ntc, md, cells, nkeys, P->cell_pos);
  if (ntc * md.ng >= (1ll << 31)) { rc = fail(MPCX_ERR_UNSUPPORTED, "tile plan: too many tile vertices"); goto done; }
  TP_CK(cub::DeviceRadixSort::SortKeys(nullptr, tb, nkeys, nkeys2, (int)(ntc * md.ng), 0, 64, s));
  TP_CK(need_tmp(tb));
  TP_CK(cub::DeviceRadixSort::SortKeys(tmp, tb, nkeys, nkeys2, (int)(ntc * md.ng), 0, 64, s));
  TP_CK(tp_alloc(&nsel, 1));
  TP_CK(cub::DeviceSelect::Unique(nullptr, tb, nkeys2, nkeys, nsel, (int)(ntc * md.ng), s));
  TP_CK(need_tmp(tb));
  TP_CK(cub::DeviceSelect::Unique(tmp, tb, nkeys2, nkeys, nsel, (int)(ntc * md.ng), s));
  TP_CK(cudaMemcpyAsync(&nun, nsel, sizeof(long long), cudaMemcpyDeviceToHost, s));
  TP_CK(cudaStreamSynchronize(s));
  P->total_nodes = nun;
  TP_CK(tp_alloc(&P->tile_node_off, P->nt + 1));
  k_tp_tile_offsets<<<tp_grid(P->nt + 1), 256, 0, s>>>(nkeys, nun, P->nt, P->tile_node_off);
  TP_CK(tp_alloc(&P->node_ids, nun)); TP_CK(tp_alloc(&P->cell_nodes, ntc * md.ng));
  k_tp_cell_nodes<<<tp_grid(std::max(ntc, nun)), 256, 0, s>>>(ckeys2, ntc, md, cells, nkeys, P->tile_node_off, P->node_ids, nun,
                                                             P->cell_nodes);
  TP_CK(cudaMemcpyAsync(h_off.data(), P->tile_node_off, sizeof(int) * (P->nt + 1), cudaMemcpyDeviceToHost, s));
  TP_CK(cudaStreamSynchronize(s));
  for (int t = 0; t < P->nt; ++t) P->max_nodes = std::max(P->max_nodes, h_off[t + 1] - h_off[t]);
  if (P->max_nodes > 65535) { rc = fail(MPCX_ERR_UNSUPPORTED, "tile plan: too many vertices in one tile"); goto done; }
  P->max_nodes = (P->max_nodes + 1) & ~1;  // keeps the shared-memory sections 16-byte aligned
  cudaFree(nkeys); nkeys = nullptr; cudaFree(nkeys2); nkeys2 = nullptr;

  // 7.-9. sources per dest, dests ordered by descending count inside every tile
  TP_CK(tp_alloc(&cnt, nnz));
  TP_CK(cudaMemsetAsync(cnt, 0, sizeof(unsigned) * (size_t)nnz, s));
  k_tp_count_sources<<<tp_grid(ntc, 128), 128, 0, s>>>(ckeys2, ntc, dm0->map, dm1->map, nd0, nd1, bs0, bs1, cells, rank, P->R,
                                                      dstart, bc0, bc1, A, cnt);
  TP_CK(tp_alloc(&dkeys, nnz)); TP_CK(tp_alloc(&dkeys2, nnz));
  k_tp_dest_keys<<<tp_grid(nrows), 256, 0, s>>>(dstart, nrows, nnz, P->R, P->tile_dest_off, cnt, dkeys);
  if (nnz >= (1ll << 31)) { rc = fail(MPCX_ERR_UNSUPPORTED, "tile plan: nnz >= 2^31 on one device"); goto done; }
  TP_CK(cub::DeviceRadixSort::SortKeys(nullptr, tb, dkeys, dkeys2, (int)nnz, 0, 64, s));
  TP_CK(need_tmp(tb));
  TP_CK(cub::DeviceRadixSort::SortKeys(tmp, tb, dkeys, dkeys2, (int)nnz, 0, 64, s));
  cudaFree(dkeys); dkeys = nullptr;
  TP_CK(tp_alloc(&newidx, nnz)); TP_CK(tp_alloc(&P->dest_row, nnz)); TP_CK(tp_alloc(&P->dest_pos, nnz));
  TP_CK(tp_alloc(&cnt_sorted, nnz)); TP_CK(tp_alloc(&P->dest_src_end, nnz));
  k_tp_unpack_dests<<<tp_grid(nnz), 256, 0, s>>>(dkeys2, nnz, P->R, nrows, dstart, P->tile_dest_off, cnt, newidx, P->dest_row,
                                                 P->dest_pos, cnt_sorted);
  TP_CK(cub::DeviceScan::InclusiveSum(nullptr, tb, cnt_sorted, cnt_sorted, (int)nnz, s));
  TP_CK(need_tmp(tb));
  TP_CK(cub::DeviceScan::InclusiveSum(tmp, tb, cnt_sorted, cnt_sorted, (int)nnz, s));
  TP_CK(tp_alloc(&totals, P->nt + 1)); TP_CK(tp_alloc(&P->tile_src_off, P->nt + 1));
  TP_CK(cudaMemsetAsync(totals, 0, sizeof(long long) * (P->nt + 1), s));
  k_tp_tile_src_totals<<<tp_grid(P->nt), 256, 0, s>>>(cnt_sorted, P->tile_dest_off, P->nt, totals);
  TP_CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, totals, P->tile_src_off, P->nt + 1, s));
  TP_CK(need_tmp(tb));
  TP_CK(cub::DeviceScan::ExclusiveSum(tmp, tb, totals, P->tile_src_off, P->nt + 1, s));
  k_tp_dest_src_end<<<tp_grid(nnz), 256, 0, s>>>(cnt_sorted, dkeys2, P->tile_dest_off, nnz, P->dest_src_end);
  h_doff.resize(P->nt + 1);
  TP_CK(cudaMemcpyAsync(h_doff.data(), P->tile_src_off, sizeof(long long) * (P->nt + 1), cudaMemcpyDeviceToHost, s));
  TP_CK(cudaStreamSynchronize(s));
  P->total_src = h_doff[P->nt];
  for (int t = 0; t < P->nt; ++t) P->max_src = std::max<long long>(P->max_src, h_doff[t + 1] - h_doff[t]);
  if (P->max_src > 65535 + 8) { rc = fail(MPCX_ERR_UNSUPPORTED, "tile plan: source segment overflow"); goto done; }
  cudaFree(cnt_sorted); cnt_sorted = nullptr;

  // 10./11. source lists
  TP_CK(tp_alloc(&P->src, P->total_src + 8));
  TP_CK(cudaMemsetAsync(P->src, 0, sizeof(uint16_t) * (size_t)(P->total_src + 8), s));
  k_tp_fill_sources<<<tp_grid(ntc, 128), 128, 0, s>>>(ckeys2, ntc, dm0->map, dm1->map, nd0, nd1, bs0, bs1, cells, rank, P->R, dstart,
                                                     bc0, bc1, A, cnt, P->tile_cell_off, P->tile_dest_off, P->tile_src_off, newidx,
                                                     P->dest_src_end, P->cap, P->src);
  k_tp_sort_sources<<<tp_grid(nnz), 256, 0, s>>>(dkeys2, nnz, P->tile_dest_off, P->tile_src_off, P->dest_src_end, P->src);
  TP_CK(cudaGetLastError());
  TP_CK(cudaStreamSynchronize(s));
  P->total_dests = nnz;
  P->bytes = (long long)sizeof(uint16_t) * (P->total_src + ntc * md.ng + 3 * nnz) + (long long)sizeof(int) * (nun + ntc + nrows)
             + (long long)(P->nt + 1) * 24;

done:
  cudaFree(mm); cudaFree(rowkey); cudaFree(rowkey2); cudaFree(ckeys); cudaFree(ckeys2); cudaFree(nkeys); cudaFree(nkeys2);
  cudaFree(dkeys); cudaFree(dkeys2); cudaFree(iota); cudaFree(order); cudaFree(rank); cudaFree(pcount); cudaFree(poff);
  cudaFree(dstart); cudaFree(cnt_sorted); cudaFree(totals); cudaFree(nsel); cudaFree(cnt); cudaFree(newidx); cudaFree(tmp);
  if (rc != MPCX_OK) { tile_plan_free(P); P = nullptr; }
  *out = P;
  return rc;
}
}  // namespace
