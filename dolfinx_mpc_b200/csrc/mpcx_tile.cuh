// mpcx_tile.cuh -- "cell tile" assembly of the bulk (constraint-free) cells.
//
// The reference inserts every element matrix with MatSetValuesBlockedLocal (cpp/assemble_matrix.cpp:546), i.e.
// a search + add per entry.  The direct GPU equivalent needs one red.global.add.f64 per entry, and at 16
// entries per P1 tetrahedron that kernel is bound by the ~1.3 cycles/lane issue rate of RED, not by HBM
// (profiles/r01_a, r01_b: 7.3 ms at 256^3 = 20.6 SM-cycles per cell, 15 % of the HBM roofline).  Here the
// entries are first combined inside a tile of neighbouring cells, in shared memory, without atomics:
//
//   setup (once per pattern / dofmap / bc set, mpcx_tile_plan_create, all on the device):
//     the cells are ordered along a Morton curve through the mesh and cut into tiles of C consecutive cells;
//     for every tile the plan holds its distinct vertices, the tile-local vertex ids of its cells, its
//     distinct CSR entries ("dests") and, per dest, the element-matrix entries ("sources") summing into it;
//   assembly (k_ctile_matrix_p1, one CTA of C threads per tile):
//     phase 0  every global read of the tile (vertex coordinates, plan records) as independent coalesced loads
//              into shared memory -- vertex coordinates are fetched once per tile, not once per cell,
//     phase 1  thread = cell: closed-form element matrix -> shared element buffer (conflict-free stores),
//     phase 2  thread = dest: sums its sources from the buffer, ONE red.global.add.f64 per (tile, dest).
//   A P1 tetrahedron mesh has about 4 distinct dests per cell in a 512-cell tile instead of 16 entries per
//   cell, so the RED traffic drops 4x, and the gathers of x[x_dofmap] / row_ptr disappear from the hot loop.
//
// Cells holding slave dofs are excluded (the `skip` flags) and handled by the elimination kernel.
#pragma once
#include <cub/cub.cuh>

namespace
{
#define MPCX_CT_INVALID 0xffffffffu
#define MPCX_TILE_THREADS 512
// element-buffer stride in doubles: odd, so that the 16 entries of one cell fall into 16 different bank pairs
// (phase 2 reads many entries of the same few cells at once)
#define MPCX_TILE_STRIDE (MPCX_TILE_THREADS + 1)


struct TilePlan
{
  int nt = 0, C = 0, ne = 0, ng = 0, nd0 = 0, nd1 = 0;
  int max_nodes = 0, max_dests = 0;
  long long nrows = 0, n_bulk = 0, total_nodes = 0, total_dests = 0, total_src = 0, bytes = 0;
  int *cell_pos = nullptr, *tile_node_off = nullptr, *node_ids = nullptr, *dest_k = nullptr, *tile_ns = nullptr, *tile_nd = nullptr;
  long long* tile_dest_off = nullptr;
  uint16_t *cell_nodes = nullptr, *dest_end = nullptr, *src = nullptr, *cell_rows = nullptr;
  int vec = 0;  // 1: vector plan (dests = row dofs, ne = nd0)
};

struct TilePlanD  // what the kernel sees
{
  int C, max_nodes, max_dests;
  long long n_bulk;
  const int *cell_pos, *tile_node_off, *node_ids, *dest_k, *tile_ns, *tile_nd;
  const long long* tile_dest_off;
  const uint16_t *cell_nodes, *dest_end, *src, *cell_rows;
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

// Morton code of the cell centroid; skipped cells sort last
__global__ void k_tp_cell_codes(MeshD mesh, BBox bb, const int* __restrict__ cells, long long nc,
                                const int8_t* __restrict__ skip, unsigned long long* __restrict__ code,
                                int* __restrict__ iota)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  iota[i] = (int)i;
  if (skip && skip[i]) { code[i] = ~0ull; return; }
  const int cell = cells ? cells[i] : (int)i;
  double c[3] = {0, 0, 0};
  for (int g = 0; g < mesh.ng; ++g)
  {
    const double* p = mesh.x + (long long)mesh.xd[(long long)cell * mesh.ng + g] * mesh.xs;
    for (int k = 0; k < 3; ++k) c[k] += p[k];
  }
  unsigned long long m = 0;
  for (int k = 0; k < 3; ++k)
  {
    double u = (c[k] / mesh.ng - bb.lo[k]) * bb.inv[k];
    u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
    m |= spread21((unsigned long long)(u * 2097151.0)) << k;
  }
  code[i] = m;
}

__global__ void k_tp_count_bulk(const unsigned long long* __restrict__ sorted_code, long long nc, long long* __restrict__ n_bulk)
{
  // first index holding the "skipped" code (codes are sorted ascending)
  if (blockIdx.x || threadIdx.x) return;
  long long lo = 0, hi = nc;
  while (lo < hi)
  {
    const long long mid = (lo + hi) >> 1;
    if (sorted_code[mid] != ~0ull) lo = mid + 1; else hi = mid;
  }
  *n_bulk = lo;
}

// One CTA per tile, thread = cell.  pass 0 counts the distinct vertices / dests / valid sources of the tile,
// pass 1 (after the host scanned the counts) writes the plan records.
template <int NT, int NEc, int NGc>
__global__ void __launch_bounds__(NT)
k_ct_build(int pass, const int* __restrict__ order, long long n_bulk, const int* __restrict__ cells, MeshD mesh,
           const int* __restrict__ dm0, const int* __restrict__ dm1, int nd0, int nd1, int bs0, int bs1,
           const int8_t* __restrict__ bc0, const int8_t* __restrict__ bc1, CsrD A, int* __restrict__ tile_nn,
           int* __restrict__ tile_nd, int* __restrict__ tile_ns, const int* __restrict__ tile_node_off,
           const long long* __restrict__ tile_dest_off, int* __restrict__ cell_pos, int* __restrict__ node_ids,
           uint16_t* __restrict__ cell_nodes, int* __restrict__ dest_k, uint16_t* __restrict__ dest_end,
           uint16_t* __restrict__ src, int extra_off, int vec, uint16_t* __restrict__ cell_rows)
{
  using SortD = cub::BlockRadixSort<unsigned, NT, NEc, unsigned short>;
  using SortN = cub::BlockRadixSort<unsigned, NT, NGc, unsigned short>;
  using DiscD = cub::BlockDiscontinuity<unsigned, NT>;
  using Scan = cub::BlockScan<int, NT>;
  extern __shared__ __align__(16) unsigned char ct_smem[];
  auto& sortd = *reinterpret_cast<typename SortD::TempStorage*>(ct_smem);
  auto& sortn = *reinterpret_cast<typename SortN::TempStorage*>(ct_smem);
  auto& disc = *reinterpret_cast<typename DiscD::TempStorage*>(ct_smem);
  auto& scan = *reinterpret_cast<typename Scan::TempStorage*>(ct_smem);
  const int t = blockIdx.x, cl = threadIdx.x;
  const long long first = (long long)t * NT;
  const int nc_t = (int)((n_bulk - first) < NT ? (n_bulk - first) : NT);
  const bool active = cl < nc_t;
  const int pos = active ? order[first + cl] : -1;
  const int cell = active ? (cells ? cells[pos] : pos) : 0;
  if (pass == 1) cell_pos[first + cl] = pos;

  // ---- vertices
  {
    unsigned keys[NGc];
    unsigned short vals[NGc];
#pragma unroll
    for (int g = 0; g < NGc; ++g)
    {
      keys[g] = active ? (unsigned)mesh.xd[(long long)cell * NGc + g] : MPCX_CT_INVALID;
      vals[g] = (unsigned short)(cl * NGc + g);
    }
    SortN(sortn).Sort(keys, vals);
    __syncthreads();
    int head[NGc];
    DiscD(disc).FlagHeads(head, keys, cub::Inequality());
    __syncthreads();
    int h = 0;
#pragma unroll
    for (int g = 0; g < NGc; ++g)
    {
      head[g] = head[g] && keys[g] != MPCX_CT_INVALID;
      h += head[g];
    }
    int hoff, total;
    Scan(scan).ExclusiveSum(h, hoff, total);
    __syncthreads();
    if (pass == 0)
    {
      if (cl == 0) tile_nn[t] = total;
    }
    else
    {
      const int noff = tile_node_off[t];
      int li = hoff - 1;  // run index of the item before this thread's first item
#pragma unroll
      for (int g = 0; g < NGc; ++g)
      {
        if (keys[g] == MPCX_CT_INVALID) continue;
        if (head[g])
        {
          ++li;
          node_ids[noff + li] = (int)keys[g];
        }
        cell_nodes[first * NGc + vals[g]] = (uint16_t)li;
      }
    }
  }
  __syncthreads();

  // ---- dests and sources
  {
    unsigned keys[NEc];
    unsigned short vals[NEc];
    const int n1 = nd1 * bs1;
#pragma unroll
    for (int e = 0; e < NEc; ++e)
    {
      keys[e] = MPCX_CT_INVALID;
      vals[e] = (unsigned short)(e * MPCX_TILE_STRIDE + cl);
      if (active && vec)
        keys[e] = (unsigned)dm0[(long long)cell * nd0 + e];  // vector plan (bs == 1): dest = row dof of local entry e
      else if (active)
      {
        const int p = e / n1, q = e - p * n1;
        const int r = dm0[(long long)cell * nd0 + p / bs0] * bs0 + p % bs0;
        const int c = dm1[(long long)cell * nd1 + q / bs1] * bs1 + q % bs1;
        // bc rows / columns are zeroed before insertion (cpp/assemble_matrix.cpp:513-533): no source at all
        if (!((bc0 && bc0[r]) || (bc1 && bc1[c])))
        {
          const long long k = csr_find(A, r, c);
          if (k < 0) g_dev_err = MPCX_ERR_PATTERN; else keys[e] = (unsigned)k;
        }
      }
    }
    SortD(sortd).Sort(keys, vals);
    __syncthreads();
    int head[NEc];
    DiscD(disc).FlagHeads(head, keys, cub::Inequality());
    __syncthreads();
    int h = 0, nv = 0;
#pragma unroll
    for (int e = 0; e < NEc; ++e)
    {
      const bool valid = keys[e] != MPCX_CT_INVALID;
      head[e] = head[e] && valid;
      h += head[e];
      nv += valid;
    }
    int hoff, total, voff, vtotal;
    Scan(scan).ExclusiveSum(h, hoff, total);
    __syncthreads();
    Scan(scan).ExclusiveSum(nv, voff, vtotal);
    __syncthreads();
    if (pass == 0)
    {
      if (cl == 0) { tile_nd[t] = total; tile_ns[t] = vtotal; }
    }
    else
    {
      // Dests are emitted in order of DESCENDING source count (stable), so that the lanes of a warp of the
      // assembly kernel's reduction phase run the same number of iterations (diagonal entries collect ~24
      // element entries, off-diagonals 4-6: in CSR order a warp would idle for most of the longest lane's loop).
      unsigned char* xp = ct_smem + extra_off;
      unsigned* dkey = reinterpret_cast<unsigned*>(xp);                       // [NT*NEc]   CSR entry of dest d
      unsigned short* dstart = reinterpret_cast<unsigned short*>(dkey + NT * NEc);  // [NT*NEc+1] first source rank of d
      unsigned short* nst = dstart + NT * NEc + 8;                            // [NT*NEc]   first source rank after reordering
      unsigned short* npos = nst + NT * NEc;                                  // [NT*NEc]   position of d after reordering
      const long long doff = tile_dest_off[t];
      uint16_t* s = src + first * NEc;
      {
        int di = hoff;
#pragma unroll
        for (int e = 0; e < NEc; ++e)
          if (keys[e] != MPCX_CT_INVALID && head[e])
          {
            dkey[di] = keys[e];
            dstart[di] = (unsigned short)(cl * NEc + e);  // valid keys sort first: rank == position among the valid sources
            ++di;
          }
        if (cl == 0) dstart[total] = (unsigned short)vtotal;
      }
      __syncthreads();
      unsigned ckey[NEc];
      unsigned short cval[NEc];
#pragma unroll
      for (int e = 0; e < NEc; ++e)
      {
        const int d = cl * NEc + e;
        const int c = d < total ? (int)dstart[d + 1] - (int)dstart[d] : 0;
        ckey[e] = d < total ? (unsigned)(63 - (c < 63 ? c : 63)) : 64u;
        cval[e] = (unsigned short)d;
      }
      SortD(sortd).Sort(ckey, cval, 0, 7);  // LSD radix sort: stable, so the plan is deterministic
      __syncthreads();
      int cnt[NEc], csum = 0;
#pragma unroll
      for (int e = 0; e < NEc; ++e)
      {
        const int d = cval[e];
        cnt[e] = ckey[e] != 64u ? (int)dstart[d + 1] - (int)dstart[d] : 0;
        csum += cnt[e];
      }
      int coff, ctotal;
      Scan(scan).ExclusiveSum(csum, coff, ctotal);
      __syncthreads();
#pragma unroll
      for (int e = 0; e < NEc; ++e)
      {
        if (ckey[e] == 64u) continue;
        const int d = cval[e], p = cl * NEc + e;
        nst[d] = (unsigned short)coff;
        npos[d] = (unsigned short)p;
        coff += cnt[e];
        dest_k[doff + p] = (int)dkey[d];
        dest_end[doff + p] = (uint16_t)coff;
      }
      __syncthreads();
      {
        int d = hoff - 1;  // items before the thread's first head continue the previous thread's dest
#pragma unroll
        for (int e = 0; e < NEc; ++e)
        {
          if (keys[e] == MPCX_CT_INVALID) continue;
          if (head[e]) ++d;
          const int rank = cl * NEc + e;
          s[(int)nst[d] + (rank - (int)dstart[d])] = vals[e];
          if (vec)  // tile-local row of (cell, local entry): lets the kernel stage per-row data once per tile
            cell_rows[first * NEc + (vals[e] % MPCX_TILE_STRIDE) * NEc + vals[e] / MPCX_TILE_STRIDE] = npos[d];
        }
      }
    }
  }
}

// ------------------------------------------------------------------ the assembly kernel
// Shared-memory layout of one tile (sections 16-byte aligned):
//   Xs[max_nodes][3] f64 | ebuf[NE][C] f64 | dk[max_dests] i32 | ssrc[NE*C] u16 | cnode[C][NV] u16 | dend[max_dests] u16
__host__ __device__ inline size_t tile_smem_bytes(int ne, int nv, int C, int max_nodes, int max_dests)
{
  auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
  return al(24 * (size_t)max_nodes) + al(8 * (size_t)ne * (C + 1)) + al(4 * (size_t)max_dests) + al(2 * (size_t)ne * C)
         + al(2 * (size_t)nv * C) + al(2 * (size_t)max_dests) + 16;
}

// ---- 1-D TMA (cp.async.bulk global -> shared, completion on an mbarrier) for the contiguous plan records
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

template <int TD, bool TMA>
__global__ void __launch_bounds__(MPCX_TILE_THREADS)
k_ctile_matrix_p1(TilePlanD P, IntD in, MeshD mesh, CsrD A)
{
  constexpr int NV = TD + 1, NE = NV * NV, NT = MPCX_TILE_THREADS;
  extern __shared__ __align__(16) unsigned char tile_smem[];
  auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
  unsigned char* sp = tile_smem;
  double* Xs = reinterpret_cast<double*>(sp); sp += al(24 * (size_t)P.max_nodes);
  double* ebuf = reinterpret_cast<double*>(sp); sp += al(8 * (size_t)NE * MPCX_TILE_STRIDE);
  int* dk = reinterpret_cast<int*>(sp); sp += al(4 * (size_t)P.max_dests);
  uint16_t* ssrc = reinterpret_cast<uint16_t*>(sp); sp += al(2 * (size_t)NE * NT);
  uint16_t* cnode = reinterpret_cast<uint16_t*>(sp); sp += al(2 * (size_t)NV * NT);
  uint16_t* dend = reinterpret_cast<uint16_t*>(sp); sp += al(2 * (size_t)P.max_dests);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sp);
  const int t = blockIdx.x, tid = threadIdx.x;
  const long long first = (long long)t * NT;
  const int nc_t = (int)((P.n_bulk - first) < NT ? (P.n_bulk - first) : NT);
  const int n0 = P.tile_node_off[t], nn_t = P.tile_node_off[t + 1] - n0;
  const long long d0 = P.tile_dest_off[t];  // multiple of 8: 16-byte aligned dest records
  const int nd_t = P.tile_nd[t];
  const int ns_t = P.tile_ns[t];

  // phase 0: every global read of the tile.  The contiguous plan records travel by TMA bulk copies issued by
  // one thread; meanwhile all threads gather the vertex coordinates (the only indirect read).
  if (TMA)
  {
    if (tid == 0)
    {
      mbar_init(bar, 1);
      const unsigned b_src = (unsigned)(((ns_t + 7) / 8) * 16), b_cn = (unsigned)(2 * NV * NT);
      const unsigned nd8 = (unsigned)((nd_t + 7) & ~7);
      mbar_expect_tx(bar, b_src + b_cn + nd8 * 6);
      if (b_src) tma_load_1d(ssrc, P.src + first * NE, b_src, bar);
      tma_load_1d(cnode, P.cell_nodes + first * NV, b_cn, bar);
      if (nd8)
      {
        tma_load_1d(dk, P.dest_k + d0, nd8 * 4, bar);
        tma_load_1d(dend, P.dest_end + d0, nd8 * 2, bar);
      }
    }
  }
  for (int i = tid; i < nn_t; i += NT)
  {
    const double* p = mesh.x + (long long)__ldg(P.node_ids + n0 + i) * mesh.xs;
    if (mesh.xs == 4)
    {
      const double2 a = __ldg(reinterpret_cast<const double2*>(p));
      Xs[3 * i] = a.x; Xs[3 * i + 1] = a.y;
    }
    else
    {
      Xs[3 * i] = __ldg(p);
      Xs[3 * i + 1] = __ldg(p + 1);
    }
    Xs[3 * i + 2] = __ldg(p + 2);
  }
  if (!TMA)
  {
    {
      const uint4* g = reinterpret_cast<const uint4*>(P.src + first * NE);
      uint4* s = reinterpret_cast<uint4*>(ssrc);
      for (int i = tid; i < (ns_t + 7) / 8; i += NT) s[i] = __ldg(g + i);
    }
    for (int k = tid; k < nd_t; k += NT)
    {
      dk[k] = __ldg(P.dest_k + d0 + k);
      dend[k] = __ldg(P.dest_end + d0 + k);
    }
    if (NV == 4)
      reinterpret_cast<uint2*>(cnode)[tid] = __ldg(reinterpret_cast<const uint2*>(P.cell_nodes) + first + tid);
    else
      for (int i = tid; i < NT * NV; i += NT) cnode[i] = __ldg(P.cell_nodes + first * NV + i);
  }
  __syncthreads();  // Xs complete; the mbarrier initialisation is visible to every thread
  if (TMA) mbar_wait(bar, 0);

  // phase 1: thread = cell; element matrix -> element buffer, entry-major (conflict-free stores)
  if (tid < nc_t)
  {
    double X[NV][3];
#pragma unroll
    for (int v = 0; v < NV; ++v)
    {
      const int l = cnode[tid * NV + v];
      X[v][0] = Xs[3 * l];
      X[v][1] = Xs[3 * l + 1];
      X[v][2] = TD == 3 ? Xs[3 * l + 2] : 0.0;
    }
    P1Geom<TD> G;
    p1_geometry<TD>(X, G);
    double w[NV], Ae[NV][NV];
    if (in.kernel == MPCX_KERNEL_LAPLACE_VARCOEF)
    {
      const long long index = __ldg(P.cell_pos + first + tid);
      p1_load_w<TD>(in, index, in.cells ? __ldg(in.cells + index) : (int)index, w);
    }
    p1_element<TD>(in.kernel, G, in.c, w, Ae);
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
      for (int j = 0; j < NV; ++j) ebuf[(i * NV + j) * MPCX_TILE_STRIDE + tid] = Ae[i][j];
  }
  __syncthreads();

  // phase 2: thread = dest; sum of its sources (4 independent chains), one RED per (tile, dest)
  for (int k = tid; k < nd_t; k += NT)
  {
    const int beg = k ? dend[k - 1] : 0, end = dend[k];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int p = beg;
    for (; p + 4 <= end; p += 4)
    {
      const int a = ssrc[p], b = ssrc[p + 1], c = ssrc[p + 2], d = ssrc[p + 3];
      s0 += ebuf[a]; s1 += ebuf[b]; s2 += ebuf[c]; s3 += ebuf[d];
    }
    for (; p < end; ++p) s0 += ebuf[ssrc[p]];
    atomicAdd(A.val + dk[k], (s0 + s1) + (s2 + s3));
  }
}

// ------------------------------------------------------------------ vector tile kernel (P1 source term)
// Same three phases for the load vector b_i = c0 |K|/((d+1)(d+2)) (f_i + sum_j f_j)
// (cpp/assemble_vector.cpp:163-185 with the P1 source kernel): 4 element entries per cell, dests = the row dofs
// of the tile, one red.global.add.f64 per (tile, row) instead of one per (cell, vertex).  When the coefficient
// lives in the same space as the test function its values are staged once per tile row.
//   Xs[max_nodes][3] f64 | ebuf[NV][C+1] f64 | fs[max_dests] f64 | dk[max_dests] i32 | ssrc[NV*C] u16 |
//   cnode[C][NV] u16 | crow[C][NV] u16 | dend[max_dests] u16 | mbarrier
__host__ __device__ inline size_t vtile_smem_bytes(int nv, int C, int max_nodes, int max_dests)
{
  auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
  return al(24 * (size_t)max_nodes) + al(8 * (size_t)nv * (C + 1)) + al(8 * (size_t)max_dests) + al(4 * (size_t)max_dests)
         + 3 * al(2 * (size_t)nv * C) + al(2 * (size_t)max_dests) + 16;
}

template <int TD>
__global__ void __launch_bounds__(MPCX_TILE_THREADS)
k_ctile_vector_p1(TilePlanD P, IntD in, MeshD mesh, int w_by_row, double* __restrict__ b)
{
  constexpr int NV = TD + 1, NT = MPCX_TILE_THREADS;
  extern __shared__ __align__(16) unsigned char tile_smem[];
  auto al = [](size_t b_) { return (b_ + 15) & ~(size_t)15; };
  unsigned char* sp = tile_smem;
  double* Xs = reinterpret_cast<double*>(sp); sp += al(24 * (size_t)P.max_nodes);
  double* ebuf = reinterpret_cast<double*>(sp); sp += al(8 * (size_t)NV * MPCX_TILE_STRIDE);
  double* fs = reinterpret_cast<double*>(sp); sp += al(8 * (size_t)P.max_dests);
  int* dk = reinterpret_cast<int*>(sp); sp += al(4 * (size_t)P.max_dests);
  uint16_t* ssrc = reinterpret_cast<uint16_t*>(sp); sp += al(2 * (size_t)NV * NT);
  uint16_t* cnode = reinterpret_cast<uint16_t*>(sp); sp += al(2 * (size_t)NV * NT);
  uint16_t* crow = reinterpret_cast<uint16_t*>(sp); sp += al(2 * (size_t)NV * NT);
  uint16_t* dend = reinterpret_cast<uint16_t*>(sp); sp += al(2 * (size_t)P.max_dests);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(sp);
  const int t = blockIdx.x, tid = threadIdx.x;
  const long long first = (long long)t * NT;
  const int nc_t = (int)((P.n_bulk - first) < NT ? (P.n_bulk - first) : NT);
  const int n0 = P.tile_node_off[t], nn_t = P.tile_node_off[t + 1] - n0;
  const long long d0 = P.tile_dest_off[t];
  const int nd_t = P.tile_nd[t];
  const int ns_t = P.tile_ns[t];

  if (tid == 0)
  {
    mbar_init(bar, 1);
    const unsigned b_src = (unsigned)(((ns_t + 7) / 8) * 16), b_cn = (unsigned)(2 * NV * NT);
    const unsigned nd8 = (unsigned)((nd_t + 7) & ~7);
    mbar_expect_tx(bar, b_src + 2 * b_cn + nd8 * 6);
    if (b_src) tma_load_1d(ssrc, P.src + first * NV, b_src, bar);
    tma_load_1d(cnode, P.cell_nodes + first * NV, b_cn, bar);
    tma_load_1d(crow, P.cell_rows + first * NV, b_cn, bar);
    if (nd8)
    {
      tma_load_1d(dk, P.dest_k + d0, nd8 * 4, bar);
      tma_load_1d(dend, P.dest_end + d0, nd8 * 2, bar);
    }
  }
  for (int i = tid; i < nn_t; i += NT)
  {
    const double* p = mesh.x + (long long)__ldg(P.node_ids + n0 + i) * mesh.xs;
    if (mesh.xs == 4)
    {
      const double2 a = __ldg(reinterpret_cast<const double2*>(p));
      Xs[3 * i] = a.x; Xs[3 * i + 1] = a.y;
    }
    else
    {
      Xs[3 * i] = __ldg(p);
      Xs[3 * i + 1] = __ldg(p + 1);
    }
    Xs[3 * i + 2] = __ldg(p + 2);
  }
  if (w_by_row)  // coefficient in the test space: one read per tile row (straight from the plan, no wait on the TMA)
    for (int k = tid; k < nd_t; k += NT) fs[k] = __ldg(in.wnodal + __ldg(P.dest_k + d0 + k));
  __syncthreads();
  mbar_wait(bar, 0);

  // phase 1: thread = cell
  if (tid < nc_t)
  {
    double X[NV][3];
#pragma unroll
    for (int v = 0; v < NV; ++v)
    {
      const int l = cnode[tid * NV + v];
      X[v][0] = Xs[3 * l];
      X[v][1] = Xs[3 * l + 1];
      X[v][2] = TD == 3 ? Xs[3 * l + 2] : 0.0;
    }
    P1Geom<TD> G;
    p1_geometry<TD>(X, G);
    double f[NV], fsum = 0.0;
    if (w_by_row)
    {
#pragma unroll
      for (int v = 0; v < NV; ++v) f[v] = fs[crow[tid * NV + v]];
    }
    else
    {
      const long long index = __ldg(P.cell_pos + first + tid);
      p1_load_w<TD>(in, index, in.cells ? __ldg(in.cells + index) : (int)index, f);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) fsum += f[v];
    const double s = in.c[0] * G.vol / double((TD + 1) * (TD + 2));
#pragma unroll
    for (int v = 0; v < NV; ++v) ebuf[v * MPCX_TILE_STRIDE + tid] = s * (f[v] + fsum);
  }
  __syncthreads();

  // phase 2: thread = row of the tile
  for (int k = tid; k < nd_t; k += NT)
  {
    const int beg = k ? dend[k - 1] : 0, end = dend[k];
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int p = beg;
    for (; p + 4 <= end; p += 4)
    {
      const int a = ssrc[p], bb = ssrc[p + 1], c = ssrc[p + 2], d = ssrc[p + 3];
      s0 += ebuf[a]; s1 += ebuf[bb]; s2 += ebuf[c]; s3 += ebuf[d];
    }
    for (; p < end; ++p) s0 += ebuf[ssrc[p]];
    atomicAdd(b + dk[k], (s0 + s1) + (s2 + s3));
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
  cudaFree(P->cell_pos); cudaFree(P->tile_node_off); cudaFree(P->node_ids); cudaFree(P->dest_k); cudaFree(P->tile_ns); cudaFree(P->tile_nd);
  cudaFree(P->tile_dest_off); cudaFree(P->cell_nodes); cudaFree(P->dest_end); cudaFree(P->src); cudaFree(P->cell_rows);
  delete P;
}

inline unsigned tp_grid(long long n, int b = 256) { return (unsigned)((n + b - 1) / b > 0 ? (n + b - 1) / b : 1); }

template <int NEc, int NGc>
cudaError_t ct_build_launch(int pass, int nt, cudaStream_t s, const int* order, long long n_bulk, const int* cells, MeshD md,
                            const mpcx_dofmap* dm0, const mpcx_dofmap* dm1, const int8_t* bc0, const int8_t* bc1, CsrD A,
                            int* tile_nn, int* tile_nd, TilePlan* P)
{
  constexpr int NT = MPCX_TILE_THREADS;
  auto kern = k_ct_build<NT, NEc, NGc>;
  size_t smem = sizeof(typename cub::BlockRadixSort<unsigned, NT, NEc, unsigned short>::TempStorage);
  smem = std::max(smem, sizeof(typename cub::BlockRadixSort<unsigned, NT, NGc, unsigned short>::TempStorage));
  smem = std::max(smem, sizeof(typename cub::BlockDiscontinuity<unsigned, NT>::TempStorage));
  smem = (std::max(smem, sizeof(typename cub::BlockScan<int, NT>::TempStorage)) + 15) & ~(size_t)15;
  const int extra_off = (int)smem;
  smem += (size_t)NT * NEc * 4 + 3 * ((size_t)NT * NEc + 8) * 2 + 16;  // dkey, dstart, nst, npos of the count ordering
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<nt, NT, smem, s>>>(pass, order, n_bulk, cells, md, dm0->map, dm1->map, dm0->nd, dm1->nd, dm0->bs, dm1->bs, bc0, bc1, A,
                            tile_nn, tile_nd, P->tile_ns, P->tile_node_off, P->tile_dest_off, P->cell_pos, P->node_ids,
                            P->cell_nodes, P->dest_k, P->dest_end, P->src, extra_off, P->vec, P->cell_rows);
  return cudaGetLastError();
}

// Acsr == nullptr builds a VECTOR plan: dests are the row dofs of dm0 (bs == 1), ne = nd0 entries per cell.
int tile_plan_build(const mpcx_mesh* mesh, const mpcx_dofmap* dm0, const mpcx_dofmap* dm1, const int32_t* cells,
                    long long nc, const int8_t* skip, const int8_t* bc0, const int8_t* bc1, const mpcx_csr* Acsr,
                    cudaStream_t s, TilePlan** out)
{
  int rc = MPCX_OK;
  TilePlan* P = new TilePlan();
  const MeshD md{mesh->x, mesh->x_dofmap, mesh->ng, mesh->x_stride};
  const bool vec = Acsr == nullptr;
  const CsrD A{vec ? nullptr : (const long long*)Acsr->row_ptr, vec ? nullptr : Acsr->col, vec ? nullptr : Acsr->val};
  constexpr int C = MPCX_TILE_THREADS;
  P->vec = vec ? 1 : 0;
  P->C = C; P->nd0 = dm0->nd; P->nd1 = dm1->nd; P->ng = mesh->ng; P->nrows = vec ? dm0->num_dofs : Acsr->num_rows;
  P->ne = vec ? dm0->nd : dm0->nd * dm0->bs * dm1->nd * dm1->bs;
  unsigned long long *mm = nullptr, *code = nullptr, *code2 = nullptr;
  int *iota = nullptr, *order = nullptr, *tile_nn = nullptr;
  long long* nb_dev = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0, tb = 0;
  std::vector<int> h_nn, h_nd, noff;
  std::vector<long long> h_doff;
  long long n_dests = 0, alloc_dests = 0;
  BBox bb;
  int variant = 0;
  auto need_tmp = [&](size_t bytes) -> cudaError_t {
    if (bytes <= tmp_bytes) return cudaSuccess;
    cudaFree(tmp);
    tmp = nullptr;
    tmp_bytes = bytes + bytes / 8;
    return cudaMalloc(&tmp, tmp_bytes);
  };
  auto build = [&](int pass) -> cudaError_t {
    if (variant == 4) return ct_build_launch<4, 4>(pass, P->nt, s, order, P->n_bulk, cells, md, dm0, dm1, bc0, bc1, A, tile_nn, P->tile_nd, P);
    if (variant == 3) return ct_build_launch<3, 3>(pass, P->nt, s, order, P->n_bulk, cells, md, dm0, dm1, bc0, bc1, A, tile_nn, P->tile_nd, P);
    if (variant == 16) return ct_build_launch<16, 4>(pass, P->nt, s, order, P->n_bulk, cells, md, dm0, dm1, bc0, bc1, A, tile_nn, P->tile_nd, P);
    return ct_build_launch<9, 3>(pass, P->nt, s, order, P->n_bulk, cells, md, dm0, dm1, bc0, bc1, A, tile_nn, P->tile_nd, P);
  };

  if (vec && dm0->bs == 1 && P->ne == 4 && mesh->ng == 4) variant = 4;
  else if (vec && dm0->bs == 1 && P->ne == 3 && mesh->ng == 3) variant = 3;
  else if (!vec && P->ne == 16 && mesh->ng == 4) variant = 16;
  else if (!vec && P->ne == 9 && mesh->ng == 3) variant = 9;
  else { rc = fail(MPCX_ERR_UNSUPPORTED, "tile plan: only scalar P1 simplices (3x3 / 4x4 element matrices, 3 / 4 element vectors) so far"); goto done; }
  if ((!vec && Acsr->nnz >= (1ll << 31) - 1) || nc >= (1ll << 31)) { rc = fail(MPCX_ERR_UNSUPPORTED, "tile plan: nnz or cells >= 2^31 on one device"); goto done; }

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
  // 2. cells along the Morton curve (skipped cells last), tiles of C consecutive cells
  TP_CK(tp_alloc(&code, nc)); TP_CK(tp_alloc(&code2, nc)); TP_CK(tp_alloc(&iota, nc)); TP_CK(tp_alloc(&order, nc));
  TP_CK(tp_alloc(&nb_dev, 1));
  k_tp_cell_codes<<<tp_grid(nc), 256, 0, s>>>(md, bb, cells, nc, skip, code, iota);
  TP_CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, code, code2, iota, order, (int)nc, 0, 64, s));
  TP_CK(need_tmp(tb));
  TP_CK(cub::DeviceRadixSort::SortPairs(tmp, tb, code, code2, iota, order, (int)nc, 0, 64, s));
  k_tp_count_bulk<<<1, 32, 0, s>>>(code2, nc, nb_dev);
  TP_CK(cudaMemcpyAsync(&P->n_bulk, nb_dev, sizeof(long long), cudaMemcpyDeviceToHost, s));
  TP_CK(cudaStreamSynchronize(s));
  cudaFree(code); code = nullptr; cudaFree(code2); code2 = nullptr; cudaFree(iota); iota = nullptr;
  P->nt = (int)((P->n_bulk + C - 1) / C);
  if (P->nt == 0) goto done;  // nothing but slave cells: an empty plan is valid

  // 3. pass 0: sizes of every tile
  TP_CK(tp_alloc(&tile_nn, P->nt)); TP_CK(tp_alloc(&P->tile_nd, P->nt)); TP_CK(tp_alloc(&P->tile_ns, P->nt));
  TP_CK(build(0));
  h_nn.resize(P->nt); h_nd.resize(P->nt);
  TP_CK(cudaMemcpyAsync(h_nn.data(), tile_nn, sizeof(int) * P->nt, cudaMemcpyDeviceToHost, s));
  TP_CK(cudaMemcpyAsync(h_nd.data(), P->tile_nd, sizeof(int) * P->nt, cudaMemcpyDeviceToHost, s));
  TP_CK(cudaStreamSynchronize(s));
  noff.assign(P->nt + 1, 0);
  h_doff.assign(P->nt + 1, 0);
  for (int t = 0; t < P->nt; ++t)
  {
    noff[t + 1] = noff[t] + h_nn[t];
    h_doff[t + 1] = h_doff[t] + ((h_nd[t] + 7) & ~7);  // 16-byte aligned records for the TMA bulk copies
    n_dests += h_nd[t];
    P->max_nodes = std::max(P->max_nodes, h_nn[t]);
    P->max_dests = std::max(P->max_dests, h_nd[t]);
  }
  P->total_nodes = noff[P->nt];
  P->total_dests = n_dests;
  alloc_dests = h_doff[P->nt] + 8;
  P->total_src = (long long)P->nt * C * P->ne;
  P->max_nodes = (P->max_nodes + 1) & ~1;
  P->max_dests = (P->max_dests + 7) & ~7;  // the TMA copies move whole groups of 8 dest records
  TP_CK(tp_alloc(&P->tile_node_off, P->nt + 1)); TP_CK(tp_alloc(&P->tile_dest_off, P->nt + 1));
  TP_CK(cudaMemcpyAsync(P->tile_node_off, noff.data(), sizeof(int) * (P->nt + 1), cudaMemcpyHostToDevice, s));
  TP_CK(cudaMemcpyAsync(P->tile_dest_off, h_doff.data(), sizeof(long long) * (P->nt + 1), cudaMemcpyHostToDevice, s));
  TP_CK(cudaStreamSynchronize(s));
  // 4. pass 1: the records
  TP_CK(tp_alloc(&P->cell_pos, (long long)P->nt * C)); TP_CK(tp_alloc(&P->node_ids, P->total_nodes));
  TP_CK(tp_alloc(&P->cell_nodes, (long long)P->nt * C * P->ng)); TP_CK(tp_alloc(&P->dest_k, alloc_dests));
  TP_CK(tp_alloc(&P->dest_end, alloc_dests)); TP_CK(tp_alloc(&P->src, P->total_src + 8));
  TP_CK(cudaMemsetAsync(P->cell_nodes, 0, sizeof(uint16_t) * (size_t)P->nt * C * P->ng, s));
  TP_CK(cudaMemsetAsync(P->dest_k, 0, sizeof(int) * (size_t)alloc_dests, s));
  TP_CK(cudaMemsetAsync(P->dest_end, 0, sizeof(uint16_t) * (size_t)alloc_dests, s));
  if (vec)
  {
    TP_CK(tp_alloc(&P->cell_rows, (long long)P->nt * C * P->ne));
    TP_CK(cudaMemsetAsync(P->cell_rows, 0, sizeof(uint16_t) * (size_t)P->nt * C * P->ne, s));
  }
  TP_CK(build(1));
  TP_CK(cudaStreamSynchronize(s));
  P->bytes = (long long)sizeof(uint16_t) * (P->total_src + (long long)P->nt * C * P->ng + P->total_dests)
             + (long long)sizeof(int) * (P->total_nodes + P->total_dests + (long long)P->nt * C) + (long long)(P->nt + 1) * 16
             + (vec ? (long long)sizeof(uint16_t) * P->nt * C * P->ne : 0);

done:
  cudaFree(mm); cudaFree(code); cudaFree(code2); cudaFree(iota); cudaFree(order); cudaFree(tile_nn);
  cudaFree(nb_dev); cudaFree(tmp);
  if (rc != MPCX_OK) { tile_plan_free(P); P = nullptr; }
  *out = P;
  return rc;
}
}  // namespace
