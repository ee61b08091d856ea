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
//     for every tile the plan holds its distinct vertices, the tile-local vertex ids of its cells and one record
//     per distinct CSR entry ("dest") the tile contributes to, ordered by descending number of contributing element
//     entries ("sources") and cut into groups of 32 (one per lane of a warp).  Every source owns one slot of the
//     tile's element buffer: slot = base(group) + 33 * i + lane for the i-th source of the dest held by `lane`.
//     In CSR order the dests are cut into runs of (nearly) consecutive entries, gaps and 16-byte padding being
//     zero-filled positions of the tile's staging buffer.  A symmetric plan (same dofmap and bc markers on both
//     sides; every tile kernel's element matrix is symmetric) keeps records for the upper triangle only, each
//     feeding entry (r, c) and entry (c, r);
//   assembly (k_ptile_matrix_p1): persistent CTAs walk the tiles; per tile
//     - the cell records (vertex ids, slots) and dest records arrive by TMA bulk copies (cp.async.bulk + mbarrier,
//       SASS UBLKCP) issued one tile ahead; the tile's vertex coordinates are gathered once per tile -- not once
//       per cell -- through registers, also one tile ahead,
//     - phase 1, thread = cell: closed-form element matrix, each stored entry written to its slot,
//     - phase 2, thread = record: sums its column of the group's [count][33] slot block -- conflict-free reads, same
//       trip count on every lane -- and stores the sum at the record's staging position(s),
//     - one TMA bulk reduction per run (cp.reduce.async.bulk.global.shared::cta.add.f64, SASS UBLKRED): the copy
//       engine adds the run into A.val in L2 -- about 35 bulk operations per tile instead of ~2100
//       red.global.add.f64 lane operations through the LSU (1.3 cycles each, profiles/r01_g: the REDs were half of
//       the L1/LSU time that bounded the kernel); the CTA moves on to the next tile meanwhile.
//   A P1 tetrahedron mesh has about 4 distinct dests per cell in a 448-cell tile instead of 16 entries per
//   cell, and the gathers of x[x_dofmap] / row_ptr disappear from the hot loop.
//   Requirements of the bulk reductions: A.val 16-byte aligned, capacity rounded up to an even number of entries
//   (a run may be padded with one zero past the last entry).
//
// Cells holding slave dofs are excluded (the `skip` flags) and handled by the elimination kernel.
// The load vector uses the same machinery with 4 entries per cell and dests = row dofs (k_ptile_vector_p1); a tile
// touches few, scattered rows, so those sums are added with one red.global.add.f64 per (tile, row).
#pragma once
#include <cub/cub.cuh>

namespace
{
#define MPCX_CT_INVALID 0xffffffffu
#define MPCX_CT_NOSLOT 0xffffu
// cells per tile = threads per CTA of the tile kernels (448: two CTAs with double-buffered records fit one SM)
#ifndef MPCX_TILE_THREADS
#define MPCX_TILE_THREADS 448
#endif
// slot of the i-th source of the dest held by lane l of a group: base + 33 i + l.  The odd stride spreads the
// sources of one dest (written by neighbouring cells at the same time) over the banks; lanes still read consecutive slots.
#define MPCX_CT_GSTRIDE 33

// dests whose CSR positions (row dofs) differ by at most this much share a run; the gaps are zero-filled.  16 merges
// the partial rows of a tile's boundary vertices with their neighbours along a line of consecutive dofs: ~35 bulk
// reductions per tile instead of ~117 with a gap of 4 (the copy engine takes a few hundred cycles per operation)
#ifndef MPCX_CT_RUNGAP
#define MPCX_CT_RUNGAP 16
#endif
#define MPCX_CT_MAXRUNS 2048
// 1: bank-aware placement of the sources of every matrix record (see k_ct_build)
#ifndef MPCX_CT_BANKOPT
#define MPCX_CT_BANKOPT 1
#endif
// staging positions per tile above which the run gap is halved (see k_ct_build)
#ifndef MPCX_CT_STAGECAP
#define MPCX_CT_STAGECAP 3200
#endif
// ... but never at the price of more than this many runs per tile
#ifndef MPCX_CT_RUNCAP
#define MPCX_CT_RUNCAP 256
#endif

struct TilePlan
{
  int nt = 0, C = 0, ne = 0, ng = 0, nd0 = 0, nd1 = 0;
  int max_nodes = 0, max_dests = 0, max_slots = 0, max_runs = 0, max_stage = 0;
  long long n_iface = 0;  // bulk cells touching a ghost row: they fill the first ceil(n_iface / C) tiles
  long long nrows = 0, nvals = 0, n_bulk = 0, total_nodes = 0, total_dests = 0, total_slots = 0, total_runs = 0, bytes = 0;
  long long total_stage = 0;  // staging positions over all tiles = fp64 adds the copy engine performs per assembly
  int *cell_pos = nullptr, *tile_node_off = nullptr, *node_ids = nullptr, *dest_k = nullptr, *tile_nd = nullptr,
      *tile_slots = nullptr, *tile_nr = nullptr, *tile_stage = nullptr;
  int2* runs = nullptr;
  int4* hdr = nullptr;  // per-tile header, 3 x int4 (see load_hdr)
  unsigned* ginfo = nullptr;
  long long *tile_dest_off = nullptr, *tile_run_off = nullptr;
  uint16_t *cell_nodes = nullptr, *dest_spos = nullptr, *dest_spos2 = nullptr, *cell_slot = nullptr, *cell_rows = nullptr;
  uint8_t* dest_cnt = nullptr;
  uint16_t* slot_cell = nullptr;       // vector plans: tile cell of the source held by every slot (inverse of cell_slot)
  long long* tile_slot_off = nullptr;  // vector plans: first entry of the tile in slot_cell (multiple of 8)
  // scatter plan of the slave cells (mpcx_tile_plan_add_slave_cells): per slave cell the insertions of the elimination
  long long *sp_off = nullptr, *sp_pos = nullptr, sp_cells = 0, sp_total = 0;
  uint8_t* sp_ent = nullptr;
  int *sp_ca = nullptr, *sp_cb = nullptr;
  int vec = 0;  // 1: vector plan (dests = row dofs, ne = nd0)
  int sym = 0;  // 1: symmetric matrix plan (upper-triangular records feed both (r, c) and (c, r))
  int ns = 0;   // slots per cell record: ne, or nd (nd + 1) / 2 for a symmetric plan
};

struct TilePlanD  // what the kernel sees
{
  int C, max_nodes, max_dests, max_slots, max_runs, max_stage;
  long long n_bulk;
  const int *cell_pos, *tile_node_off, *node_ids, *dest_k, *tile_nd, *tile_nr, *tile_stage;
  const int2* runs;
  const int4* hdr;
  const unsigned* ginfo;
  const long long *tile_dest_off, *tile_run_off;
  const uint16_t *cell_nodes, *dest_spos, *dest_spos2, *cell_slot, *cell_rows;
  const uint8_t* dest_cnt;
  const uint16_t* slot_cell;
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
  int hilbert;  // 1: Hilbert curve instead of Morton order (MPCX_TILE_CURVE=hilbert)
};

// Skilling's transform: 20-bit coordinates -> "transposed" Hilbert index (interleave X[0], X[1], X[2], X[0] highest)
__device__ __forceinline__ void hilbert_transpose(unsigned X[3])
{
  const unsigned M = 1u << 19;
  for (unsigned Q = M; Q > 1; Q >>= 1)
  {
    const unsigned Pm = Q - 1;
    for (int i = 0; i < 3; ++i)
    {
      if (X[i] & Q) X[0] ^= Pm;
      else { const unsigned t = (X[0] ^ X[i]) & Pm; X[0] ^= t; X[i] ^= t; }
    }
  }
  X[1] ^= X[0]; X[2] ^= X[1];
  unsigned t = 0;
  for (unsigned Q = M; Q > 1; Q >>= 1)
    if (X[2] & Q) t ^= Q - 1;
  X[0] ^= t; X[1] ^= t; X[2] ^= t;
}

// sum over the (non-skipped) cells of the extents of their bounding boxes, per axis, and their number: the mean cell
// size the Morton lattice is aligned with (tile_plan_build)
__global__ void k_tp_cell_extent(MeshD mesh, const int* __restrict__ cells, long long nc, const int8_t* __restrict__ skip,
                                 double* __restrict__ sum4)
{
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += (long long)gridDim.x * blockDim.x)
  {
    if (skip && skip[i]) continue;
    const int cell = cells ? cells[i] : (int)i;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int g = 0; g < mesh.ng; ++g)
    {
      const double* p = mesh.x + (long long)mesh.xd[(long long)cell * mesh.ng + g] * mesh.xs;
      for (int k = 0; k < 3; ++k) { lo[k] = p[k] < lo[k] ? p[k] : lo[k]; hi[k] = p[k] > hi[k] ? p[k] : hi[k]; }
    }
    for (int k = 0; k < 3; ++k) a[k] += hi[k] - lo[k];
    a[3] += 1.0;
  }
  for (int k = 0; k < 4; ++k)
  {
    for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
    if ((threadIdx.x & 31) == 0 && a[k] != 0.0) atomicAdd(sum4 + k, a[k]);
  }
}

// Morton code of the cell centroid; skipped cells sort last
// cells that touch a ghost row (a block >= first_ghost_block of dm) sort before all others: bit 61 clear / set
#define MPCX_TP_INTERIOR_BIT (1ull << 61)
__global__ void k_tp_cell_codes(MeshD mesh, BBox bb, const int* __restrict__ cells, long long nc,
                                const int8_t* __restrict__ skip, const int* __restrict__ dm, int nd, long long first_ghost_block,
                                unsigned long long* __restrict__ code, int* __restrict__ iota)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  iota[i] = (int)i;
  if (skip && skip[i]) { code[i] = ~0ull; return; }
  const int cell = cells ? cells[i] : (int)i;
  bool iface = false;
  if (first_ghost_block > 0)
    for (int k = 0; k < nd; ++k) iface |= dm[(long long)cell * nd + k] >= first_ghost_block;
  double c[3] = {0, 0, 0};
  for (int g = 0; g < mesh.ng; ++g)
  {
    const double* p = mesh.x + (long long)mesh.xd[(long long)cell * mesh.ng + g] * mesh.xs;
    for (int k = 0; k < 3; ++k) c[k] += p[k];
  }
  unsigned long long m = 0;
  unsigned q[3];
  for (int k = 0; k < 3; ++k)
  {
    double u = (c[k] / mesh.ng - bb.lo[k]) * bb.inv[k];
    u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
    q[k] = (unsigned)(u * 1048575.0);  // 20 bits per axis: bits 60.. stay free
  }
  if (bb.hilbert)
  {
    hilbert_transpose(q);
    m = (spread21(q[0]) << 2) | (spread21(q[1]) << 1) | spread21(q[2]);
  }
  else
    for (int k = 0; k < 3; ++k) m |= spread21(q[k]) << k;
  code[i] = m | (iface ? 0ull : MPCX_TP_INTERIOR_BIT);
}

// first index whose code is >= bound (codes sorted ascending)
__global__ void k_tp_count_below(const unsigned long long* __restrict__ sorted_code, long long nc, unsigned long long bound,
                                 long long* __restrict__ out)
{
  if (blockIdx.x || threadIdx.x) return;
  long long lo = 0, hi = nc;
  while (lo < hi)
  {
    const long long mid = (lo + hi) >> 1;
    if (sorted_code[mid] < bound) lo = mid + 1; else hi = mid;
  }
  *out = lo;
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

// One CTA per tile, thread = cell.  pass 0 sizes the tile (distinct vertices, dest records, runs, element-buffer
// slots), pass 1 (after the host scanned the sizes) writes the plan records.
template <int NT, int NEc, int NGc>
__global__ void __launch_bounds__(NT)
k_ct_build(int pass, const int* __restrict__ order, long long n_bulk, const int* __restrict__ cells, MeshD mesh,
           const int* __restrict__ dm0, const int* __restrict__ dm1, int nd0, int nd1, int bs0, int bs1,
           const int8_t* __restrict__ bc0, const int8_t* __restrict__ bc1, CsrD A, long long nrows, int vec, int sym,
           int extra_off, int* __restrict__ tile_nn, int* __restrict__ tile_nd, int* __restrict__ tile_slots,
           int* __restrict__ tile_nr, int* __restrict__ tile_stage,
           const int* __restrict__ tile_node_off, const long long* __restrict__ tile_dest_off,
           const long long* __restrict__ tile_run_off, int* __restrict__ cell_pos, int* __restrict__ node_ids,
           uint16_t* __restrict__ cell_nodes, int* __restrict__ dest_k, uint8_t* __restrict__ dest_cnt,
           uint16_t* __restrict__ dest_spos, uint16_t* __restrict__ dest_spos2, unsigned* __restrict__ ginfo,
           int2* __restrict__ runs,
           uint16_t* __restrict__ cell_slot, uint16_t* __restrict__ cell_rows,
           const long long* __restrict__ tile_slot_off, uint16_t* __restrict__ slot_cell)
{
  using SortD = cub::BlockRadixSort<unsigned, NT, NEc, unsigned short>;
  using SortN = cub::BlockRadixSort<unsigned, NT, NGc, unsigned short>;
  using DiscD = cub::BlockDiscontinuity<unsigned, NT>;
  using Scan = cub::BlockScan<int, NT>;
  constexpr int N = NT * NEc;  // most sources, dests and staging positions a tile can have
  constexpr int MR = MPCX_CT_MAXRUNS < N ? MPCX_CT_MAXRUNS : N;
  extern __shared__ __align__(16) unsigned char ct_smem[];
  auto& sortd = *reinterpret_cast<typename SortD::TempStorage*>(ct_smem);
  auto& sortn = *reinterpret_cast<typename SortN::TempStorage*>(ct_smem);
  auto& disc = *reinterpret_cast<typename DiscD::TempStorage*>(ct_smem);
  auto& scan = *reinterpret_cast<typename Scan::TempStorage*>(ct_smem);
  unsigned* dkey = reinterpret_cast<unsigned*>(ct_smem + extra_off);  // [N]      key (CSR entry / row dof) of dest d, ascending
  int* gbase = reinterpret_cast<int*>(dkey + N);                      // [N/32+1] first slot of dest group g
  int* rk0 = gbase + N / 32 + 1;                                      // [MR]     first (even) key of run r
  int* rend = rk0 + MR;                                               // [MR]     end (even) key of run r, then its staging offset
  unsigned short* dstart = reinterpret_cast<unsigned short*>(rend + MR);  // [N+8]  first source rank of d
  unsigned short* dj = dstart + N + 8;                                 // [N]     staging position of dest d
  unsigned short* rcnt = dj + N;                                       // [N]     sources of the record of dest d (0: no record)
  unsigned short* npos = rcnt + N;                                     // [N]     position of d's record in count order
  unsigned short* pcnt = npos + N;                                     // [N+32]  source count of record p
  unsigned short* part = pcnt + N + 32;                                // [N]     dest whose record collects d's sources
  unsigned short* sp2 = part + N;                                      // [N]     staging position of the transposed entry
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
    // Tile-local vertex NUMBERS are free.  Phase 1 of the kernels gathers the coordinates of vertex v of 16 cells with
    // one LDS.64 per half-warp and component: two different vertices whose numbers agree mod 16 cost an extra
    // wavefront (12 such loads per cell: 41 M of the fused kernel's 238 M excess wavefronts at 256^3, profiles/r02_c).
    // Greedy colouring, vertex by vertex in id order: the residue (mod 16) no other vertex of the same (half-warp,
    // v) groups holds yet, classes filled evenly; the tile's vertex list is padded to a multiple of 16 (holes repeat the
    // first vertex).
    const int nnp = (total + 15) & ~15;
    if (pass == 0)
    {
      if (cl == 0) tile_nn[t] = nnp;
    }
    else
    {
      constexpr int NI = NT * NGc, NGRP = (NT / 16) * NGc;
      unsigned char* vgrp = ct_smem + extra_off;                                              // [NI] group of sorted item s
      unsigned short* vstart = reinterpret_cast<unsigned short*>(vgrp + ((NI + 15) & ~15));   // [NI + 8] first item of vertex li
      unsigned short* vperm = vstart + NI + 8;                                                // [NI] number given to vertex li
      unsigned short* vmask = vperm + NI;                                                     // [NGRP] residues taken per group
      int* vsh = reinterpret_cast<int*>(vmask + NGRP + (NGRP & 1));                           // [16] class sizes, [16] first key
      const int noff = tile_node_off[t];
      int li = hoff - 1;  // run index of the item before this thread's first item
#pragma unroll
      for (int g = 0; g < NGc; ++g)
      {
        if (keys[g] == MPCX_CT_INVALID) continue;
        const int sidx = cl * NGc + g;  // valid keys sort first
        if (head[g]) vstart[++li] = (unsigned short)sidx;
        vgrp[sidx] = (unsigned char)(((vals[g] / NGc) >> 4) * NGc + vals[g] % NGc);
      }
      if (cl == 0) { vstart[total] = (unsigned short)(nc_t * NGc); vsh[16] = (int)keys[0]; }
      for (int i = cl; i < NGRP; i += NT) vmask[i] = 0;
      if (cl < 16) vsh[cl] = 0;
      __syncthreads();
      if (cl == 0)
      {
        const int capmax = nnp >> 4;
        for (int u = 0; u < total; ++u)
        {
          const int s0 = vstart[u], s1 = vstart[u + 1];
          int best = -1;
          if (MPCX_CT_BANKOPT != 0)
          {
            unsigned taken = 0u;
            for (int q = s0; q < s1; ++q) taken |= vmask[vgrp[q]];
            int bestc = 1 << 30;
            for (int r = 0; r < 16; ++r)
            {
              if (vsh[r] >= capmax) continue;
              int c = vsh[r];  // among conflict-free residues: the emptiest class
              if ((taken >> r) & 1u)
              {
                c = 1 << 16;  // no free residue left in some group: the one fewest of the vertex's groups hold
                for (int q = s0; q < s1; ++q) c += ((vmask[vgrp[q]] >> r) & 1u) << 8;
                c += vsh[r];
              }
              if (c < bestc) { bestc = c; best = r; }
            }
          }
          else
            best = u & 15, best = vsh[best] < capmax ? best : -1;
          if (best < 0)
            for (int r = 0; r < 16 && best < 0; ++r)
              if (vsh[r] < capmax) best = r;
          vperm[u] = (unsigned short)(best + 16 * vsh[best]);
          ++vsh[best];
          for (int q = s0; q < s1; ++q) vmask[vgrp[q]] |= (unsigned short)(1u << best);
        }
      }
      __syncthreads();
      const int key0 = vsh[16];
      for (int i = cl; i < nnp; i += NT) node_ids[noff + i] = key0;
      __syncthreads();
      li = hoff - 1;
#pragma unroll
      for (int g = 0; g < NGc; ++g)
      {
        if (keys[g] == MPCX_CT_INVALID) continue;
        if (head[g])
        {
          ++li;
          node_ids[noff + vperm[li]] = (int)keys[g];
        }
        cell_nodes[first * NGc + vals[g]] = vperm[li];
      }
    }
  }
  __syncthreads();

  // ---- dests and sources
  unsigned keys[NEc];
  unsigned short vals[NEc];
  const int n1 = nd1 * bs1;
  int degenerate = 0;
#pragma unroll
  for (int e = 0; e < NEc; ++e)
  {
    keys[e] = MPCX_CT_INVALID;
    vals[e] = (unsigned short)(cl * NEc + e);  // (cell of the tile, local entry)
    if (active && vec)
      keys[e] = (unsigned)dm0[(long long)cell * nd0 + e];  // vector plan (bs == 1): dest = row dof of local entry e
    else if (active)
    {
      const int p = e / n1, q = e - p * n1;
      const int r = dm0[(long long)cell * nd0 + p / bs0] * bs0 + p % bs0;
      const int c = dm1[(long long)cell * nd1 + q / bs1] * bs1 + q % bs1;
      // bc rows / columns are zeroed before insertion (cpp/assemble_matrix.cpp:513-533): no source at all
      if (sym && p != q && r == c) degenerate = 1;  // a cell listing one dof twice: no symmetric plan
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

  // ---- staging layout: the dests in key order, cut into runs wherever two keys differ by more than the gap; a run
  // covers [even floor of its first key, even ceiling past its last key): gaps and padding stay zero.  The gap starts
  // at RUNGAP and is halved while the tile's staging buffer would exceed STAGECAP positions (a few tiles at corners
  // of the Morton curve would otherwise size the shared memory of every CTA).
  int rid[NEc], roff = 0, nruns = 0, stage = 0, loff = 0;
  int rlen[NEc];
  bool runs_ok = false;
  // candidates: RUNGAP, then RUNGAP / 2 ... RUNGAP / 16 while the staging buffer is above STAGECAP -- provided the
  // number of runs (bulk operations of the copy engine, a few hundred cycles each) stays below RUNCAP; otherwise back
  // to RUNGAP, whatever the size
  for (int trial = 0;; ++trial)
  {
    const unsigned gap = trial < 5 ? (unsigned)MPCX_CT_RUNGAP >> trial : (unsigned)MPCX_CT_RUNGAP;
    int nf = 0;
#pragma unroll
    for (int e = 0; e < NEc; ++e)
    {
      const int d = cl * NEc + e;
      rid[e] = (d < total && (d == 0 || dkey[d] - dkey[d - 1] > gap)) ? 1 : 0;
      nf += rid[e];
    }
    Scan(scan).ExclusiveSum(nf, roff, nruns);
    __syncthreads();
    runs_ok = nruns <= MR;
    {
      int r = roff - 1;
#pragma unroll
      for (int e = 0; e < NEc; ++e)
      {
        const int d = cl * NEc + e;
        const bool hd = rid[e] != 0;
        if (hd) ++r;
        rid[e] = r;
        if (d >= total || !runs_ok) continue;
        if (hd) rk0[r] = (int)(dkey[d] & ~1u);
        if (d == total - 1 || dkey[d + 1] - dkey[d] > gap) rend[r] = (int)((dkey[d] + 2u) & ~1u);
      }
    }
    __syncthreads();
    int lsum = 0;
#pragma unroll
    for (int e = 0; e < NEc; ++e)
    {
      const int r = cl * NEc + e;
      rlen[e] = (runs_ok && r < nruns) ? rend[r] - rk0[r] : 0;
      lsum += rlen[e];
    }
    Scan(scan).ExclusiveSum(lsum, loff, stage);
    __syncthreads();
    if (vec || trial == 5 || (runs_ok && stage <= MPCX_CT_STAGECAP && (trial == 0 || nruns <= MPCX_CT_RUNCAP))) break;
  }
  bool ok = runs_ok && stage < 65536;  // otherwise the tile does not fit the plan format (the host reports it)
  if (ok)
  {
    const long long ro = pass == 1 ? tile_run_off[t] : 0;
#pragma unroll
    for (int e = 0; e < NEc; ++e)
    {
      const int r = cl * NEc + e;
      if (r >= nruns) continue;
      rend[r] = loff;  // from here on: staging offset of the run
      if (pass == 1) runs[ro + r] = make_int2(rk0[r], loff | (rlen[e] << 16));
      loff += rlen[e];
    }
  }
  __syncthreads();
  int bad = degenerate;
#pragma unroll
  for (int e = 0; e < NEc; ++e)
  {
    const int d = cl * NEc + e;
    if (d >= total || !ok) continue;
    dj[d] = (unsigned short)(rend[rid[e]] + (int)dkey[d] - rk0[rid[e]]);
    rcnt[d] = (unsigned short)((int)dstart[d + 1] - (int)dstart[d]);
    part[d] = (unsigned short)d;
  }
  __syncthreads();
  // ---- symmetric plan: the record of an upper-triangular dest (row <= col) also feeds the transposed entry, lower
  // dests get no record; the transposed entry lies in the same tile because the same cells contribute to both
  if (sym && ok)
  {
#pragma unroll
    for (int e = 0; e < NEc; ++e)
    {
      const int d = cl * NEc + e;
      if (d >= total) continue;
      const long long k = dkey[d];
      long long lo = 0, hi = nrows;  // row of entry k: last row with row_ptr <= k
      while (hi - lo > 1)
      {
        const long long mid = (lo + hi) >> 1;
        if (A.rp[mid] <= k) lo = mid; else hi = mid;
      }
      const int a = (int)lo, b = A.col[k];
      if (a == b) { sp2[d] = dj[d]; continue; }
      const long long kT = csr_find(A, b, a);
      int l2 = 0, h2 = total;  // position of the transposed entry among the tile's dests
      while (l2 < h2)
      {
        const int mid = (l2 + h2) >> 1;
        if ((long long)dkey[mid] < kT) l2 = mid + 1; else h2 = mid;
      }
      if (kT < 0 || l2 >= total || (long long)dkey[l2] != kT
          || (int)dstart[l2 + 1] - (int)dstart[l2] != (int)dstart[d + 1] - (int)dstart[d])
      {
        bad = 1;  // pattern or contributions not symmetric after all
        continue;
      }
      if (a < b) sp2[d] = dj[l2];
      else { rcnt[d] = 0; part[d] = (unsigned short)l2; }
    }
  }
  bad = __syncthreads_or(bad);
  if (bad) ok = false;

  // ---- records in order of DESCENDING source count (stable LSD radix sort, so the plan is deterministic): the lanes
  // of a warp of the reduction phase then run the same number of iterations (diagonal entries collect ~24 element
  // entries, off-diagonals 4-6; in CSR order a warp would idle for most of its longest lane's loop).
  unsigned ckey[NEc];
  unsigned short cval[NEc];
  int nr_local = 0, big = 0;
#pragma unroll
  for (int e = 0; e < NEc; ++e)
  {
    const int d = cl * NEc + e;
    const int c = (ok && d < total) ? (int)rcnt[d] : 0;
    ckey[e] = c > 0 ? (unsigned)(63 - (c < 63 ? c : 63)) : 64u;
    cval[e] = (unsigned short)d;
    nr_local += c > 0;
    big |= c > 255;
  }
  SortD(sortd).Sort(ckey, cval, 0, 7);
  __syncthreads();
  int dummy, nrec;
  Scan(scan).ExclusiveSum(nr_local, dummy, nrec);
  big = __syncthreads_or(big);
  if (big) ok = false;  // a record count must fit 8 bits
#pragma unroll
  for (int e = 0; e < NEc; ++e)
  {
    const int p = cl * NEc + e;  // position in count order; the records sort first
    if (ckey[e] == 64u) continue;
    pcnt[p] = rcnt[cval[e]];
    npos[cval[e]] = (unsigned short)p;
  }
  __syncthreads();
  // groups of 32 records; group g owns the slot block [gbase[g], gbase[g] + GSTRIDE * (largest count in g))
  const int ngroups = (nrec + 31) >> 5;
  int gmax = 0;
  if (cl < ngroups)
    for (int l = 0; l < 32; ++l)
    {
      const int p = cl * 32 + l;
      const int c = p < nrec ? (int)pcnt[p] : 0;
      gmax = c > gmax ? c : gmax;
    }
  int goff, slots;
  Scan(scan).ExclusiveSum(gmax * MPCX_CT_GSTRIDE, goff, slots);
  __syncthreads();
  if (pass == 0)
  {
    // a tile that does not fit reports a size no launch can satisfy: the host turns it into MPCX_ERR_UNSUPPORTED
    if (cl == 0)
    {
      tile_nd[t] = ok ? nrec : (1 << 30);
      tile_slots[t] = slots; tile_nr[t] = ok ? nruns : 0; tile_stage[t] = ok ? stage : 0;
    }
    return;
  }
  if (!ok || slots >= (int)MPCX_CT_NOSLOT) { g_dev_err = MPCX_ERR_UNSUPPORTED; return; }  // the host checked pass 0
  const long long doff = tile_dest_off[t];  // multiple of 128
  if (cl < ngroups)
  {
    gbase[cl] = goff;
    ginfo[(doff >> 5) + cl] = (unsigned)goff | ((unsigned)gmax << 16);  // first slot | largest count of the group
  }
#pragma unroll
  for (int e = 0; e < NEc; ++e)
  {
    if (ckey[e] == 64u) continue;
    const int p = cl * NEc + e, d = cval[e];
    dest_cnt[doff + p] = (uint8_t)pcnt[p];
    dest_spos[doff + p] = dj[d];
    if (sym) dest_spos2[doff + p] = sp2[d];
    if (vec) dest_k[doff + p] = (int)dkey[d];  // row dof: lets the kernel stage per-row data once per tile
  }
  __syncthreads();
  // ---- bank-aware order of the sources of every record (matrix plans).  The i-th source of the record held by lane l
  // of group g lives in slot gbase[g] + 33 i + l, i.e. in 8-byte bank (gbase[g] + l + i) mod 16; WHICH contributing cell
  // gets which i is free.  The cell threads of one half-warp store element entry `si` with one STS.64: 16 slots that
  // should fall into 16 different banks (random placement costs ~3 wavefronts instead of 1: 44 % of all shared-memory
  // wavefronts of the fused kernel were bank conflicts, profiles/r02_a).  Deterministic coloured greedy: the records
  // of one colour pick, source by source, the free position whose bank is least loaded in the table
  // T[half-warp][si][bank] (read-only during the pick), then all of them book their picks; a second sweep re-picks
  // every record against the complete table.
  const bool opt = !vec && ok && (MPCX_CT_BANKOPT != 0);
  unsigned short* srcv = dj;                                    // [N] (cell, entry) of the source at sorted position s
  unsigned char* isrc = reinterpret_cast<unsigned char*>(sp2);  // [N] position picked for the source at sorted position s
  unsigned* T32 = reinterpret_cast<unsigned*>(pcnt);            // [NT/16][NS][16] byte counters
  if (opt)
  {
    constexpr int NVc = NGc;
    const int NSo = sym ? NVc * (NVc + 1) / 2 : NEc;
    const int twords = (NT / 16) * NSo * 4;
#pragma unroll
    for (int e = 0; e < NEc; ++e) srcv[cl * NEc + e] = vals[e];
    for (int i = cl; i < twords; i += NT) T32[i] = 0u;
    __syncthreads();
    auto slot_group = [&](unsigned short sv) {
      const int cidx = sv / NEc, ent = sv - cidx * NEc;
      int si = ent;
      if (sym)
      {
        int i = ent / NVc, j = ent - i * NVc;
        if (i > j) { const int tmp = i; i = j; j = tmp; }  // the cell stores the upper-triangular twin
        si = i * NVc - (i * (i - 1)) / 2 + (j - i);
      }
      return ((cidx >> 4) * NSo + si) * 16;
    };
    const unsigned char* T8 = reinterpret_cast<const unsigned char*>(T32);
    constexpr int NCOL = 8, NSWEEP = 2;
    for (int round = 0; round < NCOL * NSWEEP; ++round)
    {
      const unsigned col = (unsigned)(round % NCOL);
      const bool refine = round >= NCOL;
      // pick (T read-only), after taking the record's own entries out of the table in the refinement sweep
      for (int pass2 = 0; pass2 < 3; ++pass2)
      {
        // pass2 0: un-book (refinement only), 1: pick, 2: book
        if (pass2 == 0 && !refine) continue;
        for (int d = cl; d < total; d += NT)
        {
          if (rcnt[d] == 0 || ((((unsigned)d * 2654435761u) >> 29) & (NCOL - 1)) != col) continue;
          const int s0 = dstart[d], cnt = (int)dstart[d + 1] - s0;
          const int p = npos[d];
          const int B = gbase[p >> 5] + (p & 31);
          if (pass2 == 1)
          {
            unsigned used = 0u;
            for (int k = 0; k < cnt; ++k)
            {
              int best = k;
              if (cnt <= 32)
              {
                const int gi = slot_group(srcv[s0 + k]);
                int bestc = 1 << 30;
                for (int j = 0; j < cnt; ++j)
                {
                  const int i = k + j < cnt ? k + j : k + j - cnt;  // ties keep the cell order
                  if ((used >> i) & 1u) continue;
                  const int c = T8[gi + ((B + i) & 15)];
                  if (c < bestc) { bestc = c; best = i; }
                }
                used |= 1u << best;
              }
              isrc[s0 + k] = (unsigned char)best;
            }
          }
          else
          {
            for (int k = 0; k < cnt; ++k)
            {
              const int idx = slot_group(srcv[s0 + k]) + ((B + (int)isrc[s0 + k]) & 15);
              const unsigned one = 1u << (8 * (idx & 3));
              if (pass2 == 2) atomicAdd(&T32[idx >> 2], one); else atomicSub(&T32[idx >> 2], one);
            }
          }
        }
        __syncthreads();
      }
    }
  }
  // Vector plans: the fused kernel GATHERS through the inverse slot map -- the thread of row record p reads, at step i,
  // the 16-byte pair of the cell in its slot i with one LDS.128; the 8 records of a quarter-warp (p / 8) are served in
  // one wavefront when their 8 cells differ mod 8.  Which cell of a row sits in which slot is free: one thread per
  // quarter-warp walks the steps and gives every record, in turn, a remaining cell whose residue no other record of
  // the quarter holds at that step (records past their count read the zero pair at cellv[C + (p & 7)] until their
  // last group of four steps ends: residue p & 7).
  const bool optv = vec && ok && (MPCX_CT_BANKOPT != 0);
  if (optv)
  {
    unsigned short* pinv = pcnt;  // [nrec] dest of record p (the counts were written out above)
#pragma unroll
    for (int e = 0; e < NEc; ++e) srcv[cl * NEc + e] = vals[e];
    for (int d = cl; d < total; d += NT)
      if (rcnt[d] > 0) pinv[npos[d]] = (unsigned short)d;
    __syncthreads();
    for (int q = cl; 8 * q < nrec; q += NT)
    {
      int s0[8], cn[8];
      unsigned used_src[8];
      int maxc = 0;
      for (int l = 0; l < 8; ++l)
      {
        const int p = 8 * q + l;
        s0[l] = 0; cn[l] = 0; used_src[l] = 0u;
        if (p >= nrec) continue;
        const int d = pinv[p];
        s0[l] = dstart[d]; cn[l] = (int)dstart[d + 1] - s0[l];
        if (cn[l] > 32)  // (never on simplicial meshes) identity order, takes no part in the pattern
        {
          for (int k = 0; k < cn[l]; ++k) isrc[s0[l] + k] = (unsigned char)k;
          cn[l] = 0;
        }
        maxc = cn[l] > maxc ? cn[l] : maxc;
      }
      for (int i = 0; i < maxc; ++i)
      {
        unsigned used_res = 0u;
        for (int l = 0; l < 8; ++l)
          if (cn[l] > 0 && cn[l] <= i && i < ((cn[l] + 3) & ~3)) used_res |= 1u << l;
        for (int ll = 0; ll < 8; ++ll)
        {
          const int l = (ll + i) & 7;  // the first pick rotates
          if (i >= cn[l]) continue;
          int pick = -1, fallback = -1;
          for (int k = 0; k < cn[l] && pick < 0; ++k)
          {
            if ((used_src[l] >> k) & 1u) continue;
            if (fallback < 0) fallback = k;
            if (!((used_res >> ((srcv[s0[l] + k] / NEc) & 7)) & 1u)) pick = k;
          }
          if (pick < 0) pick = fallback;
          used_src[l] |= 1u << pick;
          used_res |= 1u << ((srcv[s0[l] + pick] / NEc) & 7);
          isrc[s0[l] + pick] = (unsigned char)i;
        }
      }
    }
    __syncthreads();
  }
  {
    constexpr int NVc = NGc;                                  // P1: entries form an NVc x NVc matrix
    const int NS = sym ? NVc * (NVc + 1) / 2 : NEc;           // slots per cell record
    int d = hoff - 1;  // items before the thread's first head continue the previous thread's dest
#pragma unroll
    for (int e = 0; e < NEc; ++e)
    {
      const bool valid = keys[e] != MPCX_CT_INVALID;
      if (valid && head[e]) ++d;
      const int cidx = vals[e] / NEc, ent = vals[e] - cidx * NEc;
      int si = ent;
      if (sym)
      {
        const int i = ent / NVc, j = ent - i * NVc;
        if (i > j) continue;  // the kernel stores the upper triangle only
        si = i * NVc - (i * (i - 1)) / 2 + (j - i);
      }
      uint16_t sl = (uint16_t)slots;  // bc-zeroed entry: stored to the spare slot nobody reads
      if (valid)
      {
        // i-th source of dest d (same cell order in the transposed dest, whose record holds the picked positions)
        const int rank = cl * NEc + e - (int)dstart[d];
        const int i_src = (opt || optv) ? (int)isrc[(int)dstart[part[d]] + rank] : rank;
        const int p = npos[part[d]];
        sl = (uint16_t)(gbase[p >> 5] + MPCX_CT_GSTRIDE * i_src + (p & 31));
        if (vec)
        {
          cell_rows[first * NEc + vals[e]] = (uint16_t)p;
          slot_cell[tile_slot_off[t] + sl] = (uint16_t)cidx;
        }
      }
      cell_slot[first * NS + cidx * NS + si] = sl;
    }
  }
}

// ------------------------------------------------------------------ the assembly kernels
// Shared memory of one CTA (sections 16-byte aligned; record arrays sized for whole groups of 128 records):
//   Xs[max_nodes][3] f64 | stage[max_stage] f64 (matrix) | ebuf[max_slots + 1] f64 | fs[max_dests] f64 (vector) |
//   R (two copies in the matrix kernel): runs[max_runs] int2 (matrix) | gi[max_dests/32] u32 | dk[max_dests] i32
//      (vector) | spos[max_dests] u16, spos2[max_dests] u16 (matrix / symmetric) | dcnt[max_dests] u8 |
//   C: cnode[C][NV] u16 | cslot[C][NS] u16 | crow[C][NV] u16 (vector) | 3 mbarriers
__host__ __device__ inline size_t tile_smem_bytes(int max_nodes, int max_dests, int max_slots, int max_runs, int max_stage,
                                                  int C, int nv, int ns, bool vec, bool sym)
{
  auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
  size_t b = al(24 * (size_t)max_nodes) + al(8 * (size_t)(max_slots + 1));
  const size_t rec = al(4 * (size_t)(max_dests / 32)) + al((size_t)max_dests);  // gi, dcnt
  if (vec) b += rec + al(8 * (size_t)max_dests) + al(4 * (size_t)max_dests) + al(2 * (size_t)C * nv);
  else b += al(8 * (size_t)max_stage) + 2 * (rec + al(8 * (size_t)max_runs) + (sym ? 2 : 1) * al(2 * (size_t)max_dests));
  return b + al(2 * (size_t)C * nv) + al(2 * (size_t)C * ns) + 32;
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

__device__ __forceinline__ void tma_reduce_add_f64(double* dst, const double* src_smem, unsigned bytes)
{
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}

struct TileRec  // dest-side records of one tile
{
  int2* runs;
  unsigned* gi;
  int* dk;
  uint16_t *spos, *spos2;
  uint8_t* dcnt;
  unsigned long long* bar;
};
struct TileSmem
{
  double *Xs, *stage, *ebuf, *fs;
  TileRec R0;          // first copy of the dest-side records
  unsigned rstride;    // bytes to the second copy (matrix kernel)
  uint16_t *cnode, *cslot, *crow;
  unsigned long long* barC;
};

// copy c (0 / 1) of the dest-side records: pointer arithmetic instead of an indexed array keeps TileSmem in registers
__device__ __forceinline__ TileRec tile_rec(const TileSmem& S, unsigned c)
{
  const size_t o = (size_t)c * S.rstride;
  TileRec R;
  R.runs = reinterpret_cast<int2*>(reinterpret_cast<unsigned char*>(S.R0.runs) + o);
  R.gi = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(S.R0.gi) + o);
  R.dk = S.R0.dk;
  R.spos = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(S.R0.spos) + o);
  R.spos2 = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(S.R0.spos2) + o);
  R.dcnt = S.R0.dcnt + o;
  R.bar = S.R0.bar + c;
  return R;
}

__device__ __forceinline__ TileSmem tile_carve(unsigned char* sp, const TilePlanD& P, int nv, int ns, bool vec, bool sym)
{
  auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
  TileSmem S;
  S.Xs = reinterpret_cast<double*>(sp); sp += al(24 * (size_t)P.max_nodes);
  S.stage = reinterpret_cast<double*>(sp); if (!vec) sp += al(8 * (size_t)P.max_stage);
  S.ebuf = reinterpret_cast<double*>(sp); sp += al(8 * (size_t)(P.max_slots + 1));
  S.fs = reinterpret_cast<double*>(sp); if (vec) sp += al(8 * (size_t)P.max_dests);
  unsigned char* r0 = sp;
  TileRec& R = S.R0;
  R.runs = reinterpret_cast<int2*>(sp); if (!vec) sp += al(8 * (size_t)P.max_runs);
  R.gi = reinterpret_cast<unsigned*>(sp); sp += al(4 * (size_t)(P.max_dests / 32));
  R.dk = reinterpret_cast<int*>(sp); if (vec) sp += al(4 * (size_t)P.max_dests);
  R.spos = reinterpret_cast<uint16_t*>(sp); if (!vec) sp += al(2 * (size_t)P.max_dests);
  R.spos2 = reinterpret_cast<uint16_t*>(sp); if (!vec && sym) sp += al(2 * (size_t)P.max_dests);
  R.dcnt = reinterpret_cast<uint8_t*>(sp); sp += al((size_t)P.max_dests);
  S.rstride = (unsigned)(sp - r0);
  if (!vec) sp += S.rstride;  // second copy
  S.cnode = reinterpret_cast<uint16_t*>(sp); sp += al(2 * (size_t)P.C * nv);
  S.cslot = reinterpret_cast<uint16_t*>(sp); sp += al(2 * (size_t)P.C * ns);
  S.crow = reinterpret_cast<uint16_t*>(sp); if (vec) sp += al(2 * (size_t)P.C * nv);
  R.bar = reinterpret_cast<unsigned long long*>(sp);
  S.barC = R.bar + 2;
  return S;
}

// Per-tile header (3 x int4): {vertex offset, vertices, records, runs} {staging size, -, record offset lo, hi}
// {run offset lo, hi, -, -}.  Every thread reads the same 48 bytes (one broadcast transaction per warp).
struct TileHdr
{
  int node_off, nn, nd, nr, stage;
  long long dest_off, run_off;
};
__device__ __forceinline__ TileHdr load_hdr(const int4* __restrict__ hdr, int t)
{
  const int4 a = __ldg(hdr + 3 * (long long)t), b = __ldg(hdr + 3 * (long long)t + 1), c = __ldg(hdr + 3 * (long long)t + 2);
  TileHdr h;
  h.node_off = a.x; h.nn = a.y; h.nd = a.z; h.nr = a.w; h.stage = b.x;
  h.dest_off = (long long)(unsigned)b.z | ((long long)b.w << 32);
  h.run_off = (long long)(unsigned)c.x | ((long long)c.y << 32);
  return h;
}

// TMA bulk copies of one tile's dest-side records (R) and cell-side records (C), issued by one thread
template <bool VEC, bool SYM>
__device__ __forceinline__ void tma_records(const TileRec& R, const TilePlanD& P, const TileHdr& h)
{
  const unsigned nd16 = (unsigned)((h.nd + 15) & ~15), ng4 = (unsigned)((((h.nd + 31) >> 5) + 3) & ~3);
  const unsigned nr2 = VEC ? 0u : (unsigned)((h.nr + 1) & ~1);
  mbar_expect_tx(R.bar, nd16 * (VEC ? 5 : (SYM ? 5 : 3)) + ng4 * 4 + nr2 * 8);
  if (!nd16) return;
  tma_load_1d(R.dcnt, P.dest_cnt + h.dest_off, nd16, R.bar);
  tma_load_1d(R.gi, P.ginfo + (h.dest_off >> 5), ng4 * 4, R.bar);
  if (VEC)
    tma_load_1d(R.dk, P.dest_k + h.dest_off, nd16 * 4, R.bar);
  else
  {
    tma_load_1d(R.spos, P.dest_spos + h.dest_off, nd16 * 2, R.bar);
    if (SYM) tma_load_1d(R.spos2, P.dest_spos2 + h.dest_off, nd16 * 2, R.bar);
    if (nr2) tma_load_1d(R.runs, P.runs + h.run_off, nr2 * 8, R.bar);
  }
}
template <bool VEC>
__device__ __forceinline__ void tma_cells(const TileSmem& S, const TilePlanD& P, int t, int nv, int ns)
{
  const long long first = (long long)t * P.C;
  const unsigned bn = (unsigned)(2 * nv * P.C), bs = (unsigned)(2 * ns * P.C);
  mbar_expect_tx(S.barC, bn + bs + (VEC ? bn : 0u));
  tma_load_1d(S.cnode, P.cell_nodes + first * nv, bn, S.barC);
  tma_load_1d(S.cslot, P.cell_slot + first * ns, bs, S.barC);
  if (VEC) tma_load_1d(S.crow, P.cell_rows + first * nv, bn, S.barC);
}

__device__ __forceinline__ void load_vertex(const MeshD& mesh, int node, double& x0, double& x1, double& x2)
{
  const double* p = mesh.x + (long long)node * mesh.xs;
  if (mesh.xs == 4)
  {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p));
    x0 = a.x; x1 = a.y;
  }
  else
  {
    x0 = __ldg(p); x1 = __ldg(p + 1);
  }
  x2 = __ldg(p + 2);
}

// sum of record k = column (k & 31) of its group's slot block; the trip count is the group's largest count
// (warp-uniform: no divergence bookkeeping), rows past the lane's own count are not read
__device__ __forceinline__ double tile_record_sum(const double* ebuf, const TileRec& R, int k)
{
  const unsigned g = R.gi[k >> 5];
  const double* e = ebuf + (g & 0xffffu) + (k & 31);
  const int cmax = (int)(g >> 16), cnt = R.dcnt[k];
  double s0 = 0.0, s1 = 0.0;
  int i = 0;
#pragma unroll 1
  for (; i + 2 <= cmax; i += 2, e += 2 * MPCX_CT_GSTRIDE)
  {
    if (i < cnt) s0 += e[0];
    if (i + 1 < cnt) s1 += e[MPCX_CT_GSTRIDE];
  }
  if (i < cnt) s0 += e[0];
  return s0 + s1;
}

// Persistent CTAs (2 per SM) walk the tiles with stride gridDim.x.  While tile t computes, the records of tile t+1
// arrive by TMA, its vertex coordinates travel through registers and the copy engine is still adding tile t-1 into
// A.val: the per-tile latency chain (header -> vertex ids -> coordinates -> records) and the bulk reductions are
// off the critical path (profiles/r01_k: that chain, not a throughput limit, bounded the one-tile-per-CTA kernel).
//   top      vertex id of tile t+1 -> register
//   phase 1  wait C(t); thread = cell: element matrix -> slots; then x[vertex id] of t+1 -> registers;
//            issuing lanes: wait until the reductions of t-1 have read the staging buffer
//   sync 1   TMA C(t+1), TMA R(t+1) into the other record buffer; zero the staging buffer
//   sync 1b
//   phase 2  wait R(t); thread = record: column sum -> staging position(s); vertex registers -> Xs
//   sync 2   warp 0 issues the TMA bulk reduce-add of the tile's runs and moves on
// SYM: the element matrix is symmetric and so are dofmaps and bc markers of both sides: only the upper triangle is
// stored and summed, every record feeds entry (r, c) and entry (c, r).
template <int TD, bool SYM>
__global__ void __launch_bounds__(MPCX_TILE_THREADS, 2)
k_ptile_matrix_p1(TilePlanD P, int nt, IntD in, MeshD mesh, CsrD A)
{
  constexpr int NV = TD + 1, NS = SYM ? NV * (NV + 1) / 2 : NV * NV, NT = MPCX_TILE_THREADS;
  extern __shared__ __align__(16) unsigned char tile_smem[];
  const TileSmem S = tile_carve(tile_smem, P, NV, NS, false, SYM);
  const int tid = threadIdx.x;
  const bool issuer = tid < 32;  // warp 0 hands the runs to the copy engine, one run per lane and trip
  int t = blockIdx.x;
  if (t >= nt) return;
  if (tid == 0) { mbar_init(S.R0.bar, 1); mbar_init(S.R0.bar + 1, 1); mbar_init(S.barC, 1); }
  __syncthreads();
  TileHdr h = load_hdr(P.hdr, t), hn = h;
  if (tid == 0) { tma_records<false, SYM>(S.R0, P, h); tma_cells<false>(S, P, t, NV, NS); }
  for (int i = tid; i < h.nn; i += NT)
    load_vertex(mesh, __ldg(P.node_ids + h.node_off + i), S.Xs[3 * i], S.Xs[3 * i + 1], S.Xs[3 * i + 2]);
  int tn = t + gridDim.x;
  bool has_next = tn < nt;
  if (has_next) hn = load_hdr(P.hdr, tn);
  __syncthreads();
  for (unsigned it = 0;; ++it)
  {
    const TileRec R = tile_rec(S, it & 1);
    const long long first = (long long)t * NT;
    const int nc_t = (int)((P.n_bulk - first) < NT ? (P.n_bulk - first) : NT);
    int nid = -1;
    if (has_next && tid < hn.nn) nid = __ldg(P.node_ids + hn.node_off + tid);

    // phase 1: thread = cell; element matrix entries to their slots
    mbar_wait(S.barC, it & 1);
    if (tid < nc_t)
    {
      double X[NV][3];
#pragma unroll
      for (int v = 0; v < NV; ++v)
      {
        const int l = S.cnode[tid * NV + v];
        X[v][0] = S.Xs[3 * l];
        X[v][1] = S.Xs[3 * l + 1];
        X[v][2] = TD == 3 ? S.Xs[3 * l + 2] : 0.0;
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
      uint16_t slot[NS];
      if (NS % 2 == 0)
      {
        const unsigned* sp = reinterpret_cast<const unsigned*>(S.cslot) + tid * (NS / 2);
#pragma unroll
        for (int e = 0; e < NS / 2; ++e)
        {
          const unsigned ww = sp[e];
          slot[2 * e] = ww & 0xffff; slot[2 * e + 1] = ww >> 16;
        }
      }
      else
      {
#pragma unroll
        for (int e = 0; e < NS; ++e) slot[e] = S.cslot[tid * NS + e];
      }
      int si = 0;
#pragma unroll
      for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int j = SYM ? i : 0; j < NV; ++j) S.ebuf[slot[si++]] = Ae[i][j];  // bc-zeroed entries go to the tile's spare slot
    }
    double xg0 = 0.0, xg1 = 0.0, xg2 = 0.0;
    if (nid >= 0) load_vertex(mesh, nid, xg0, xg1, xg2);
    if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // reductions of t-1 have read the staging buffer
    __syncthreads();  // 1: element buffer complete; cell records, Xs, staging buffer and the other record buffer are free
    if (tid == 0 && has_next) { tma_cells<false>(S, P, tn, NV, NS); tma_records<false, SYM>(tile_rec(S, (it & 1) ^ 1), P, hn); }
    for (int i = tid; i < (h.stage >> 1); i += NT) reinterpret_cast<double2*>(S.stage)[i] = make_double2(0.0, 0.0);
    __syncthreads();  // 1b: staging buffer zeroed

    // phase 2: thread = record; the sum goes to its staging position (and the transposed entry's)
    mbar_wait(R.bar, (it >> 1) & 1);
    for (int k = tid; k < h.nd; k += NT)
    {
      const double v = tile_record_sum(S.ebuf, R, k);
      S.stage[R.spos[k]] = v;
      if (SYM) S.stage[R.spos2[k]] = v;
    }
    if (nid >= 0) { S.Xs[3 * tid] = xg0; S.Xs[3 * tid + 1] = xg1; S.Xs[3 * tid + 2] = xg2; }
    if (has_next)
      for (int i = tid + NT; i < hn.nn; i += NT)  // a tile with more vertices than threads (never on simplicial meshes)
        load_vertex(mesh, __ldg(P.node_ids + hn.node_off + i), S.Xs[3 * i], S.Xs[3 * i + 1], S.Xs[3 * i + 2]);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // staging writes -> visible to the copy engine
    __syncthreads();  // 2: staging buffer and next Xs complete; element buffer free

    // warp 0 issues the bulk reductions (uniform operands: the lanes take turns) and moves on to the next tile
    if (issuer)
    {
      for (int r = tid; r < h.nr; r += 32)
      {
        const int2 rr = R.runs[r];
        tma_reduce_add_f64(A.val + rr.x, S.stage + (rr.y & 0xffff), (unsigned)(rr.y >> 16) * 8u);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (!has_next) break;
    h = hn; t = tn; tn += gridDim.x; has_next = tn < nt;
    if (has_next) hn = load_hdr(P.hdr, tn);
  }
  if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the staging buffer must outlive the reads
}

// Load vector b_i = c0 |K|/((d+1)(d+2)) (f_i + sum_j f_j) (cpp/assemble_vector.cpp:163-185 with the P1 source
// kernel), same pipeline with NV entries per cell and dests = the row dofs of the tile.  A tile has ~0.4 rows per
// cell scattered through b, so the sums are added with one red.global.add.f64 per (tile, row) -- no staging, no
// phase 3.  When the coefficient lives in the test space (w_by_row) its values are staged once per tile row.
template <int TD>
__global__ void __launch_bounds__(MPCX_TILE_THREADS)
k_ptile_vector_p1(TilePlanD P, int nt, IntD in, MeshD mesh, int w_by_row, double* __restrict__ b)
{
  constexpr int NV = TD + 1, NT = MPCX_TILE_THREADS;
  extern __shared__ __align__(16) unsigned char tile_smem[];
  const TileSmem S = tile_carve(tile_smem, P, NV, NV, true, false);
  const int tid = threadIdx.x;
  int t = blockIdx.x;
  if (t >= nt) return;
  const TileRec R = S.R0;
  if (tid == 0) { mbar_init(R.bar, 1); mbar_init(S.barC, 1); }
  __syncthreads();
  TileHdr h = load_hdr(P.hdr, t), hn = h;
  if (tid == 0) { tma_records<true, false>(R, P, h); tma_cells<true>(S, P, t, NV, NV); }
  for (int i = tid; i < h.nn; i += NT)
    load_vertex(mesh, __ldg(P.node_ids + h.node_off + i), S.Xs[3 * i], S.Xs[3 * i + 1], S.Xs[3 * i + 2]);
  if (w_by_row)
    for (int k = tid; k < h.nd; k += NT) S.fs[k] = __ldg(in.wnodal + __ldg(P.dest_k + h.dest_off + k));
  int tn = t + gridDim.x;
  bool has_next = tn < nt;
  if (has_next) hn = load_hdr(P.hdr, tn);
  __syncthreads();
  unsigned phase = 0;
  for (;;)
  {
    const long long first = (long long)t * NT;
    const int nc_t = (int)((P.n_bulk - first) < NT ? (P.n_bulk - first) : NT);
    int nid = -1, fid = -1;
    if (has_next && tid < hn.nn) nid = __ldg(P.node_ids + hn.node_off + tid);
    if (has_next && w_by_row && tid < hn.nd) fid = __ldg(P.dest_k + hn.dest_off + tid);

    mbar_wait(S.barC, phase);
    if (tid < nc_t)
    {
      double X[NV][3];
#pragma unroll
      for (int v = 0; v < NV; ++v)
      {
        const int l = S.cnode[tid * NV + v];
        X[v][0] = S.Xs[3 * l];
        X[v][1] = S.Xs[3 * l + 1];
        X[v][2] = TD == 3 ? S.Xs[3 * l + 2] : 0.0;
      }
      P1Geom<TD> G;
      p1_geometry<TD>(X, G);
      double f[NV], fsum = 0.0;
      if (w_by_row)
      {
#pragma unroll
        for (int v = 0; v < NV; ++v) f[v] = S.fs[S.crow[tid * NV + v]];
      }
      else
      {
        const long long index = __ldg(P.cell_pos + first + tid);
        p1_load_w<TD>(in, index, in.cells ? __ldg(in.cells + index) : (int)index, f);
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) fsum += f[v];
      const double sc = in.c[0] * G.vol * (1.0 / double((TD + 1) * (TD + 2)));
#pragma unroll
      for (int v = 0; v < NV; ++v) S.ebuf[S.cslot[tid * NV + v]] = sc * (f[v] + fsum);  // a vector plan has no bc-zeroed entries
    }
    double xg0 = 0.0, xg1 = 0.0, xg2 = 0.0, fg = 0.0;
    if (nid >= 0) load_vertex(mesh, nid, xg0, xg1, xg2);
    if (fid >= 0) fg = __ldg(in.wnodal + fid);
    __syncthreads();  // 1: element buffer complete; cell records, Xs and fs are free
    if (tid == 0 && has_next) tma_cells<true>(S, P, tn, NV, NV);

    mbar_wait(R.bar, phase);
    for (int k = tid; k < h.nd; k += NT) atomicAdd(b + R.dk[k], tile_record_sum(S.ebuf, R, k));
    if (nid >= 0) { S.Xs[3 * tid] = xg0; S.Xs[3 * tid + 1] = xg1; S.Xs[3 * tid + 2] = xg2; }
    if (fid >= 0) S.fs[tid] = fg;
    if (has_next)
    {
      for (int i = tid + NT; i < hn.nn; i += NT)
        load_vertex(mesh, __ldg(P.node_ids + hn.node_off + i), S.Xs[3 * i], S.Xs[3 * i + 1], S.Xs[3 * i + 2]);
      if (w_by_row)
        for (int k = tid + NT; k < hn.nd; k += NT) S.fs[k] = __ldg(in.wnodal + __ldg(P.dest_k + hn.dest_off + k));
    }
    __syncthreads();  // 2: dest records are free; next Xs / fs complete
    if (!has_next) break;
    if (tid == 0) tma_records<true, false>(R, P, hn);
    h = hn; t = tn; tn += gridDim.x; has_next = tn < nt; phase ^= 1u;
    if (has_next) hn = load_hdr(P.hdr, tn);
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
  cudaFree(P->cell_pos); cudaFree(P->tile_node_off); cudaFree(P->node_ids); cudaFree(P->dest_k); cudaFree(P->tile_nd);
  cudaFree(P->tile_slots); cudaFree(P->tile_nr); cudaFree(P->tile_stage); cudaFree(P->runs); cudaFree(P->hdr); cudaFree(P->ginfo);
  cudaFree(P->tile_dest_off); cudaFree(P->tile_run_off); cudaFree(P->cell_nodes); cudaFree(P->dest_cnt); cudaFree(P->dest_spos);
  cudaFree(P->dest_spos2); cudaFree(P->cell_slot); cudaFree(P->cell_rows); cudaFree(P->slot_cell); cudaFree(P->tile_slot_off);
  cudaFree(P->sp_off); cudaFree(P->sp_pos); cudaFree(P->sp_ent); cudaFree(P->sp_ca); cudaFree(P->sp_cb);
  delete P;
}

inline TilePlanD tile_plan_view(const TilePlan* P)
{
  return TilePlanD{P->C, P->max_nodes, P->max_dests, P->max_slots, P->max_runs, P->max_stage, P->n_bulk, P->cell_pos,
                   P->tile_node_off, P->node_ids, P->dest_k, P->tile_nd, P->tile_nr, P->tile_stage, P->runs, P->hdr, P->ginfo,
                   P->tile_dest_off, P->tile_run_off, P->cell_nodes, P->dest_spos, P->dest_spos2, P->cell_slot, P->cell_rows,
                   P->dest_cnt, P->slot_cell};
}

inline unsigned tp_grid(long long n, int b = 256) { return (unsigned)((n + b - 1) / b > 0 ? (n + b - 1) / b : 1); }

template <int NEc, int NGc>
cudaError_t ct_build_launch(int pass, cudaStream_t s, const int* order, const int* cells, MeshD md, const mpcx_dofmap* dm0,
                            const mpcx_dofmap* dm1, const int8_t* bc0, const int8_t* bc1, CsrD A, int* tile_nn, TilePlan* P)
{
  constexpr int NT = MPCX_TILE_THREADS, N = NT * NEc;
  auto kern = k_ct_build<NT, NEc, NGc>;
  size_t smem = sizeof(typename cub::BlockRadixSort<unsigned, NT, NEc, unsigned short>::TempStorage);
  smem = std::max(smem, sizeof(typename cub::BlockRadixSort<unsigned, NT, NGc, unsigned short>::TempStorage));
  smem = std::max(smem, sizeof(typename cub::BlockDiscontinuity<unsigned, NT>::TempStorage));
  smem = (std::max(smem, sizeof(typename cub::BlockScan<int, NT>::TempStorage)) + 15) & ~(size_t)15;
  const int extra_off = (int)smem;
  constexpr int MR = MPCX_CT_MAXRUNS < N ? MPCX_CT_MAXRUNS : N;
  // dkey, gbase, rk0, rend | dstart, dj, rcnt, npos, pcnt, part, sp2
  smem += (size_t)N * 4 + (size_t)(N / 32 + 1) * 4 + 2 * (size_t)MR * 4 + ((size_t)N + 8 + 3 * (size_t)N + N + 32 + 2 * (size_t)N) * 2 + 16;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<P->nt, NT, smem, s>>>(pass, order, P->n_bulk, cells, md, dm0->map, dm1->map, dm0->nd, dm1->nd, dm0->bs, dm1->bs, bc0,
                               bc1, A, P->nrows, P->vec, P->sym, extra_off, tile_nn, P->tile_nd, P->tile_slots, P->tile_nr,
                               P->tile_stage, P->tile_node_off, P->tile_dest_off, P->tile_run_off, P->cell_pos, P->node_ids,
                               P->cell_nodes, P->dest_k, P->dest_cnt, P->dest_spos, P->dest_spos2, P->ginfo, P->runs,
                               P->cell_slot, P->cell_rows, P->tile_slot_off, P->slot_cell);
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
  P->nvals = vec ? dm0->num_dofs : Acsr->nnz;
  // symmetric plan: same dofmap and bc markers on both sides (every tile kernel's element matrix is symmetric)
  // (MPCX_TILE_SYM=0 forces the general plan, read at every plan creation: used by the tests)
  const char* sym_env = getenv("MPCX_TILE_SYM");
  P->sym = (!vec && dm0->map == dm1->map && dm0->bs == 1 && dm1->bs == 1 && bc0 == bc1 && Acsr->num_rows < (1ll << 31)
            && !(sym_env && sym_env[0] == '0')) ? 1 : 0;
  P->ne = vec ? dm0->nd : dm0->nd * dm0->bs * dm1->nd * dm1->bs;
  P->ns = P->sym ? dm0->nd * (dm0->nd + 1) / 2 : P->ne;
  unsigned long long *mm = nullptr, *code = nullptr, *code2 = nullptr;
  int *iota = nullptr, *order = nullptr, *tile_nn = nullptr;
  long long* nb_dev = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0, tb = 0;
  std::vector<int> h_nn, h_nd, h_sl, h_nr, h_st, noff, h_hdr;
  std::vector<long long> h_doff, h_roff, h_soff;
  bool fits = true;
  long long alloc_dests = 0;
  BBox bb;
  {
    const char* cv = getenv("MPCX_TILE_CURVE");
    bb.hilbert = (cv && cv[0] == 'h') ? 1 : 0;
  }
  int variant = 0;
  auto need_tmp = [&](size_t bytes) -> cudaError_t {
    if (bytes <= tmp_bytes) return cudaSuccess;
    cudaFree(tmp);
    tmp = nullptr;
    tmp_bytes = bytes + bytes / 8;
    return cudaMalloc(&tmp, tmp_bytes);
  };
  auto build = [&](int pass) -> cudaError_t {
    if (variant == 4) return ct_build_launch<4, 4>(pass, s, order, cells, md, dm0, dm1, bc0, bc1, A, tile_nn, P);
    if (variant == 3) return ct_build_launch<3, 3>(pass, s, order, cells, md, dm0, dm1, bc0, bc1, A, tile_nn, P);
    if (variant == 16) return ct_build_launch<16, 4>(pass, s, order, cells, md, dm0, dm1, bc0, bc1, A, tile_nn, P);
    return ct_build_launch<9, 3>(pass, s, order, cells, md, dm0, dm1, bc0, bc1, A, tile_nn, P);
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
    // one scale for the three axes (the largest extent): Morton cells are cubes in PHYSICAL space, so that the tiles
    // of a flat slab (a z-slab partition of a cube) are as compact as those of the cube itself; per-axis scaling
    // turned them into 8 x 8 x 1 pancakes with 40 % more vertices and 7 x more runs per tile
    double ext = 0.0;
    for (int k = 0; k < 3; ++k) ext = std::max(ext, dec_f64(h[3 + k]) - dec_f64(h[k]));
    for (int k = 0; k < 3; ++k)
    {
      bb.lo[k] = dec_f64(h[k]);
      bb.inv[k] = ext > 0.0 ? 1.0 / ext : 0.0;
    }
    // Lattice ALIGNED with the cells: on a structured mesh (every Kuhn tetrahedron spans exactly one cube of the grid)
    // the tiles should be unions of whole cubes, which they are when a cube is exactly 2^m Morton cells wide; with the
    // plain bounding-box scale a 255-cube axis leaves every level of the curve cutting through cubes: 1.64 tiles per
    // matrix entry and 0.44 tile vertices per cell instead of 1.46 and 0.37 (tools/complete_entries.py model; the
    // same mesh with 256 cubes per axis is aligned by accident).  n_k = extent / (mean cell extent) cells along axis k,
    // the same 2^m sub-cells per cell on every axis: isotropic in cell counts, hence in space for isotropic cells --
    // the slab case of the comment above keeps its cubic tiles.  Unstructured meshes: harmless.  MPCX_TILE_ALIGN=0: off.
    {
      const char* al = getenv("MPCX_TILE_ALIGN");
      double* sum4 = nullptr;
      if (!(al && al[0] == '0') && nc > 0 && tp_alloc(&sum4, 4) == cudaSuccess)
      {
        double hs[4] = {0, 0, 0, 0};
        cudaMemsetAsync(sum4, 0, sizeof(hs), s);
        k_tp_cell_extent<<<148 * 4, 256, 0, s>>>(md, cells, nc, skip, sum4);
        cudaMemcpyAsync(hs, sum4, sizeof(hs), cudaMemcpyDeviceToHost, s);
        cudaStreamSynchronize(s);
        cudaFree(sum4);
        long long nk[3] = {1, 1, 1}, nmax = 1;
        bool usable = hs[3] > 0.0;
        for (int k = 0; k < 3 && usable; ++k)
        {
          const double ek = dec_f64(h[3 + k]) - dec_f64(h[k]), ck = hs[k] / hs[3];
          if (ek > 0.0 && ck > 0.0) nk[k] = std::max(1ll, (long long)llround(ek / ck));
          nmax = std::max(nmax, nk[k]);
        }
        if (usable && nmax <= (1ll << 19))
        {
          int m = 0;
          while ((nmax << (m + 1)) <= (1ll << 20)) ++m;
          for (int k = 0; k < 3; ++k)
          {
            const double ek = dec_f64(h[3 + k]) - dec_f64(h[k]);
            bb.inv[k] = ek > 0.0 ? (double)(nk[k] << m) / (ek * 1048575.0) : 0.0;
          }
        }
      }
    }
    // tuning: MPCX_TILE_STRETCH="fx,fy,fz" divides the Morton scale of an axis by f (tiles f times longer along it)
    if (const char* st = getenv("MPCX_TILE_STRETCH"))
    {
      double f[3] = {1.0, 1.0, 1.0};
      if (sscanf(st, "%lf,%lf,%lf", &f[0], &f[1], &f[2]) >= 1)
        for (int k = 0; k < 3; ++k)
          if (f[k] > 0.0) bb.inv[k] /= f[k];
    }
  }
  // 2. cells along the Morton curve (skipped cells last), tiles of C consecutive cells
  TP_CK(tp_alloc(&code, nc)); TP_CK(tp_alloc(&code2, nc)); TP_CK(tp_alloc(&iota, nc)); TP_CK(tp_alloc(&order, nc));
  TP_CK(tp_alloc(&nb_dev, 1));
  k_tp_cell_codes<<<tp_grid(nc), 256, 0, s>>>(md, bb, cells, nc, skip, dm0->map, dm0->nd,
                                              dm0->num_owned_dofs > 0 && dm0->num_owned_dofs < dm0->num_dofs ? dm0->num_owned_dofs / dm0->bs : 0,
                                              code, iota);
  TP_CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, code, code2, iota, order, (int)nc, 0, 64, s));
  TP_CK(need_tmp(tb));
  TP_CK(cub::DeviceRadixSort::SortPairs(tmp, tb, code, code2, iota, order, (int)nc, 0, 64, s));
  k_tp_count_bulk<<<1, 32, 0, s>>>(code2, nc, nb_dev);
  TP_CK(cudaMemcpyAsync(&P->n_bulk, nb_dev, sizeof(long long), cudaMemcpyDeviceToHost, s));
  TP_CK(cudaStreamSynchronize(s));
  k_tp_count_below<<<1, 32, 0, s>>>(code2, nc, MPCX_TP_INTERIOR_BIT, nb_dev);
  TP_CK(cudaMemcpyAsync(&P->n_iface, nb_dev, sizeof(long long), cudaMemcpyDeviceToHost, s));
  TP_CK(cudaStreamSynchronize(s));
  cudaFree(code); code = nullptr; cudaFree(code2); code2 = nullptr; cudaFree(iota); iota = nullptr;
  P->nt = (int)((P->n_bulk + C - 1) / C);
  if (P->nt == 0) goto done;  // nothing but slave cells: an empty plan is valid

  // 3. pass 0: sizes of every tile
  TP_CK(tp_alloc(&tile_nn, P->nt)); TP_CK(tp_alloc(&P->tile_nd, P->nt)); TP_CK(tp_alloc(&P->tile_slots, P->nt));
  TP_CK(tp_alloc(&P->tile_nr, P->nt)); TP_CK(tp_alloc(&P->tile_stage, P->nt));
  TP_CK(build(0));
  h_nn.resize(P->nt); h_nd.resize(P->nt); h_sl.resize(P->nt); h_nr.resize(P->nt); h_st.resize(P->nt);
  TP_CK(cudaMemcpyAsync(h_st.data(), P->tile_stage, sizeof(int) * P->nt, cudaMemcpyDeviceToHost, s));
  TP_CK(cudaMemcpyAsync(h_nr.data(), P->tile_nr, sizeof(int) * P->nt, cudaMemcpyDeviceToHost, s));
  TP_CK(cudaMemcpyAsync(h_nn.data(), tile_nn, sizeof(int) * P->nt, cudaMemcpyDeviceToHost, s));
  TP_CK(cudaMemcpyAsync(h_nd.data(), P->tile_nd, sizeof(int) * P->nt, cudaMemcpyDeviceToHost, s));
  TP_CK(cudaMemcpyAsync(h_sl.data(), P->tile_slots, sizeof(int) * P->nt, cudaMemcpyDeviceToHost, s));
  TP_CK(cudaStreamSynchronize(s));
  noff.assign(P->nt + 1, 0);
  h_doff.assign(P->nt + 1, 0);
  h_roff.assign(P->nt + 1, 0);
  h_soff.assign(P->nt + 1, 0);
  for (int t = 0; t < P->nt; ++t)
  {
    h_soff[t + 1] = h_soff[t] + ((h_sl[t] + 1 + 7) & ~7);  // slots + the spare one, 16-byte granules of uint16
    if (h_nd[t] >= (1 << 30)) { fits = false; break; }  // too many runs / staging positions for the plan format
    h_roff[t + 1] = h_roff[t] + ((h_nr[t] + 1) & ~1);
    P->total_runs += h_nr[t];
    P->max_runs = std::max(P->max_runs, h_nr[t]);
    P->max_stage = std::max(P->max_stage, h_st[t]);
    P->total_stage += h_st[t];
    noff[t + 1] = noff[t] + h_nn[t];
    h_doff[t + 1] = h_doff[t] + ((h_nd[t] + 127) & ~127);  // 16-byte aligned dest / group records for the TMA bulk copies
    P->total_dests += h_nd[t];
    P->total_slots += h_sl[t];
    P->max_nodes = std::max(P->max_nodes, h_nn[t]);
    P->max_dests = std::max(P->max_dests, h_nd[t]);
    P->max_slots = std::max(P->max_slots, h_sl[t]);
  }
  P->total_nodes = noff[P->nt];
  alloc_dests = h_doff[P->nt] + 128;
  P->max_nodes = (P->max_nodes + 1) & ~1;
  P->max_dests = (P->max_dests + 127) & ~127;  // the TMA copies move whole groups of records
  P->max_runs = (P->max_runs + 1) & ~1;
  if (!fits || P->max_slots >= (int)MPCX_CT_NOSLOT
      || tile_smem_bytes(P->max_nodes, P->max_dests, P->max_slots, P->max_runs, P->max_stage, C, mesh->ng, P->ns, vec, P->sym != 0) > 226 * 1024)
  {
    rc = fail(MPCX_ERR_UNSUPPORTED, "tile plan: a tile needs more element-buffer slots than shared memory holds");
    goto done;
  }
  TP_CK(tp_alloc(&P->tile_node_off, P->nt + 1)); TP_CK(tp_alloc(&P->tile_dest_off, P->nt + 1));
  TP_CK(tp_alloc(&P->tile_run_off, P->nt + 1));
  TP_CK(cudaMemcpyAsync(P->tile_run_off, h_roff.data(), sizeof(long long) * (P->nt + 1), cudaMemcpyHostToDevice, s));
  h_hdr.resize(12 * (size_t)P->nt);
  for (int t = 0; t < P->nt; ++t)
  {
    int* q = h_hdr.data() + 12 * (size_t)t;
    q[0] = noff[t]; q[1] = h_nn[t]; q[2] = h_nd[t]; q[3] = h_nr[t];
    q[4] = h_st[t]; q[5] = 0; q[6] = (int)(unsigned)(h_doff[t] & 0xffffffffll); q[7] = (int)(h_doff[t] >> 32);
    q[8] = (int)(unsigned)(h_roff[t] & 0xffffffffll); q[9] = (int)(h_roff[t] >> 32);
    q[10] = (int)(unsigned)(h_soff[t] & 0xffffffffll); q[11] = (int)(h_soff[t] >> 32); q[5] = h_sl[t];
  }
  TP_CK(tp_alloc(&P->hdr, 3 * (long long)P->nt));
  TP_CK(cudaMemcpyAsync(P->hdr, h_hdr.data(), sizeof(int) * h_hdr.size(), cudaMemcpyHostToDevice, s));
  TP_CK(cudaMemcpyAsync(P->tile_node_off, noff.data(), sizeof(int) * (P->nt + 1), cudaMemcpyHostToDevice, s));
  TP_CK(cudaMemcpyAsync(P->tile_dest_off, h_doff.data(), sizeof(long long) * (P->nt + 1), cudaMemcpyHostToDevice, s));
  TP_CK(cudaStreamSynchronize(s));
  // 4. pass 1: the records
  TP_CK(tp_alloc(&P->cell_pos, (long long)P->nt * C)); TP_CK(tp_alloc(&P->node_ids, P->total_nodes));
  TP_CK(tp_alloc(&P->cell_nodes, (long long)P->nt * C * P->ng)); TP_CK(tp_alloc(&P->dest_k, vec ? alloc_dests : 1));
  TP_CK(tp_alloc(&P->dest_cnt, alloc_dests)); TP_CK(tp_alloc(&P->dest_spos, alloc_dests));
  TP_CK(tp_alloc(&P->ginfo, alloc_dests / 32)); TP_CK(tp_alloc(&P->runs, h_roff[P->nt] + 2));
  TP_CK(cudaMemsetAsync(P->runs, 0, sizeof(int2) * (size_t)(h_roff[P->nt] + 2), s));
  TP_CK(cudaMemsetAsync(P->dest_spos, 0, sizeof(uint16_t) * (size_t)alloc_dests, s));
  TP_CK(tp_alloc(&P->cell_slot, (long long)P->nt * C * P->ns));
  if (P->sym)
  {
    TP_CK(tp_alloc(&P->dest_spos2, alloc_dests));
    TP_CK(cudaMemsetAsync(P->dest_spos2, 0, sizeof(uint16_t) * (size_t)alloc_dests, s));
  }
  TP_CK(cudaMemsetAsync(P->cell_nodes, 0, sizeof(uint16_t) * (size_t)P->nt * C * P->ng, s));
  TP_CK(cudaMemsetAsync(P->dest_k, 0xff, sizeof(int) * (size_t)(vec ? alloc_dests : 1), s));  // fillers: -1
  TP_CK(cudaMemsetAsync(P->dest_cnt, 0, sizeof(uint8_t) * (size_t)alloc_dests, s));
  TP_CK(cudaMemsetAsync(P->ginfo, 0, sizeof(unsigned) * (size_t)(alloc_dests / 32), s));
  TP_CK(cudaMemsetAsync(P->cell_slot, 0, sizeof(uint16_t) * (size_t)P->nt * C * P->ns, s));
  if (vec)
  {
    TP_CK(tp_alloc(&P->cell_rows, (long long)P->nt * C * P->ne));
    TP_CK(cudaMemsetAsync(P->cell_rows, 0, sizeof(uint16_t) * (size_t)P->nt * C * P->ne, s));
    TP_CK(tp_alloc(&P->slot_cell, h_soff[P->nt] + 8));
    TP_CK(cudaMemsetAsync(P->slot_cell, 0, sizeof(uint16_t) * (size_t)(h_soff[P->nt] + 8), s));
    TP_CK(tp_alloc(&P->tile_slot_off, P->nt + 1));
    TP_CK(cudaMemcpyAsync(P->tile_slot_off, h_soff.data(), sizeof(long long) * (P->nt + 1), cudaMemcpyHostToDevice, s));
  }
  TP_CK(build(1));
  TP_CK(cudaStreamSynchronize(s));
  // bytes one assembly reads from the plan
  P->bytes = (long long)sizeof(uint16_t) * ((long long)P->nt * C * (P->ns + (vec ? P->ne : 0) + P->ng) + (P->sym ? 2 : 1) * P->total_dests)
             + P->total_dests
             + (long long)sizeof(int) * (P->total_nodes + (vec ? P->total_dests : 0) + P->total_dests / 32 + 2 * P->total_runs)
             + (long long)P->nt * 32;

done:
  cudaFree(mm); cudaFree(code); cudaFree(code2); cudaFree(iota); cudaFree(order); cudaFree(tile_nn);
  cudaFree(nb_dev); cudaFree(tmp);
  if (rc != MPCX_OK) { tile_plan_free(P); P = nullptr; }
  *out = P;
  return rc;
}
}  // namespace
