// mpcx_tile_fused.cuh -- ONE pass over the bulk cells for the stiffness matrix AND the load vector.
//
// The reference's LinearProblem.solve (python/src/dolfinx_mpc/problem.py:539-582) calls assemble_matrix and
// assemble_vector back to back; each walks every cell, gathers its vertex coordinates and evaluates the element
// geometry (cpp/assemble_matrix.cpp:488-506, cpp/assemble_vector.cpp:163-185).  Round 1 mirrored that with two tile
// kernels (k_ptile_matrix_p1 3.44 ms + k_ptile_vector_p1 1.67 ms at 256^3, the second one bound by the very
// shared-memory vertex reads the first had just done).  Here a cell is visited once: its thread reads the tile-local
// vertex ids, gathers the coordinates from the tile's vertex buffer, evaluates the P1 geometry and writes
//   * the element matrix entries to the slots of the MATRIX plan
//       records -> column sums -> CSR-ordered staging buffer -> TMA bulk reduce-add per run (as k_ptile_matrix_p1),
//   * ONE pair (s_c, s_c F_c) to a cell-indexed buffer, s_c = c0 |K| / ((d+1)(d+2)), F_c = sum of f over the cell's
//     vertices.  The P1 source vector is b_r = sum_{c containing r} s_c (f_r + F_c) = f_r S1_r + S2_r with
//     S1_r = sum s_c, S2_r = sum s_c F_c: every cell hands the SAME pair to all its vertices, so the (d+1) scattered
//     stores per cell of a slot buffer become one conflict-free 16-byte store, and the thread of row record r gathers
//     the pairs of its cells through the VECTOR plan's inverse slot map (slot -> tile cell), then issues one
//     red.global.add.f64 per (tile, row).
// Both plans were built over the same Morton order and skip flags, so they cut the cells into the same tiles.
// The dest-side records are single-buffered: their TMA load for tile t+1 is issued by warp 0 right after it has
// handed tile t's runs to the copy engine, and lands during phase 1 of tile t+1.
#pragma once

namespace
{
// Phase timeline of the fused kernel (tuning builds only: -DMPCX_TRACE, tools/probe_trace.py): clock64 of warp 0, a
// middle and the last warp at the phase boundaries of the first iterations of the first CTAs.
#ifdef MPCX_TRACE
__device__ long long* g_trace = nullptr;
#define MPCX_STAMP(k)                                                                                              \
  do {                                                                                                             \
    if (g_trace && blockIdx.x < 8 && it < 64 && (tid & 31) == 0)                                                   \
      g_trace[(((long long)blockIdx.x * 64 + it) * (MPCX_TILE_THREADS / 32) + (tid >> 5)) * 12 + (k)] = clock64(); \
  } while (0)
#else
#define MPCX_STAMP(k)
#endif

struct FusedSmem
{
  double *Xs, *stage, *ebuf, *fs;
  double2* cellv;
  TileRec R;  // matrix records
  unsigned* giv;
  int* dkv;
  uint8_t* dcntv;
  uint16_t *vinc, *cnode, *cslot, *crow;
  int* hb;  // ring of 3 tile headers (matrix plan: 12 ints, vector plan: 12 ints), filled two tiles ahead
  unsigned long long *barC, *barR;
};

__host__ __device__ inline size_t fused_smem_bytes(const TilePlanD& P, const TilePlanD& Q, int nv, int ns, bool sym)
{
  auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
  size_t b = al(24 * (size_t)P.max_nodes) + al(8 * (size_t)P.max_stage) + al(8 * (size_t)(P.max_slots + 1))
             + al(8 * (size_t)Q.max_dests) + al(16 * (size_t)(P.C + 8));
  b += al(8 * (size_t)P.max_runs) + al(4 * (size_t)(P.max_dests / 32)) + (sym ? 2 : 1) * al(2 * (size_t)P.max_dests)
       + al((size_t)P.max_dests);
  b += al(4 * (size_t)(Q.max_dests / 32)) + al(4 * (size_t)Q.max_dests) + al((size_t)Q.max_dests);
  b += al(2 * (size_t)(Q.max_slots + 9)) + 2 * al(2 * (size_t)P.C * nv) + al(2 * (size_t)P.C * ns);
  return b + 32 + 3 * 24 * sizeof(int);
}

__device__ __forceinline__ FusedSmem fused_carve(unsigned char* sp, const TilePlanD& P, const TilePlanD& Q, int nv, int ns,
                                                 bool sym)
{
  auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
  FusedSmem S;
  S.Xs = reinterpret_cast<double*>(sp); sp += al(24 * (size_t)P.max_nodes);
  S.stage = reinterpret_cast<double*>(sp); sp += al(8 * (size_t)P.max_stage);
  S.ebuf = reinterpret_cast<double*>(sp); sp += al(8 * (size_t)(P.max_slots + 1));
  S.fs = reinterpret_cast<double*>(sp); sp += al(8 * (size_t)Q.max_dests);
  S.cellv = reinterpret_cast<double2*>(sp); sp += al(16 * (size_t)(P.C + 8));
  S.R.runs = reinterpret_cast<int2*>(sp); sp += al(8 * (size_t)P.max_runs);
  S.R.gi = reinterpret_cast<unsigned*>(sp); sp += al(4 * (size_t)(P.max_dests / 32));
  S.R.dk = nullptr;
  S.R.spos = reinterpret_cast<uint16_t*>(sp); sp += al(2 * (size_t)P.max_dests);
  S.R.spos2 = reinterpret_cast<uint16_t*>(sp); if (sym) sp += al(2 * (size_t)P.max_dests);
  S.R.dcnt = reinterpret_cast<uint8_t*>(sp); sp += al((size_t)P.max_dests);
  S.giv = reinterpret_cast<unsigned*>(sp); sp += al(4 * (size_t)(Q.max_dests / 32));
  S.dkv = reinterpret_cast<int*>(sp); sp += al(4 * (size_t)Q.max_dests);
  S.dcntv = reinterpret_cast<uint8_t*>(sp); sp += al((size_t)Q.max_dests);
  S.vinc = reinterpret_cast<uint16_t*>(sp); sp += al(2 * (size_t)(Q.max_slots + 9));
  S.cnode = reinterpret_cast<uint16_t*>(sp); sp += al(2 * (size_t)P.C * nv);
  S.cslot = reinterpret_cast<uint16_t*>(sp); sp += al(2 * (size_t)P.C * ns);
  S.crow = reinterpret_cast<uint16_t*>(sp); sp += al(2 * (size_t)P.C * nv);
  S.hb = reinterpret_cast<int*>(sp); sp += 3 * 24 * sizeof(int);
  S.barC = reinterpret_cast<unsigned long long*>(sp);
  S.barR = S.barC + 1;
  S.R.bar = S.barR;
  return S;
}

// cell-side records of one tile: vertex ids and matrix slots (matrix plan), row records of the vertices (vector plan)
__device__ __forceinline__ void fused_tma_cells(const FusedSmem& S, const TilePlanD& P, const TilePlanD& Q, int t, int nv, int ns)
{
  const long long first = (long long)t * P.C;
  const unsigned bn = (unsigned)(2 * nv * P.C), bs = (unsigned)(2 * ns * P.C);
  mbar_expect_tx(S.barC, 2 * bn + bs);
  tma_load_1d(S.cnode, P.cell_nodes + first * nv, bn, S.barC);
  tma_load_1d(S.cslot, P.cell_slot + first * ns, bs, S.barC);
  tma_load_1d(S.crow, Q.cell_rows + first * nv, bn, S.barC);
}

// vector records of one tile: counts, group info, row dofs and the inverse slot map (second arrival on barR)
__device__ __forceinline__ long long hdr64(const int* h, int w) { return (long long)(unsigned)h[w] | ((long long)h[w + 1] << 32); }

__device__ __forceinline__ void fused_tma_vrecords(const FusedSmem& S, const TilePlanD& Q, const int* hv)
{
  const int nd = hv[2], slots = hv[5];
  const long long dest_off = hdr64(hv, 6), slot_off = hdr64(hv, 10);
  const unsigned nd16 = (unsigned)((nd + 15) & ~15), ng4 = (unsigned)((((nd + 31) >> 5) + 3) & ~3);
  const unsigned ns8 = nd16 ? (unsigned)((slots + 1 + 7) & ~7) : 0u;
  mbar_expect_tx(S.barR, nd16 * 5 + ng4 * 4 + ns8 * 2);
  if (!nd16) return;
  tma_load_1d(S.dcntv, Q.dest_cnt + dest_off, nd16, S.barR);
  tma_load_1d(S.giv, Q.ginfo + (dest_off >> 5), ng4 * 4, S.barR);
  tma_load_1d(S.dkv, Q.dest_k + dest_off, nd16 * 4, S.barR);
  tma_load_1d(S.vinc, Q.slot_cell + slot_off, ns8 * 2, S.barR);
}

// matrix records of one tile from its header words in shared memory (first arrival on barR)
template <bool SYM>
__device__ __forceinline__ void fused_tma_mrecords(const FusedSmem& S, const TilePlanD& P, const int* hm)
{
  TileHdr h;
  h.node_off = hm[0]; h.nn = hm[1]; h.nd = hm[2]; h.nr = hm[3]; h.stage = hm[4];
  h.dest_off = hdr64(hm, 6); h.run_off = hdr64(hm, 8);
  tma_records<false, SYM>(S.R, P, h);
}

// P: matrix plan, Q: vector plan of the same tiling.  ina: bilinear integral (Laplace / mass / variable-coefficient
// Laplace), inL: the P1 source term with its coefficient in the test space.  Persistent CTAs, 2 per SM.  Tile headers
// travel through a 3-slot ring in shared memory, loaded two tiles ahead by 24 threads (keeping them in registers of
// every thread cost spills whose reloads missed the small L1 left beside 2 x 100 KB of shared memory).
//   top      header words of tile t+2 -> register (24 threads); vertex id / row dof of tile t+1 -> registers
//   phase 1  wait C(t); thread = cell: geometry once, element matrix -> matrix slots, (s, s F) -> cellv; then
//            x[vertex id], f[row dof] of t+1 -> registers; header words -> ring; warp 0 waits until the reductions of
//            t-1 have read the staging buffer
//   sync 1   TMA C(t+1); zero the staging buffer
//   sync 1b
//   phase 2  wait R(t); thread = record: matrix column sum -> staging position(s); row record: gather the cell pairs
//            -> red.global.add (f of the row travels in a register of its thread); registers -> Xs, fs of t+1
//   sync 2   warp 0: TMA bulk reduce-add of the runs, then TMA R(t+1)
template <int TD, bool SYM>
__global__ void __launch_bounds__(MPCX_TILE_THREADS, 2)
k_ptile_system_p1(TilePlanD P, TilePlanD Q, int t_begin, int nt, IntD ina, IntD inL, MeshD mesh, CsrD A, double* __restrict__ b)
{
  constexpr int NV = TD + 1, NS = SYM ? NV * (NV + 1) / 2 : NV * NV, NT = MPCX_TILE_THREADS;
  extern __shared__ __align__(16) unsigned char tile_smem[];
  const FusedSmem S = fused_carve(tile_smem, P, Q, NV, NS, SYM);
  const int tid = threadIdx.x;
  const bool issuer = tid < 32;
  const int G = gridDim.x;
  int t = t_begin + blockIdx.x;  // tiles [t_begin, nt)
  if (t >= nt) return;
  // header loader threads: word w of the matrix (w < 12) or vector (w >= 12) header
  const int hw_i = tid - 32;
  const bool loader = hw_i >= 0 && hw_i < 24;
  const int* hsrc = loader ? reinterpret_cast<const int*>(hw_i < 12 ? P.hdr : Q.hdr) + (hw_i < 12 ? hw_i : hw_i - 12) : nullptr;
  if (tid == 0) { mbar_init(S.barC, 1); mbar_init(S.barR, 2); }
  if (tid < 8) S.cellv[NT + tid] = make_double2(0.0, 0.0);  // zero pairs read by records past their count, one per bank group
  if (loader)
  {
    S.hb[hw_i] = __ldg(hsrc + 12 * (long long)t);
    if (t + G < nt) S.hb[24 + hw_i] = __ldg(hsrc + 12 * (long long)(t + G));
  }
  __syncthreads();
  {
    const int* hm = S.hb;
    const int* hv = S.hb + 12;
    if (tid == 0)
    {
      fused_tma_mrecords<SYM>(S, P, hm);
      fused_tma_vrecords(S, Q, hv);
      fused_tma_cells(S, P, Q, t, NV, NS);
    }
    const int nn = hm[1], node_off = hm[0], ndv = hv[2];
    const long long doffv = hdr64(hv, 6);
    for (int i = tid; i < nn; i += NT)
      load_vertex(mesh, __ldg(P.node_ids + node_off + i), S.Xs[3 * i], S.Xs[3 * i + 1], S.Xs[3 * i + 2]);
    for (int k = tid; k < ndv; k += NT) S.fs[k] = __ldg(inL.wnodal + __ldg(Q.dest_k + doffv + k));
  }
  const double cL = inL.c[0] * (1.0 / double((TD + 1) * (TD + 2)));
  // thread NT-1-k serves row record k and carries f of that row in a register
  double fcur = 0.0;
  {
    const int* hv = S.hb + 12;
    if (NT - 1 - tid < hv[2]) fcur = __ldg(inL.wnodal + __ldg(Q.dest_k + hdr64(hv, 6) + (NT - 1 - tid)));
  }
  __syncthreads();
  int tn = t + G;
  for (unsigned it = 0;; ++it)
  {
    const bool has_next = tn < nt;
    const int* hm = S.hb + 24 * (it % 3);          // this tile
    const int* hmn = S.hb + 24 * ((it + 1) % 3);   // next tile of this CTA
    const long long first = (long long)t * NT;
    const int nc_t = (int)((P.n_bulk - first) < NT ? (P.n_bulk - first) : NT);
    int hword = 0;
    const bool load2 = loader && tn + G < nt;
    if (load2) hword = __ldg(hsrc + 12 * (long long)(tn + G));
    int nid = -1, fid = -1;
    if (has_next)
    {
      if (tid < hmn[1]) nid = __ldg(P.node_ids + hmn[0] + tid);
      if (NT - 1 - tid < hmn[12 + 2]) fid = __ldg(Q.dest_k + hdr64(hmn + 12, 6) + (NT - 1 - tid));
    }

    // phase 1: thread = cell
    MPCX_STAMP(0);
    mbar_wait(S.barC, it & 1);
    MPCX_STAMP(1);
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
      P1Geom<TD> Gm;
      p1_geometry<TD>(X, Gm);
      {
        double F = 0.0;
#pragma unroll
        for (int v = 0; v < NV; ++v) F += S.fs[S.crow[tid * NV + v]];
        const double sc = cL * Gm.vol;
        S.cellv[tid] = make_double2(sc, sc * F);
      }
      double w[NV], Ae[NV][NV];
      if (ina.kernel == MPCX_KERNEL_LAPLACE_VARCOEF)
      {
        const long long index = __ldg(P.cell_pos + first + tid);
        p1_load_w<TD>(ina, index, ina.cells ? __ldg(ina.cells + index) : (int)index, w);
      }
      p1_element<TD>(ina.kernel, Gm, ina.c, w, Ae);
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
        for (int j = SYM ? i : 0; j < NV; ++j) S.ebuf[slot[si++]] = Ae[i][j];  // bc-zeroed entries: the spare slot
    }
    double xg0 = 0.0, xg1 = 0.0, xg2 = 0.0, fg = 0.0;
    if (nid >= 0) load_vertex(mesh, nid, xg0, xg1, xg2);
    if (fid >= 0) fg = __ldg(inL.wnodal + fid);
    if (load2) S.hb[24 * ((it + 2) % 3) + hw_i] = hword;  // that slot held tile t-1's header: no reader left
    MPCX_STAMP(2);
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // this thread's reductions of t-1 have read the staging buffer
    MPCX_STAMP(3);
    __syncthreads();  // 1: element buffers complete; cell records, Xs and the staging buffer are free
    MPCX_STAMP(4);
    if (tid == 0 && has_next) fused_tma_cells(S, P, Q, tn, NV, NS);
    {
      const int half = hm[4] >> 1;
      for (int i = tid; i < half; i += NT) reinterpret_cast<double2*>(S.stage)[i] = make_double2(0.0, 0.0);
    }
    // (measured: replacing this barrier by an mbarrier that warp 0 signals after its wait -- every thread clearing its
    // share before barrier 1 -- was 2 % slower: profiles/README.md, r02_c)
    __syncthreads();  // 1b: staging buffer zeroed
    MPCX_STAMP(5);

    // phase 2: thread = record
    mbar_wait(S.barR, it & 1);
    MPCX_STAMP(6);
    {
      const int nd = hm[2];
      for (int k = tid; k < nd; k += NT)
      {
        const double v = tile_record_sum(S.ebuf, S.R, k);
        S.stage[S.R.spos[k]] = v;
        if (SYM) S.stage[S.R.spos2[k]] = v;
      }
    }
    MPCX_STAMP(7);
    // row records go to the LAST threads (the first ones hold the heaviest matrix records: those are ordered by
    // descending source count)
    {
      const int ndv = hm[12 + 2];
      for (int k = NT - 1 - tid; k < ndv; k += NT)
      {
        const unsigned g = S.giv[k >> 5];
        const uint16_t* e = S.vinc + (g & 0xffffu) + (k & 31);
        const int cnt = S.dcntv[k], zs = NT + (k & 7);
        double s1 = 0.0, s2 = 0.0;
#pragma unroll 1
        for (int i = 0; i < cnt; i += 4, e += 4 * MPCX_CT_GSTRIDE)
        {
          // four independent index -> pair loads in flight; lanes past their count read the zero pair of their bank group
          const int c0 = e[0], c1 = i + 1 < cnt ? (int)e[MPCX_CT_GSTRIDE] : zs, c2 = i + 2 < cnt ? (int)e[2 * MPCX_CT_GSTRIDE] : zs,
                    c3 = i + 3 < cnt ? (int)e[3 * MPCX_CT_GSTRIDE] : zs;
          const double2 v0 = S.cellv[c0], v1 = S.cellv[c1], v2 = S.cellv[c2], v3 = S.cellv[c3];
          s1 += (v0.x + v1.x) + (v2.x + v3.x);
          s2 += (v0.y + v1.y) + (v2.y + v3.y);
        }
        const int row = S.dkv[k];
        atomicAdd(b + row, (k < NT ? fcur : __ldg(inL.wnodal + row)) * s1 + s2);
      }
    }
    if (nid >= 0) { S.Xs[3 * tid] = xg0; S.Xs[3 * tid + 1] = xg1; S.Xs[3 * tid + 2] = xg2; }
    if (fid >= 0) S.fs[NT - 1 - tid] = fg;
    if (has_next)
    {
      const int nn = hmn[1], ndvn = hmn[12 + 2];
      if (nn > NT)  // a tile with more vertices than threads (never on simplicial meshes)
        for (int i = tid + NT; i < nn; i += NT)
          load_vertex(mesh, __ldg(P.node_ids + hmn[0] + i), S.Xs[3 * i], S.Xs[3 * i + 1], S.Xs[3 * i + 2]);
      if (ndvn > NT)
        for (int k = tid; k < ndvn - NT; k += NT) S.fs[k + NT] = __ldg(inL.wnodal + __ldg(Q.dest_k + hdr64(hmn + 12, 6) + k + NT));
    }
    MPCX_STAMP(8);
    // Every warp hands a few runs to the copy engine (run r -> lane r / NW of warp r % NW): one warp issuing all ~40
    // bulk reductions of a tile spent 3800 of the tile's 9400 cycles in that loop and was the last to arrive at barrier
    // 1 of the next tile (phase timeline, profiles/r02_trace6.txt).  The run records are read BEFORE the barrier so that
    // thread 0 can refill the record buffers right after it.
    const int nr = hm[3];
    constexpr int NW = NT / 32;
    const int my_run = (tid & 31) * NW + (tid >> 5);
    int2 rr = make_int2(0, 0);
    if (nr <= NT && my_run < nr) rr = S.R.runs[my_run];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // staging writes -> visible to the copy engine
    __syncthreads();  // 2: staging buffer, next Xs / fs complete; element buffers and records free
    MPCX_STAMP(9);

    if (nr <= NT)
    {
      if (my_run < nr) tma_reduce_add_f64(A.val + rr.x, S.stage + (rr.y & 0xffff), (unsigned)(rr.y >> 16) * 8u);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (tid == 0 && has_next) { fused_tma_mrecords<SYM>(S, P, hmn); fused_tma_vrecords(S, Q, hmn + 12); }
    }
    else if (issuer)  // more runs than threads (pathological meshes): warp 0 takes them all
    {
      for (int r = tid; r < nr; r += 32)
      {
        const int2 r2 = S.R.runs[r];
        tma_reduce_add_f64(A.val + r2.x, S.stage + (r2.y & 0xffff), (unsigned)(r2.y >> 16) * 8u);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      __syncwarp();  // every lane has read its runs: the record buffers may be refilled
      if (tid == 0 && has_next) { fused_tma_mrecords<SYM>(S, P, hmn); fused_tma_vrecords(S, Q, hmn + 12); }
    }
    MPCX_STAMP(10);
    if (!has_next) break;
    t = tn; tn += G; fcur = fg;
  }
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the staging buffer must outlive the reads
}
}  // namespace
