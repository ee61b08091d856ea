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
struct FusedSmem
{
  double *Xs, *stage, *ebuf, *fs;
  double2* cellv;
  TileRec R;  // matrix records
  unsigned* giv;
  int* dkv;
  uint8_t* dcntv;
  uint16_t *vinc, *cnode, *cslot, *crow;
  unsigned long long *barC, *barR;
};

struct VHdr  // what the vector records of one tile need from the vector plan's header
{
  int nd, slots;
  long long dest_off, slot_off;
};
__device__ __forceinline__ VHdr load_vhdr(const int4* __restrict__ hdr, int t)
{
  const int4 a = __ldg(hdr + 3 * (long long)t), b = __ldg(hdr + 3 * (long long)t + 1), c = __ldg(hdr + 3 * (long long)t + 2);
  VHdr h;
  h.nd = a.z; h.slots = b.y;
  h.dest_off = (long long)(unsigned)b.z | ((long long)b.w << 32);
  h.slot_off = (long long)(unsigned)c.z | ((long long)c.w << 32);
  return h;
}

__host__ __device__ inline size_t fused_smem_bytes(const TilePlanD& P, const TilePlanD& Q, int nv, int ns, bool sym)
{
  auto al = [](size_t b) { return (b + 15) & ~(size_t)15; };
  size_t b = al(24 * (size_t)P.max_nodes) + al(8 * (size_t)P.max_stage) + al(8 * (size_t)(P.max_slots + 1))
             + al(8 * (size_t)Q.max_dests) + al(16 * (size_t)(P.C + 1));
  b += al(8 * (size_t)P.max_runs) + al(4 * (size_t)(P.max_dests / 32)) + (sym ? 2 : 1) * al(2 * (size_t)P.max_dests)
       + al((size_t)P.max_dests);
  b += al(4 * (size_t)(Q.max_dests / 32)) + al(4 * (size_t)Q.max_dests) + al((size_t)Q.max_dests);
  b += al(2 * (size_t)(Q.max_slots + 9)) + 2 * al(2 * (size_t)P.C * nv) + al(2 * (size_t)P.C * ns);
  return b + 32;
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
  S.cellv = reinterpret_cast<double2*>(sp); sp += al(16 * (size_t)(P.C + 1));
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
__device__ __forceinline__ void fused_tma_vrecords(const FusedSmem& S, const TilePlanD& Q, const VHdr& h)
{
  const unsigned nd16 = (unsigned)((h.nd + 15) & ~15), ng4 = (unsigned)((((h.nd + 31) >> 5) + 3) & ~3);
  const unsigned ns8 = nd16 ? (unsigned)((h.slots + 1 + 7) & ~7) : 0u;
  mbar_expect_tx(S.barR, nd16 * 5 + ng4 * 4 + ns8 * 2);
  if (!nd16) return;
  tma_load_1d(S.dcntv, Q.dest_cnt + h.dest_off, nd16, S.barR);
  tma_load_1d(S.giv, Q.ginfo + (h.dest_off >> 5), ng4 * 4, S.barR);
  tma_load_1d(S.dkv, Q.dest_k + h.dest_off, nd16 * 4, S.barR);
  tma_load_1d(S.vinc, Q.slot_cell + h.slot_off, ns8 * 2, S.barR);
}

// P: matrix plan, Q: vector plan of the same tiling.  ina: bilinear integral (Laplace / mass / variable-coefficient
// Laplace), inL: the P1 source term with its coefficient in the test space.  Persistent CTAs, 2 per SM:
//   top      vertex id / row dof of tile t+1 -> registers
//   phase 1  wait C(t); thread = cell: geometry once, element matrix -> matrix slots, (s, s F) -> cellv; then
//            x[vertex id], f[row dof] of t+1 -> registers; warp 0 waits until the reductions of t-1 have read the
//            staging buffer
//   sync 1   TMA C(t+1); zero the staging buffer
//   sync 1b
//   phase 2  wait R(t); thread = record: matrix column sum -> staging position(s); row record: gather the cell pairs
//            -> red.global.add (f of the row travels in a register of its thread); registers -> Xs, fs of t+1
//   sync 2   warp 0: TMA bulk reduce-add of the runs, then TMA R(t+1)
template <int TD, bool SYM>
__global__ void __launch_bounds__(MPCX_TILE_THREADS, 2)
k_ptile_system_p1(TilePlanD P, TilePlanD Q, int nt, IntD ina, IntD inL, MeshD mesh, CsrD A, double* __restrict__ b)
{
  constexpr int NV = TD + 1, NS = SYM ? NV * (NV + 1) / 2 : NV * NV, NT = MPCX_TILE_THREADS;
  extern __shared__ __align__(16) unsigned char tile_smem[];
  const FusedSmem S = fused_carve(tile_smem, P, Q, NV, NS, SYM);
  const int tid = threadIdx.x;
  const bool issuer = tid < 32;
  int t = blockIdx.x;
  if (t >= nt) return;
  if (tid == 0) { mbar_init(S.barC, 1); mbar_init(S.barR, 2); }
  __syncthreads();
  TileHdr h = load_hdr(P.hdr, t), hn = h;
  VHdr hv = load_vhdr(Q.hdr, t), hvn = hv;
  if (tid == 0)
  {
    tma_records<false, SYM>(S.R, P, h);
    fused_tma_vrecords(S, Q, hv);
    fused_tma_cells(S, P, Q, t, NV, NS);
  }
  for (int i = tid; i < h.nn; i += NT)
    load_vertex(mesh, __ldg(P.node_ids + h.node_off + i), S.Xs[3 * i], S.Xs[3 * i + 1], S.Xs[3 * i + 2]);
  for (int k = tid; k < hv.nd; k += NT) S.fs[k] = __ldg(inL.wnodal + __ldg(Q.dest_k + hv.dest_off + k));
  int tn = t + gridDim.x;
  bool has_next = tn < nt;
  if (has_next) { hn = load_hdr(P.hdr, tn); hvn = load_vhdr(Q.hdr, tn); }
  __syncthreads();
  const double cL = inL.c[0] * (1.0 / double((TD + 1) * (TD + 2)));
  double fcur = NT - 1 - tid < hv.nd ? __ldg(inL.wnodal + __ldg(Q.dest_k + hv.dest_off + (NT - 1 - tid))) : 0.0;
  if (tid == 0) S.cellv[NT] = make_double2(0.0, 0.0);  // the pair read by lanes past their count
  for (unsigned it = 0;; ++it)
  {
    const long long first = (long long)t * NT;
    const int nc_t = (int)((P.n_bulk - first) < NT ? (P.n_bulk - first) : NT);
    int nid = -1, fid = -1;
    if (has_next && tid < hn.nn) nid = __ldg(P.node_ids + hn.node_off + tid);
    if (has_next && NT - 1 - tid < hvn.nd) fid = __ldg(Q.dest_k + hvn.dest_off + (NT - 1 - tid));  // row of this thread's record

    // phase 1: thread = cell
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
      {
        double F = 0.0;
#pragma unroll
        for (int v = 0; v < NV; ++v) F += S.fs[S.crow[tid * NV + v]];
        const double sc = cL * G.vol;
        S.cellv[tid] = make_double2(sc, sc * F);
      }
      double w[NV], Ae[NV][NV];
      if (ina.kernel == MPCX_KERNEL_LAPLACE_VARCOEF)
      {
        const long long index = __ldg(P.cell_pos + first + tid);
        p1_load_w<TD>(ina, index, ina.cells ? __ldg(ina.cells + index) : (int)index, w);
      }
      p1_element<TD>(ina.kernel, G, ina.c, w, Ae);
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
    if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // reductions of t-1 have read the staging buffer
    __syncthreads();  // 1: element buffers complete; cell records, Xs and the staging buffer are free
    if (tid == 0 && has_next) fused_tma_cells(S, P, Q, tn, NV, NS);
    for (int i = tid; i < (h.stage >> 1); i += NT) reinterpret_cast<double2*>(S.stage)[i] = make_double2(0.0, 0.0);
    __syncthreads();  // 1b: staging buffer zeroed

    // phase 2: thread = record
    mbar_wait(S.barR, it & 1);
    for (int k = tid; k < h.nd; k += NT)
    {
      const double v = tile_record_sum(S.ebuf, S.R, k);
      S.stage[S.R.spos[k]] = v;
      if (SYM) S.stage[S.R.spos2[k]] = v;
    }
    // row records go to the LAST threads (the first ones hold the heaviest matrix records: those are ordered by
    // descending source count); thread NT-1-k carries f of row k in a register (fcur)
    for (int k = NT - 1 - tid; k < hv.nd; k += NT)
    {
      const unsigned g = S.giv[k >> 5];
      const uint16_t* e = S.vinc + (g & 0xffffu) + (k & 31);
      const int cnt = S.dcntv[k];
      double s1 = 0.0, s2 = 0.0;
#pragma unroll 1
      for (int i = 0; i < cnt; i += 4, e += 4 * MPCX_CT_GSTRIDE)
      {
        // four independent index -> pair loads in flight; lanes past their count read the zero pair at cellv[NT]
        const int c0 = e[0], c1 = i + 1 < cnt ? (int)e[MPCX_CT_GSTRIDE] : NT, c2 = i + 2 < cnt ? (int)e[2 * MPCX_CT_GSTRIDE] : NT,
                  c3 = i + 3 < cnt ? (int)e[3 * MPCX_CT_GSTRIDE] : NT;
        const double2 v0 = S.cellv[c0], v1 = S.cellv[c1], v2 = S.cellv[c2], v3 = S.cellv[c3];
        s1 += (v0.x + v1.x) + (v2.x + v3.x);
        s2 += (v0.y + v1.y) + (v2.y + v3.y);
      }
      const int row = S.dkv[k];
      atomicAdd(b + row, (k < NT ? fcur : __ldg(inL.wnodal + row)) * s1 + s2);
    }
    if (nid >= 0) { S.Xs[3 * tid] = xg0; S.Xs[3 * tid + 1] = xg1; S.Xs[3 * tid + 2] = xg2; }
    if (fid >= 0) S.fs[NT - 1 - tid] = fg;
    if (has_next)
    {
      for (int i = tid + NT; i < hn.nn; i += NT)
        load_vertex(mesh, __ldg(P.node_ids + hn.node_off + i), S.Xs[3 * i], S.Xs[3 * i + 1], S.Xs[3 * i + 2]);
      for (int k = tid + NT; k < hvn.nd; k += NT) S.fs[k] = __ldg(inL.wnodal + __ldg(Q.dest_k + hvn.dest_off + k));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // staging writes -> visible to the copy engine
    __syncthreads();  // 2: staging buffer, next Xs / fs complete; element buffers and records free

    if (issuer)
    {
      for (int r = tid; r < h.nr; r += 32)
      {
        const int2 rr = S.R.runs[r];
        tma_reduce_add_f64(A.val + rr.x, S.stage + (rr.y & 0xffff), (unsigned)(rr.y >> 16) * 8u);
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      __syncwarp();  // every lane has read its runs: the record buffers may be refilled
      if (tid == 0 && has_next) { tma_records<false, SYM>(S.R, P, hn); fused_tma_vrecords(S, Q, hvn); }
    }
    if (!has_next) break;
    h = hn; hv = hvn; t = tn; tn += gridDim.x; has_next = tn < nt; fcur = fg;
    if (has_next) { hn = load_hdr(P.hdr, tn); hvn = load_vhdr(Q.hdr, tn); }
  }
  if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the staging buffer must outlive the reads
}
}  // namespace
