// mpcx_pattern.cpp -- host-side sparsity pattern with the MPC additions (cold path).
//
// Produces the CSR the reference gets from create_sparsity_pattern (cpp/utils.h:381-496):
// for every owned cell c,  (rows(c) U row-masters(c)) x (cols(c) U col-masters(c))  at block
// level, expanded by bs0 x bs1.  The reference inserts cell by cell into a
// dolfinx::la::SparsityPattern; here the pattern is built row-wise (block row -> incident
// cells -> sorted unique column candidates) on several host threads.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "mpcx.h"

namespace
{
struct Side
{
  const int32_t* dofmap;
  int32_t nd, bs;
  const mpcx_mpc_host* mpc;
  // block ids the cell touches on this side: its dofs, then the masters of its slaves
  template <typename F>
  void for_each_block(int64_t c, F&& f) const
  {
    for (int i = 0; i < nd; ++i) f(dofmap[c * nd + i]);
    if (!mpc || !mpc->cell_to_slaves_offsets) return;
    for (int32_t k = mpc->cell_to_slaves_offsets[c]; k < mpc->cell_to_slaves_offsets[c + 1]; ++k)
    {
      const int32_t s = mpc->cell_to_slaves[k];
      for (int32_t q = mpc->offsets[s]; q < mpc->offsets[s + 1]; ++q) f(mpc->masters[q] / bs);
    }
  }
};
}  // namespace

extern "C" int mpcx_create_pattern_host(const int32_t* dofmap0, int32_t nd0, int32_t bs0,
                                        const int32_t* dofmap1, int32_t nd1, int32_t bs1,
                                        int64_t num_cells, int64_t num_block_rows,
                                        const mpcx_mpc_host* mpc0, const mpcx_mpc_host* mpc1,
                                        int32_t num_threads, int64_t** row_ptr_out,
                                        int32_t** col_out, int64_t* nnz_out)
{
  if (!dofmap0 || !dofmap1 || !row_ptr_out || !col_out || !nnz_out || num_cells < 0 || num_block_rows < 0)
    return MPCX_ERR_ARG;
  const Side rows{dofmap0, nd0, bs0, mpc0}, cols{dofmap1, nd1, bs1, mpc1};
  const int64_t nbr = num_block_rows;

  // block row -> incident cells
  std::vector<int64_t> adj_ptr(nbr + 1, 0);
  for (int64_t c = 0; c < num_cells; ++c)
    rows.for_each_block(c, [&](int32_t r) { ++adj_ptr[r + 1]; });
  for (int64_t r = 0; r < nbr; ++r) adj_ptr[r + 1] += adj_ptr[r];
  std::vector<int32_t> adj(adj_ptr[nbr]);
  {
    std::vector<int64_t> fill(adj_ptr.begin(), adj_ptr.end() - 1);
    for (int64_t c = 0; c < num_cells; ++c)
      rows.for_each_block(c, [&](int32_t r) { adj[fill[r]++] = (int32_t)c; });
  }

  int nt = num_threads > 0 ? num_threads : (int)std::thread::hardware_concurrency();
  if (nt < 1) nt = 1;
  if ((int64_t)nt > nbr) nt = nbr > 0 ? (int)nbr : 1;
  std::vector<std::vector<int32_t>> tcols(nt);
  std::vector<int64_t> brow_len(nbr, 0);
  auto work = [&](int t) {
    const int64_t r0 = nbr * t / nt, r1 = nbr * (t + 1) / nt;
    std::vector<int32_t> cand;
    auto& out = tcols[t];
    for (int64_t r = r0; r < r1; ++r)
    {
      cand.clear();
      for (int64_t k = adj_ptr[r]; k < adj_ptr[r + 1]; ++k)
        cols.for_each_block(adj[k], [&](int32_t b) { cand.push_back(b); });
      std::sort(cand.begin(), cand.end());
      cand.erase(std::unique(cand.begin(), cand.end()), cand.end());
      brow_len[r] = (int64_t)cand.size();
      out.insert(out.end(), cand.begin(), cand.end());
    }
  };
  {
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& x : th) x.join();
  }

  // expand block pattern to the scalar CSR
  const int64_t nrows = nbr * bs0;
  int64_t* row_ptr = (int64_t*)std::malloc(sizeof(int64_t) * (size_t)(nrows + 1));
  if (!row_ptr) return MPCX_ERR_ALLOC;
  row_ptr[0] = 0;
  for (int64_t r = 0; r < nbr; ++r)
    for (int a = 0; a < bs0; ++a) row_ptr[r * bs0 + a + 1] = row_ptr[r * bs0 + a] + brow_len[r] * bs1;
  const int64_t nnz = row_ptr[nrows];
  int32_t* col = (int32_t*)std::malloc(sizeof(int32_t) * (size_t)(nnz > 0 ? nnz : 1));
  if (!col) { std::free(row_ptr); return MPCX_ERR_ALLOC; }
  auto expand = [&](int t) {
    const int64_t r0 = nbr * t / nt, r1 = nbr * (t + 1) / nt;
    const int32_t* src = tcols[t].data();
    for (int64_t r = r0; r < r1; ++r)
    {
      const int64_t len = brow_len[r];
      for (int a = 0; a < bs0; ++a)
      {
        int32_t* dst = col + row_ptr[r * bs0 + a];
        for (int64_t k = 0; k < len; ++k)
          for (int b = 0; b < bs1; ++b) dst[k * bs1 + b] = src[k] * bs1 + b;
      }
      src += len;
    }
  };
  {
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(expand, t);
    expand(0);
    for (auto& x : th) x.join();
  }
  *row_ptr_out = row_ptr;
  *col_out = col;
  *nnz_out = nnz;
  return MPCX_OK;
}

extern "C" void mpcx_free_host(void* p) { std::free(p); }
