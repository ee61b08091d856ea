// mpcx_pattern_gpu.cuh -- sparsity pattern with the MPC additions, built on the device.
//
// Same result as mpcx_create_pattern_host (the CSR the reference gets from create_sparsity_pattern,
// cpp/utils.h:381-496: for every owned cell c, (rows(c) U row-masters(c)) x (cols(c) U col-masters(c)) at block
// level, expanded by bs0 x bs1), but where the reference inserts cell by cell into a dolfinx::la::SparsityPattern
// (hash/sort per row on the host) this emits one 64-bit key (block row, block col) per coupling, radix-sorts the
// keys with CUB, removes duplicates and reads the CSR off the sorted list.  At 256^3 P1 that is 1.6 G keys and a
// few hundred milliseconds instead of seconds on the host plus a 3 GB upload.
#pragma once
#include <cub/cub.cuh>

namespace
{
struct PatternSide
{
  const int* dofmap;
  int nd, bs;
  const int* masters;
  const int* offsets;
  const int* c2s;
  const int* c2s_off;
};

struct Pattern
{
  long long nbr = 0, nnz_block = 0;
  int bs0 = 1, bs1 = 1, colbits = 0;
  unsigned long long* keys = nullptr;  // sorted unique (block row << colbits) | block col
  long long* start = nullptr;          // [nbr + 1] first key of every block row
};

__device__ __forceinline__ int side_extra(const PatternSide& s, long long c)
{
  int m = 0;
  if (s.c2s_off)
    for (int k = s.c2s_off[c]; k < s.c2s_off[c + 1]; ++k)
    {
      const int sl = s.c2s[k];
      m += s.offsets[sl + 1] - s.offsets[sl];
    }
  return m;
}
// i-th block of the cell on this side: its dofs, then the master blocks of its slaves
__device__ __forceinline__ int side_block(const PatternSide& s, long long c, int i)
{
  if (i < s.nd) return s.dofmap[c * s.nd + i];
  i -= s.nd;
  for (int k = s.c2s_off[c]; k < s.c2s_off[c + 1]; ++k)
  {
    const int sl = s.c2s[k], n = s.offsets[sl + 1] - s.offsets[sl];
    if (i < n) return s.masters[s.offsets[sl] + i] / s.bs;
    i -= n;
  }
  return 0;
}

__global__ void k_pat_count(PatternSide r, PatternSide c, long long nc, long long* __restrict__ cnt)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  cnt[i] = (long long)(r.nd + side_extra(r, i)) * (c.nd + side_extra(c, i));
}

__global__ void k_pat_fill(PatternSide r, PatternSide c, long long nc, const long long* __restrict__ off, int colbits,
                           unsigned long long* __restrict__ keys)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  const int n0 = r.nd + side_extra(r, i), n1 = c.nd + side_extra(c, i);
  unsigned long long* out = keys + off[i];
  if (n1 == c.nd && n1 <= 32)
  {
    int cb[32];  // the common case: no column masters
    for (int q = 0; q < n1; ++q) cb[q] = c.dofmap[i * c.nd + q];
    for (int p = 0; p < n0; ++p)
    {
      const unsigned long long rb = (unsigned long long)side_block(r, i, p) << colbits;
      for (int q = 0; q < n1; ++q) out[p * n1 + q] = rb | (unsigned)cb[q];
    }
    return;
  }
  for (int p = 0; p < n0; ++p)
  {
    const unsigned long long rb = (unsigned long long)side_block(r, i, p) << colbits;
    for (int q = 0; q < n1; ++q) out[(long long)p * n1 + q] = rb | (unsigned)side_block(c, i, q);
  }
}

__global__ void k_pat_row_start(const unsigned long long* __restrict__ keys, long long n, long long nbr, int colbits,
                                long long* __restrict__ start)
{
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > nbr) return;
  const unsigned long long target = (unsigned long long)r << colbits;
  long long lo = 0, hi = n;
  while (lo < hi)
  {
    const long long mid = (lo + hi) >> 1;
    if (keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  start[r] = lo;
}

__global__ void k_pat_export_rows(const long long* __restrict__ start, long long nbr, int bs0, int bs1,
                                  long long* __restrict__ row_ptr)
{
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // scalar row, or nbr * bs0 for the end
  if (i > nbr * bs0) return;
  if (i == nbr * bs0) { row_ptr[i] = start[nbr] * bs0 * bs1; return; }
  const long long r = i / bs0, a = i - r * bs0, len = start[r + 1] - start[r];
  row_ptr[i] = (start[r] * bs0 + a * len) * bs1;
}

__global__ void k_pat_export_cols(const unsigned long long* __restrict__ keys, const long long* __restrict__ start,
                                  long long nnz_block, int bs0, int bs1, int colbits, int* __restrict__ col)
{
  // one thread per (block entry, a): writes bs1 consecutive scalar columns
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nnz_block * bs0) return;
  const long long k = t / bs0;
  const int a = (int)(t - k * bs0);
  const unsigned long long key = keys[k];
  const long long r = (long long)(key >> colbits);
  const int cb = (int)(key & ((1ull << colbits) - 1));
  const long long len = start[r + 1] - start[r];
  int* dst = col + ((start[r] * bs0 + a * len) + (k - start[r])) * bs1;
  for (int b = 0; b < bs1; ++b) dst[b] = cb * bs1 + b;
}

void pattern_free(Pattern* P)
{
  if (!P) return;
  cudaFree(P->keys); cudaFree(P->start);
  delete P;
}

#define PT_CK(call)                                                      \
  do {                                                                   \
    cudaError_t e__ = (call);                                            \
    if (e__ != cudaSuccess) { rc = cuda_check(e__, #call); goto done; }  \
  } while (0)

int pattern_build(const mpcx_dofmap* d0, const mpcx_dofmap* d1, long long nc, long long nbr, long long nbc,
                  const mpcx_mpc* m0, const mpcx_mpc* m1, cudaStream_t s, Pattern** out)
{
  int rc = MPCX_OK;
  Pattern* P = new Pattern();
  const PatternSide R{d0->map, d0->nd, d0->bs, m0 ? m0->masters : nullptr, m0 ? m0->offsets : nullptr,
                      m0 ? m0->cell_to_slaves : nullptr, m0 ? m0->cell_to_slaves_offsets : nullptr};
  const PatternSide Cc{d1->map, d1->nd, d1->bs, m1 ? m1->masters : nullptr, m1 ? m1->offsets : nullptr,
                       m1 ? m1->cell_to_slaves : nullptr, m1 ? m1->cell_to_slaves_offsets : nullptr};
  long long *cnt = nullptr, *off = nullptr, *nsel = nullptr, total = 0, last_cnt = 0, last_off = 0;
  unsigned long long *keys = nullptr, *keys2 = nullptr;
  void* tmp = nullptr;
  size_t tb = 0, tb2 = 0;
  int rowbits = 1;
  const unsigned nb = (unsigned)((nc + 255) / 256 > 0 ? (nc + 255) / 256 : 1);
  P->nbr = nbr; P->bs0 = d0->bs; P->bs1 = d1->bs;
  while ((1ll << P->colbits) < (nbc > 1 ? nbc : 2)) ++P->colbits;
  while ((1ll << rowbits) < (nbr > 1 ? nbr : 2)) ++rowbits;
  if (P->colbits > 31 || rowbits + P->colbits > 64) { rc = fail(MPCX_ERR_UNSUPPORTED, "pattern: too many rows / columns"); goto done; }

  PT_CK(cudaMalloc(&P->start, sizeof(long long) * (size_t)(nbr + 1)));
  if (nc > 0)
  {
    PT_CK(cudaMalloc(&cnt, sizeof(long long) * (size_t)nc));
    PT_CK(cudaMalloc(&off, sizeof(long long) * (size_t)nc));
    k_pat_count<<<nb, 256, 0, s>>>(R, Cc, nc, cnt);
    PT_CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt, off, nc, s));
    PT_CK(cudaMalloc(&tmp, tb));
    PT_CK(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, off, nc, s));
    PT_CK(cudaMemcpyAsync(&last_cnt, cnt + nc - 1, sizeof(long long), cudaMemcpyDeviceToHost, s));
    PT_CK(cudaMemcpyAsync(&last_off, off + nc - 1, sizeof(long long), cudaMemcpyDeviceToHost, s));
    PT_CK(cudaStreamSynchronize(s));
    total = last_cnt + last_off;
    cudaFree(tmp); tmp = nullptr;
  }
  if (total >= (1ll << 31) - 1)
  {
    // CUB's device-wide primitives of this toolkit count items in 32 bits
    rc = fail(MPCX_ERR_UNSUPPORTED, "pattern: more than 2^31 cell couplings on one device (use the host builder)");
    goto done;
  }
  if (total > 0)
  {
    PT_CK(cudaMalloc(&keys, sizeof(unsigned long long) * (size_t)total));
    PT_CK(cudaMalloc(&keys2, sizeof(unsigned long long) * (size_t)total));
    PT_CK(cudaMalloc(&nsel, sizeof(long long)));
    k_pat_fill<<<nb, 256, 0, s>>>(R, Cc, nc, off, P->colbits, keys);
    cudaFree(cnt); cnt = nullptr; cudaFree(off); off = nullptr;
    PT_CK(cub::DeviceRadixSort::SortKeys(nullptr, tb, keys, keys2, (int)total, 0, rowbits + P->colbits, s));
    PT_CK(cub::DeviceSelect::Unique(nullptr, tb2, keys2, keys, nsel, (int)total, s));
    tb = tb > tb2 ? tb : tb2;
    PT_CK(cudaMalloc(&tmp, tb));
    PT_CK(cub::DeviceRadixSort::SortKeys(tmp, tb, keys, keys2, (int)total, 0, rowbits + P->colbits, s));
    PT_CK(cub::DeviceSelect::Unique(tmp, tb, keys2, keys, nsel, (int)total, s));
    PT_CK(cudaMemcpyAsync(&P->nnz_block, nsel, sizeof(long long), cudaMemcpyDeviceToHost, s));
    PT_CK(cudaStreamSynchronize(s));
    cudaFree(keys2); keys2 = nullptr;
    // keep only the unique keys
    PT_CK(cudaMalloc(&P->keys, sizeof(unsigned long long) * (size_t)(P->nnz_block > 0 ? P->nnz_block : 1)));
    PT_CK(cudaMemcpyAsync(P->keys, keys, sizeof(unsigned long long) * (size_t)P->nnz_block, cudaMemcpyDeviceToDevice, s));
  }
  k_pat_row_start<<<(unsigned)((nbr + 256) / 256), 256, 0, s>>>(P->keys, P->nnz_block, nbr, P->colbits, P->start);
  PT_CK(cudaStreamSynchronize(s));

done:
  cudaFree(cnt); cudaFree(off); cudaFree(keys); cudaFree(keys2); cudaFree(tmp); cudaFree(nsel);
  if (rc != MPCX_OK) { pattern_free(P); P = nullptr; }
  *out = P;
  return rc;
}
}  // namespace
