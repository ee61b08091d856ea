// Microbenchmark: throughput of the copy engine's bulk operations shared -> global as the fused tile kernel uses them:
//   op 0: cp.reduce.async.bulk ... add.f64 (SASS UBLKRED)     op 1: cp.async.bulk.global.shared::cta (plain store)
// Persistent CTAs of 448 threads, 2 per SM, each with a 25.6 KB staging buffer; per "tile" every CTA issues R
// operations of L bytes each (operation r from thread r, as in the kernel) to 16-byte aligned places of a 2 GB array
// (near: consecutive segments of a 64 KB window per tile, like the CSR rows of a tile; far: anywhere), then waits for the
// reads of the staging buffer before the next tile.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o bulk_rate bulk_rate.cu && ./bulk_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int OP>
__global__ void __launch_bounds__(448, 2) k(double* p, long long n, int R, int L8, int tiles, int far_)
{
  extern __shared__ __align__(16) double stage[];
  for (int i = threadIdx.x; i < 3200; i += blockDim.x) stage[i] = 1.0;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  unsigned long long h = (blockIdx.x + 1) * 2654435761ull;
  for (int t = 0; t < tiles; ++t)
  {
    h = h * 6364136223846793005ull + 1442695040888963407ull;
    const long long window = (long long)((h >> 20) % (unsigned long long)(n - (1 << 20))) & ~1ll;
    const int r = threadIdx.x;
    if (r < R)
    {
      unsigned long long h2 = h + r * 0x9E3779B97F4A7C15ull;
      h2 ^= h2 >> 29; h2 *= 0xBF58476D1CE4E5B9ull; h2 ^= h2 >> 32;
      long long off = far_ ? (long long)(h2 % (unsigned long long)(n - 8192)) : window + (long long)r * (8192 / (R > 0 ? R : 1) + L8);
      off &= ~1ll;
      const unsigned src = smem_u32(stage + ((r * L8) % (3200 - L8) & ~1));
      if (OP == 0)
        asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(p + off), "r"(src), "r"(L8 * 8) : "memory");
      else
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + off), "r"(src), "r"(L8 * 8) : "memory");
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main()
{
  const long long n = 1ll << 28;  // 2 GB of doubles
  double* p;
  cudaMalloc(&p, n * 8);
  cudaMemset(p, 0, n * 8);
  const int blocks = 148 * 2, threads = 448, tiles = 400;
  const size_t smem = 3200 * 8 + 80 * 1024;  // + padding so that exactly 2 CTAs fit an SM, as in the kernel
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int Rs[] = {8, 16, 33, 66, 132, 264, 448};
  for (int op = 0; op < 2; ++op)
    for (int far_ = 0; far_ < 2; ++far_)
      for (int ri = 0; ri < 7; ++ri)
      {
        const int R = Rs[ri];
        int L8 = 2800 / R;  // ~2800 doubles per tile in total, as the kernel's tiles
        L8 &= ~1;
        if (L8 < 2) L8 = 2;
        for (int rep = 0; rep < 2; ++rep)
        {
          cudaEvent_t a, b;
          cudaEventCreate(&a); cudaEventCreate(&b);
          cudaEventRecord(a);
          if (op == 0) k<0><<<blocks, threads, smem>>>(p, n, R, L8, tiles, far_);
          else k<1><<<blocks, threads, smem>>>(p, n, R, L8, tiles, far_);
          cudaEventRecord(b);
          cudaEventSynchronize(b);
          float ms;
          cudaEventElapsedTime(&ms, a, b);
          if (rep == 1)
          {
            const double ops = (double)blocks * tiles * R, elems = ops * L8;
            printf("%s %s R=%3d ops/tile x %4d B: %7.3f ms  %6.2f G ops/s  %7.1f G f64/s  %6.2f TB/s  %6.0f cycles/tile/CTA\n",
                   op == 0 ? "reduce" : "store ", far_ ? "far " : "near", R, L8 * 8, ms, ops / ms / 1e6, elems / ms / 1e6,
                   elems * 8 / ms / 1e9, ms * 1e-3 * 1.965e9 / tiles);
          }
        }
      }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
