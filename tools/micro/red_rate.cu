// Microbenchmark: issue cost of red.global.add.f64 as a function of the address pattern of a warp instruction.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o red_rate red_rate.cu && ./red_rate
#include <cstdio>
#include <cuda_runtime.h>

// pattern 0: 32 lanes -> 32 lines (stride 128 B);  1: 32 consecutive doubles (256 B);  2: groups of 3 consecutive
// doubles, groups scattered;  3: 16 lanes x 2 consecutive;  4: groups of 4 consecutive (32 B sector each)
__global__ void k(double* p, long long n, int pattern, int iters)
{
  const int lane = threadIdx.x & 31;
  long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  unsigned long long h = warp * 2654435761ull + 12345;
  for (int it = 0; it < iters; ++it)
  {
    h = h * 6364136223846793005ull + 1442695040888963407ull;
    long long base = (long long)((h >> 20) % (unsigned long long)(n - 8192));
    base &= ~15ll;
    long long idx;
    if (pattern == 0) idx = base + lane * 16;
    else if (pattern == 1) idx = base + lane;
    else if (pattern == 2) idx = base + (lane / 3) * 48 + lane % 3;
    else if (pattern == 3) idx = base + (lane / 2) * 32 + lane % 2;
    else idx = base + (lane / 4) * 64 + lane % 4;
    atomicAdd(p + idx, 1.0);
  }
}

int main()
{
  const long long n = 1ll << 28;  // 2 GB of doubles
  double* p;
  cudaMalloc(&p, n * 8);
  cudaMemset(p, 0, n * 8);
  const int iters = 256, blocks = 148 * 8, threads = 256;
  for (int pattern = 0; pattern < 5; ++pattern)
  {
    k<<<blocks, threads>>>(p, n, pattern, 8);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<<<blocks, threads>>>(p, n, pattern, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double lanes = (double)blocks * threads * iters;
    printf("pattern %d: %.3f ms, %.2f G lane-REDs/s, %.3f SM-cycles per lane (1965 MHz, 148 SMs)\n", pattern, ms,
           lanes / ms / 1e6, ms * 1e-3 * 1.965e9 * 148 / lanes);
  }
  return 0;
}
