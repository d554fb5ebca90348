// Micro-benchmark: how many random 8-byte gathers per second can a B200 sustain from an
// L2-resident vector?  (development aid for the SpMV design; not part of the product)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int U, typename TT>
__global__ void gather_kernel(const int* __restrict__ idx, const TT* __restrict__ x, TT* out, long long n) {
  long long i = (long long)blockIdx.x * blockDim.x * U + threadIdx.x;
  TT acc = 0;
  int c[U];
#pragma unroll
  for (int u = 0; u < U; u++) { long long k = i + (long long)u * blockDim.x; c[u] = k < n ? idx[k] : 0; }
#pragma unroll
  for (int u = 0; u < U; u++) acc += x[c[u]];
  if (acc == (TT)123456789) out[0] = acc;
}

template <int U, typename TT>
void run(const char* name, int region, long long n, int block) {
  std::vector<int> h(n);
  for (long long i = 0; i < n; i++) h[i] = rand() % region;
  int* d_idx; TT* d_x; TT* d_out;
  cudaMalloc(&d_idx, n * 4); cudaMalloc(&d_x, (size_t)region * sizeof(TT)); cudaMalloc(&d_out, 64);
  cudaMemcpy(d_idx, h.data(), n * 4, cudaMemcpyHostToDevice);
  cudaMemset(d_x, 0, (size_t)region * sizeof(TT));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int grid = (int)((n + (long long)block * U - 1) / ((long long)block * U));
  for (int w = 0; w < 3; w++) gather_kernel<U, TT><<<grid, block>>>(d_idx, d_x, d_out, n);
  cudaEventRecord(e0);
  const int reps = 10;
  for (int r = 0; r < reps; r++) gather_kernel<U, TT><<<grid, block>>>(d_idx, d_x, d_out, n);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  printf("%-28s U=%d block=%d region=%8d (%6.1f MB): %7.1f us  %6.1f Ggather/s  idx-stream %5.0f GB/s\n", name, U, block,
         region, region * sizeof(TT) / 1e6, ms * 1e3, n / ms / 1e6, n * 4 / ms / 1e6);
  cudaFree(d_idx); cudaFree(d_x); cudaFree(d_out);
}

int main() {
  const long long n = 11400000;
  for (int region : {100000, 1000000, 10000000}) {
    run<4, double>("f64 gather", region, n, 256);
    run<8, double>("f64 gather", region, n, 256);
    run<8, double>("f64 gather", region, n, 512);
    run<8, float>("f32 gather", region, n, 256);
  }
  printf("err %d\n", (int)cudaGetLastError());
  return 0;
}
