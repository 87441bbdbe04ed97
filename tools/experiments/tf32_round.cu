// Experiment: is (u + 0x1000) & 0xFFFFE000 the same as cvt.rna.tf32.f32 for every finite float32 (subnormals included)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tf32_round tf32_round.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(unsigned long long* bad, unsigned* first) {
  unsigned long long n = 0;
  for (unsigned long long u = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; u < (1ull << 32); u += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t x = (uint32_t)u;
    if ((x & 0x7F800000u) == 0x7F800000u) continue;                 // inf / NaN: not compared
    uint32_t a;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(a) : "f"(__uint_as_float(x)));
    const uint32_t b = (x + 0x1000u) & 0xFFFFE000u;
    if (a != b) { if (n == 0) atomicMin(first, x); ++n; }
  }
  if (n) atomicAdd(bad, n);
}
int main() {
  unsigned long long* d; unsigned* f; cudaMalloc(&d, 8); cudaMalloc(&f, 4);
  cudaMemset(d, 0, 8); cudaMemset(f, 0xFF, 4);
  k<<<148 * 8, 256>>>(d, f);
  unsigned long long h; unsigned hf;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&hf, f, 4, cudaMemcpyDeviceToHost);
  printf("finite float32 bit patterns where the two roundings differ: %llu (first 0x%08x) -- %s\n", h, hf, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
