// Experiment: cycles per tcgen05.mma.kind::tf32 (M=128, K=8, SS operands) as a function of N,
// issued back-to-back by one thread on fixed shared-memory operands.
#include <cstdio>
#include <vector>
#include "../../davo_b200/csrc/ptx.cuh"
using namespace davo;

template <int N, int M>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters, int sbo_bytes, int distinct) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + 160 * 1024);
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) ((float*)smem)[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); fence_mbar_init(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) tmem_alloc(slot, 256);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = umma_idesc_tf32(M, N);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      // `distinct` different A windows, 4 K-slices each (as the conv kernel issues them)
      const uint32_t a = a0 + (uint32_t)(i % distinct) * 128 * 3;
      const uint64_t da = (uint64_t)((a & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | (2ull << 61);
      const uint64_t db = umma_desc_sw128(b0);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) tc_mma_tf32(tm, da + 2 * kk, db + 2 * kk, idesc, 1);
    }
    long long t1 = clock64();
    tc_commit(&bar[0]);
    mbar_wait(&bar[0], 0);
    long long t2 = clock64();
    out[0] = t1 - t0; out[1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 256); }
}

template <int N, int M> void run(long long* d, int sbo) {
  const int smem = 162 * 1024 + 1024;
  cudaFuncSetAttribute((k<N, M>), cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long h[2];
  for (int iters : {1024}) {
    k<N, M><<<1, 128, smem>>>(d, iters, sbo, 28);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); return; }
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("M=%3d N=%3d sbo=%4d iters=%4d (x4 MMAs): issue %.1f cyc/MMA, complete %.1f cyc/MMA\n", M, N, sbo, iters,
           h[0] / (4.0 * iters), h[1] / (4.0 * iters));
  }
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  for (int sbo : {1024}) {
    run<16, 128>(d, sbo); run<64, 128>(d, sbo); run<128, 128>(d, sbo); run<192, 128>(d, sbo); run<256, 128>(d, sbo);
    run<16, 64>(d, sbo); run<64, 64>(d, sbo); run<128, 64>(d, sbo); run<256, 64>(d, sbo);
  }
  return 0;
}
