// Experiment: the hardware floor of tcgen05.mma.kind::tf32 (M=128, K=8, SS) per N when the issuing
// thread does nothing else: descriptors precomputed, 16 MMAs unrolled per loop trip.
#include <cstdio>
#include "../../davo_b200/csrc/ptx.cuh"
using namespace davo;

template <int N>
__global__ void __launch_bounds__(128, 1) k(long long* out, int iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + 160 * 1024);
  uint32_t* slot = (uint32_t*)(bar + 4);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) ((float*)smem)[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); fence_mbar_init(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) tmem_alloc(slot, 256);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    constexpr uint32_t idesc = umma_idesc_tf32(128, N);
    const uint64_t da = umma_desc_sw128(smem_u32(smem)), db = umma_desc_sw128(smem_u32(smem + 96 * 1024));
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        asm volatile("tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, 1;" ::"r"(tm), "l"(da + 2 * (j & 3) + 64 * (j >> 2)),
                     "l"(db + 2 * (j & 3)), "r"(idesc) : "memory");
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

template <int N> void run(long long* d) {
  const int smem = 162 * 1024 + 1024;
  cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) {
    const int iters = 256;
    k<N><<<1, 128, smem>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); return; }
    long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    if (rep) printf("N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA\n", N, h[0] / (16.0 * iters), h[1] / (16.0 * iters));
  }
}
int main() {
  long long* d; cudaMalloc(&d, 16);
  run<16>(d); run<32>(d); run<64>(d); run<96>(d); run<128>(d); run<192>(d); run<256>(d);
  return 0;
}
