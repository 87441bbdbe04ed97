// Experiment: does cta_group::2 (M=256 over a CTA pair, each CTA supplying half of the B
// operand) raise the POWER-CAPPED tensor throughput of kind::tf32 over cta_group::1 (M=128 per CTA, full B)?
// Every SM issues back-to-back MMAs on fixed shared-memory operands for a few seconds; no TMA, no epilogue:
// what differs between the two modes is only the shared-memory operand traffic per FLOP
// (cta_group::1: 4 KB of A + 8 KB of B per 128x256x8; cta_group::2: 2 x 4 KB of A + 2 x 4 KB of B per 256x256x8).
// Prints TFLOP/s over the whole run; run nvidia-smi -lms beside it for clocks and power.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_power mma_power.cu && ./mma_power [seconds]
#include <cstdio>
#include <cstdlib>
#include "../../davo_b200/csrc/ptx.cuh"
using namespace davo;

__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_2cta(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2cta(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)), "h"(mask) : "memory");
}

// MODE 1: cta_group::1, M=128, N=NN per CTA.  MODE 2: cta_group::2, M=256, N=NN per CTA pair.
template <int MODE, int NN>
__global__ void __launch_bounds__(128, 1) k(long long* out, long long iters, int random_data) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + 160 * 1024);
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = MODE == 2 ? cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) {
    // low-entropy operands (eight distinct values) or hashed pseudo-random ones in (-1, 1), TF32-rounded like the
    // conv stack's activations and weights: the tensor core's switching power depends on the data
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const float r = (float)(int)(h >> 8) * (1.0f / 8388608.0f) - 1.0f;
    ((float*)smem)[i] = random_data ? round_tf32(r) : 1.0f + (float)(i & 7) * 0.125f;
  }
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); fence_mbar_init(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) { if (MODE == 2) tmem_alloc2(slot, 256); else tmem_alloc(slot, 256); }
  tc_fence_before(); __syncthreads();
  if (MODE == 2) cluster_sync();
  tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = umma_idesc_tf32(MODE == 2 ? 256 : 128, NN);
    // eight A slabs (16 KB each) and three B slabs (32 KB / 16 KB each) so that the operand addresses move as in the conv
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
    long long t0 = clock64();
    for (long long i = 0; i < iters; ++i) {
      const uint64_t da = umma_desc_sw128(a0 + (uint32_t)(i & 3) * 16384);
      const uint64_t db = umma_desc_sw128(b0 + (uint32_t)(i % 3) * 32768);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (MODE == 2) tc_mma_tf32_2cta(tm, da + 2 * kk, db + 2 * kk, idesc, 1);
        else tc_mma_tf32(tm, da + 2 * kk, db + 2 * kk, idesc, 1);
      }
    }
    if (MODE == 2) tc_commit_2cta(&bar[0], 1); else tc_commit(&bar[0]);
    mbar_wait(&bar[0], 0);
    if (blockIdx.x == 0) out[0] = clock64() - t0;
  }
  tc_fence_before(); __syncthreads();
  if (MODE == 2) cluster_sync();
  if (warp == 0) { tc_fence_after(); if (MODE == 2) tmem_dealloc2(tm, 256); else tmem_dealloc(tm, 256); }
}

template <int MODE, int NN>
void run(long long* d, double seconds, int sms, int random_data) {
  auto* kern = k<MODE, NN>;
  const int smem = 162 * 1024 + 1024;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = MODE == 2 ? 2 : 1; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.gridDim = dim3(sms & ~1); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem; cfg.attrs = &attr; cfg.numAttrs = 1;
  const double flop_per_iter = 4.0 * 2.0 * (MODE == 2 ? 256 : 128) * NN * 8;     // per issuing thread
  const int issuers = MODE == 2 ? (sms & ~1) / 2 : (sms & ~1);
  long long iters = 20000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int pass = 0; pass < 2; ++pass) {        // pass 0 calibrates the iteration count
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, d, iters, random_data);
    cudaEventRecord(e1);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d N=%d: %s\n", MODE, NN, cudaGetErrorString(e)); return; }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long cyc = 0;
    cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    if (pass == 1)
      printf("%s data, cta_group::%d M=%d N=%d on %d SMs: %.2f s, %.1f TFLOP/s, %.1f cycles per (MMA x %d SM), mean SM clock %.0f MHz\n", random_data ? "random" : "constant", MODE,
             MODE == 2 ? 256 : 128, NN, sms & ~1, ms * 1e-3, flop_per_iter * iters * issuers / (ms * 1e-3) / 1e12,
             (double)cyc / (4.0 * iters), MODE, cyc / (ms * 1e-3) / 1e6);

    iters = (long long)(iters * seconds / (ms * 1e-3));
  }
  fflush(stdout);
}

int main(int argc, char** argv) {
  const double seconds = argc > 1 ? atof(argv[1]) : 4.0;
  long long* d; cudaMalloc(&d, 16);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  for (int random_data = 0; random_data < 2; ++random_data)
    for (int rep = 0; rep < 2; ++rep) {
      run<1, 256>(d, seconds, p.multiProcessorCount, random_data);
      run<2, 256>(d, seconds, p.multiProcessorCount, random_data);
    }
  run<1, 128>(d, seconds, p.multiProcessorCount, 1);
  run<2, 128>(d, seconds, p.multiProcessorCount, 1);
  return 0;
}
