// Experiment: can a UMMA SWIZZLE_128B K-major descriptor start at a 128-B row that is not
// 1024-B aligned (row-shifted window into a TMA-written patch)?  Tries base_offset = 0 and
// base_offset = (addr >> 7) & 7, with SBO = 1024 and 2048.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include "../../davo_b200/csrc/ptx.cuh"
using namespace davo;

constexpr int ROWS = 512, N = 64;

__global__ void __launch_bounds__(128, 1)
k(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out,
  int shift, int sbo_bytes, int use_base_off) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                 // 512 rows x 128 B = 64 KB
  uint8_t* sB = smem + ROWS * 128;    // 64 x 128 B = 8 KB
  uint64_t* bar = (uint64_t*)(sB + N * 128);
  uint32_t* slot = (uint32_t*)(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(slot, 64);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar[0], ROWS * 128 + N * 128);
    tma_load_2d(sA, &tmA, &bar[0], 0, 0);
    tma_load_2d(sA + 256 * 128, &tmA, &bar[0], 0, 256);
    tma_load_2d(sB, &tmB, &bar[0], 0, 0);
    mbar_wait(&bar[0], 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sA) + shift * 128;
    uint64_t da = (uint64_t)((a_addr & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
                  (1ull << 46) | (2ull << 61);
    if (use_base_off) da |= (uint64_t)((a_addr >> 7) & 7) << 49;
    const uint64_t db = umma_desc_sw128(smem_u32(sB));
    const uint32_t idesc = umma_idesc_tf32(128, N);
    for (int kk = 0; kk < 4; ++kk) tc_mma_tf32(tm, da + 2 * kk, db + 2 * kk, idesc, kk != 0);
    tc_commit(&bar[1]);
  }
  __syncthreads();
  mbar_wait(&bar[1], 0);
  tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t v[32];
    tmem_ld_32x32(tm + (uint32_t(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 64); }
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  Enc enc = (Enc)fp;
  std::vector<float> A(ROWS * 32), B(N * 32);
  for (auto& v : A) v = (float)((rand() % 17) - 8);       // small ints: exact in tf32
  for (auto& v : B) v = (float)((rand() % 9) - 4);
  float *dA, *dB, *dO;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dO, 128 * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap tA, tB;
  cuuint64_t st[1] = {128}; cuuint32_t es[2] = {1, 1};
  cuuint64_t dimA[2] = {32, ROWS}; cuuint32_t boxA[2] = {32, 256};
  cuuint64_t dimB[2] = {32, N}; cuuint32_t boxB[2] = {32, N};
  enc(&tA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dA, dimA, st, boxA, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  enc(&tB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dB, dimB, st, boxB, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int smem = ROWS * 128 + N * 128 + 1024 + 64;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> O(128 * N);
  for (int sbo : {1024, 2048, 1152 /* 9 rows: groups at different phases */})
    for (int ub = 0; ub < 2; ++ub)
      for (int shift : {0, 1, 2, 3, 4, 7, 8, 9}) {
        k<<<1, 128, smem>>>(tA, tB, dO, shift, sbo, ub);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sbo %d ub %d shift %d: CUDA error %s\n", sbo, ub, shift, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0; int bad = 0;
        for (int m = 0; m < 128; ++m) {
          const int row = shift + (m / 8) * (sbo / 128) + (m % 8);   // expected source row
          for (int n = 0; n < N; ++n) {
            double r = 0;
            for (int kk = 0; kk < 32; ++kk) r += (double)A[row * 32 + kk] * B[n * 32 + kk];
            double d = fabs(r - O[m * N + n]);
            if (d > maxerr) maxerr = d;
            if (d > 1e-3) ++bad;
          }
        }
        printf("sbo %4d base_off %d shift %d : max err %.3g  bad %d / %d\n", sbo, ub, shift, maxerr, bad, 128 * N);
      }
  return 0;
}
