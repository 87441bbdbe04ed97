// Experiment: does a tcgen05.mma stream slow down when consecutive instructions use different N,
// different accumulator column windows, different weight-row windows or A starts that are not
// 1024-B aligned (the column-widened cnv1 plan does all four)?  One thread, fixed smem operands.
#include <cstdio>
#include <vector>
#include "../../davo_b200/csrc/ptx.cuh"
using namespace davo;

struct Tap { int a_off, n16, dcol16, brow8; };
__constant__ Tap c_taps[128];

__global__ void __launch_bounds__(128, 1) k(long long* out, int iters, int ntaps, int fence_every, int commit_every) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + 160 * 1024);
  uint32_t* slot = (uint32_t*)(bar + 4);
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) ((float*)smem)[i] = 1.0f;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1 << 20); fence_mbar_init(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) tmem_alloc(slot, 256);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 96 * 1024);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      for (int t = 0; t < ntaps; ++t) {
        if (fence_every && t % fence_every == 0) tc_fence_after();
        if (commit_every && t % commit_every == 0) tc_commit(&bar[1]);
        const Tap td = c_taps[t];
        const uint32_t a = a0 + (uint32_t)td.a_off * 128;
        const uint64_t da = (uint64_t)((a & 0x3FFFF) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
        const uint64_t db = umma_desc_sw128(b0 + td.brow8 * 1024);
        const uint32_t id = umma_idesc_tf32(128, 0) | ((uint32_t)td.n16 << 18);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) tc_mma_tf32(tm + td.dcol16 * 16, da + 2 * kk, db + 2 * kk, id, 1);
      }
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

void run(const char* name, const std::vector<Tap>& taps, long long* d, int fence_every = 0, int commit_every = 0) {
  const int smem = 162 * 1024 + 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaMemcpyToSymbol(c_taps, taps.data(), taps.size() * sizeof(Tap));
  const int iters = 64;
  k<<<1, 128, smem>>>(d, iters, (int)taps.size(), fence_every, commit_every);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-44s %3zu taps: issue %.1f cyc/MMA, complete %.1f cyc/MMA\n", name, taps.size(),
         h[0] / (4.0 * iters * taps.size()), h[1] / (4.0 * iters * taps.size()));
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  std::vector<Tap> t;
  for (int i = 0; i < 77; ++i) t.push_back({0, 4, 0, 0});
  run("fixed: N=64, col 0, aligned A", t, d);
  t.clear(); for (int i = 0; i < 77; ++i) t.push_back({0, 4, i % 5, 0});
  run("accumulator window moves (N=64)", t, d);
  t.clear(); for (int i = 0; i < 77; ++i) t.push_back({0, 1 + i % 4, 0, 0});
  run("N changes 16..64 (col 0)", t, d);
  t.clear(); for (int i = 0; i < 77; ++i) t.push_back({0, 4, 0, (i % 4) * 2});
  run("weight row window moves", t, d);
  t.clear(); for (int i = 0; i < 77; ++i) t.push_back({(i % 4) * 2, 4, 0, 0});
  run("A start moves by 256 B (not 1024-aligned)", t, d);
  t.clear(); for (int i = 0; i < 77; ++i) t.push_back({(i % 4) * 8, 4, 0, 0});
  run("A start moves by 1024 B", t, d);
  // the cnv1 plan: pair positions c = 0..10, 7 filter rows
  t.clear();
  for (int c = -1; c <= 9; ++c)
    for (int ty = 0; ty < 7; ++ty) {
      const int g_lo = c - 2 > 0 ? c - 2 : 0, g_hi = c + 1 < 7 ? c + 1 : 7;
      t.push_back({(ty / 2) * 2, g_hi - g_lo + 1, g_lo, (2 - c + g_lo) * 2});
    }
  run("cnv1 widened plan", t, d);
  run("cnv1 plan + fence::after_thread_sync every 4 taps", t, d, 4, 0);
  run("cnv1 plan + commit every 4 taps", t, d, 0, 4);
  run("cnv1 plan + fence and commit every 4 taps", t, d, 4, 4);
  run("cnv1 plan + fence and commit every tap", t, d, 1, 1);
  // same with 128-column MMAs always
  t.clear(); for (int i = 0; i < 77; ++i) t.push_back({(i % 4) * 2, 8, 0, 0});
  run("N=128 always, A moves by 256 B", t, d);
  t.clear(); for (int i = 0; i < 77; ++i) t.push_back({(i % 28) * 3, 8, 0, 0});
  run("N=128 always, A moves by 384 B", t, d);
  t.clear(); for (int i = 0; i < 77; ++i) t.push_back({(i % 28) * 1, 8, 0, 0});
  run("N=128 always, A moves by 128 B", t, d);
  t.clear(); for (int i = 0; i < 77; ++i) t.push_back({0, 16, 0, 0});
  run("N=256, fixed", t, d);
  t.clear(); for (int i = 0; i < 77; ++i) t.push_back({(i % 28) * 1, 16, 0, 0});
  run("N=256, A moves by 128 B", t, d);
  run("N=256 + fence every tap", t, d, 1, 0);
  run("N=256 + commit every tap", t, d, 0, 1);
  run("N=256 + fence and commit every tap", t, d, 1, 1);
  run("N=256 + fence and commit every 4 taps", t, d, 4, 4);
  t.clear(); for (int i = 0; i < 77; ++i) t.push_back({(i % 28) * 1, 8, 0, 0});
  for (int e : {1, 2, 4, 8, 16, 32}) {
    char nm[64];
    snprintf(nm, 64, "N=128 + commit every %d taps", e); run(nm, t, d, 0, e);
    snprintf(nm, 64, "N=128 + fence every %d taps", e); run(nm, t, d, e, 0);
  }
  return 0;
}
