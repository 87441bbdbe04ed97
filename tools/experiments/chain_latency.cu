// Experiment: what one kernel boundary costs in a chain of small dependent kernels (the batch-1 forward is ten of them).
//   mode 0  plain stream order
//   mode 1  programmatic dependent launch: launch_dependents at entry, griddepcontrol.wait before the first read
//           (what the library does)
//   mode 2  programmatic dependent launch WITHOUT griddepcontrol.wait: the consumer polls a global counter that the
//           producer's CTAs bump (release) after their last store; launch_dependents only after the wait, so that at
//           most the next grid is resident and spinning
// Each kernel: `ctas` CTAs of 128 threads; every thread reads one float the previous kernel wrote (other CTA's slot)
// and writes one.  Time per kernel = the boundary cost + ~1 us of dependent loads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chain_latency chain_latency.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(128) link(const float* in, float* out, unsigned* ctr_prev, unsigned* ctr_mine, unsigned target,
                                            int mode, long long spin_limit) {
  if (mode == 1) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
  } else if (mode == 2) {
    if (threadIdx.x == 0 && ctr_prev) {
      const long long t0 = clock64();
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr_prev) : "memory");
        if (clock64() - t0 > spin_limit) __trap();
      } while (v < target);
    }
    __syncthreads();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  const int n = gridDim.x * blockDim.x;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float v = __ldcg(in + (i + 4099) % n);          // another CTA's value: wrong if the previous kernel is not done
  out[i] = v + 1.0f;
  if (mode == 2) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr_mine) : "memory");
    }
  }
}

int main(int argc, char** argv) {
  const int K = 200;
  float *a, *b;
  unsigned* ctr;
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaMalloc(&a, 1 << 22);
  cudaMalloc(&b, 1 << 22);
  cudaMalloc(&ctr, (K + 1) * sizeof(unsigned));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int ctas : {26, 52, 148, 592}) {
    for (int mode = 0; mode < 3; ++mode) {
      float best = 1e9f;
      bool ok = true;
      for (int rep = 0; rep < 5; ++rep) {
        cudaMemsetAsync(a, 0, 1 << 22, st);
        cudaMemsetAsync(b, 0, 1 << 22, st);
        cudaMemsetAsync(ctr, 0, (K + 1) * sizeof(unsigned), st);
        cudaStreamSynchronize(st);
        cudaEventRecord(e0, st);
        for (int k = 0; k < K; ++k) {
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3(ctas);
          cfg.blockDim = dim3(128);
          cfg.stream = st;
          cudaLaunchAttribute at[1];
          at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
          at[0].val.programmaticStreamSerializationAllowed = 1;
          cfg.attrs = at;
          cfg.numAttrs = mode ? 1 : 0;
          const float* in = (k & 1) ? b : a;
          float* out = (k & 1) ? a : b;
          unsigned* prev = k ? ctr + k - 1 : nullptr;
          cudaLaunchKernelEx(&cfg, link, in, out, prev, ctr + k, (unsigned)ctas, mode, 2000000000LL);
        }
        cudaEventRecord(e1, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
        // every value must be K: each link added 1 to what the previous one wrote
        static float h[592 * 128];
        cudaMemcpy(h, (K & 1) ? b : a, ctas * 128 * sizeof(float), cudaMemcpyDeviceToHost);
        for (int i = 0; i < ctas * 128; ++i) ok = ok && h[i] == (float)K;
      }
      printf("ctas %3d  mode %d  %.3f us per kernel  chain %s\n", ctas, mode, best * 1e3f / K, ok ? "correct" : "WRONG");
    }
  }
  return 0;
}
