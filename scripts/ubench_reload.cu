// Micro-benchmark for the B = 1 decoder pass: what does "grid barrier + every CTA re-reads the activation rows" cost, and
// how much of it is the L2 hot spot of 148 SMs asking for the same 200 lines?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/ubench_reload scripts/ubench_reload.cu
// Experiments (148 CTAs x 256 threads, cooperative launch; one atomic-counter barrier per round, WAR covered by rotating
// between two buffers):
//   ldg  R   producers write their slice into R replicas of x, CTA b re-reads replica b % R with ld.global.cg
//   bulk R   same, the re-read is ONE cp.async.bulk of 25.6 KB into shared memory (thread 0) + mbarrier wait
//   none     barrier only (reference)
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                             \
  do {                                                                                    \
    cudaError_t e_ = (x);                                                                 \
    if (e_ != cudaSuccess) {                                                              \
      fprintf(stderr, "%s failed: %s (line %d)\n", #x, cudaGetErrorString(e_), __LINE__); \
      exit(1);                                                                            \
    }                                                                                     \
  } while (0)

constexpr int THREADS = 256;
constexpr int ITER = 2000;
constexpr int XF = 5 * 1280;

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void barrier(unsigned* counter, unsigned& target) {
  __syncthreads();
  target += gridDim.x;
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    while (static_cast<int>(ld_acquire(counter) - target) < 0) {
    }
  }
  __syncthreads();
}

// mode 0: barrier only; 1: ldg reload; 2: bulk reload into smem
template <int MODE>
__global__ void k_round(unsigned* counter, unsigned base, float* x, int R, float* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bar;
  unsigned target = base;
  const int per = (XF + gridDim.x - 1) / gridDim.x;
  const unsigned bar_a = static_cast<unsigned>(__cvta_generic_to_shared(&bar));
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
  __syncthreads();
  float acc = 0.f;
  unsigned ph = 0;
  for (int i = 0; i < ITER; ++i) {
    float* xb = x + static_cast<size_t>(i & 1) * R * XF;
    if (MODE != 0) {
      for (int j = threadIdx.x; j < per * R; j += blockDim.x) {
        const int rep = j / per, k = blockIdx.x * per + (j - rep * per);
        if (k < XF) xb[static_cast<size_t>(rep) * XF + k] = static_cast<float>(i + k);
      }
    }
    barrier(counter, target);
    const float* mine = xb + static_cast<size_t>(blockIdx.x % R) * XF;
    if (MODE == 1) {
      for (int j = threadIdx.x; j < XF / 4; j += blockDim.x) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(mine) + j);
        acc += v.x + v.y + v.z + v.w;
      }
    } else if (MODE == 2) {
      if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(XF * 4) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         static_cast<unsigned>(__cvta_generic_to_shared(smem))),
                     "l"(mine), "r"(XF * 4), "r"(bar_a)
                     : "memory");
      }
      unsigned done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar_a), "r"(ph) : "memory");
      ph ^= 1u;
      acc += reinterpret_cast<const float*>(smem)[threadIdx.x];
    }
  }
  if (acc == -1.f) *sink = acc;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  unsigned* counter;
  CK(cudaMalloc(&counter, 128));
  CK(cudaMemset(counter, 0, 128));
  float *x, *sink;
  CK(cudaMalloc(&x, 2 * 32 * XF * sizeof(float)));
  CK(cudaMalloc(&sink, 64));
  unsigned base = 0;
  auto run = [&](const void* fn, int R, const char* name) {
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEvent_t a, b;
      CK(cudaEventCreate(&a));
      CK(cudaEventCreate(&b));
      void* args[] = {&counter, &base, &x, &R, &sink};
      CK(cudaEventRecord(a));
      CK(cudaLaunchCooperativeKernel(fn, dim3(sms), dim3(THREADS), args, XF * 4, 0));
      CK(cudaEventRecord(b));
      CK(cudaEventSynchronize(b));
      base += static_cast<unsigned>(ITER) * sms;
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, a, b));
      if (ms < best) best = ms;
    }
    printf("%-6s R=%2d  %8.3f us per round\n", name, R, best * 1e3f / ITER);
  };
  run(reinterpret_cast<const void*>(k_round<0>), 1, "none");
  for (int R : {1, 2, 4, 8, 16, 32}) run(reinterpret_cast<const void*>(k_round<1>), R, "ldg");
  for (int R : {1, 2, 4, 8, 16, 32}) run(reinterpret_cast<const void*>(k_round<2>), R, "bulk");
  return 0;
}
