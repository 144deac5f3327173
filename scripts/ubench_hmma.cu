// Latency / issue rate of the warp-level tensor path (mma.sync.m16n8k16 f16 -> f32, ldmatrix) on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/ubench_hmma scripts/ubench_hmma.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <int CHAINS>
__global__ void k(long long* out, float* sink, int iters) {
  uint32_t a[4] = {threadIdx.x, 1u, 2u, 3u};
  float c[CHAINS][4];
  for (int j = 0; j < CHAINS; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) mma(c[j], a, 5u, 6u);
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int j = 0; j < CHAINS; ++j) s += c[j][0] + c[j][3];
  if (s == 123.f) *sink = s;
  if (threadIdx.x % 32 == 0) out[blockIdx.x * 32 + threadIdx.x / 32] = t1 - t0;
}
__global__ void k_ldsm(long long* out, float* sink, int iters) {
  __shared__ __align__(128) uint8_t sm[8192];
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i & 1023;
  __syncthreads();
  uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(sm)) + (threadIdx.x & 31) * 128;
  uint32_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t r[4];
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr + (acc & 16)));
    acc += r[0] & 1;  // dependent address: latency chain
  }
  const long long t1 = clock64();
  if (acc == 12345u) *sink = 1.f;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
int main() {
  long long* out; float* sink;
  cudaMalloc(&out, 148 * 32 * 8); cudaMalloc(&sink, 4);
  const int iters = 4096;
  auto report = [&](const char* name, int n, int per_iter) {
    cudaDeviceSynchronize();
    long long h[32]; cudaMemcpy(h, out, sizeof(long long) * n, cudaMemcpyDeviceToHost);
    printf("%-40s %.1f cycles per HMMA-slot (warp 0), err %s\n", name, double(h[0]) / iters / per_iter, cudaGetErrorString(cudaGetLastError()));
  };
  k<1><<<1, 32>>>(out, sink, iters); report("1 warp, 1 dependent chain", 1, 1);
  k<2><<<1, 32>>>(out, sink, iters); report("1 warp, 2 chains (per HMMA)", 1, 2);
  k<4><<<1, 32>>>(out, sink, iters); report("1 warp, 4 chains (per HMMA)", 1, 4);
  k<8><<<1, 32>>>(out, sink, iters); report("1 warp, 8 chains (per HMMA)", 1, 8);
  k<4><<<1, 128>>>(out, sink, iters); report("4 warps (1/SMSP), 4 chains (per HMMA)", 4, 4);
  k<4><<<1, 256>>>(out, sink, iters); report("8 warps (2/SMSP), 4 chains (per HMMA)", 8, 4);
  k<1><<<1, 224>>>(out, sink, iters); report("7 warps, 1 chain", 7, 1);
  k_ldsm<<<1, 32>>>(out, sink, iters); report("ldmatrix.x4 dependent latency", 1, 1);
  return 0;
}
