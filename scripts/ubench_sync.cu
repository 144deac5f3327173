// Micro-benchmarks for the decoder-pass redesign: what does one grid-wide exchange cost on B200?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/ubench_sync scripts/ubench_sync.cu
//   gpurun -- './gpurun_out/ubench_sync'          (prints one line per experiment; ~2 s of GPU time)
// Experiments (148 CTAs x 256 threads, cooperative launch, ITER rounds each, time per round from CUDA events):
//   flag_barrier      per-CTA epoch flags in separate 128-byte lines, st.release / ld.acquire polling (the pass kernel's)
//   cg_grid_sync      cooperative_groups::grid_group::sync()
//   atomic_barrier    one atomicAdd per CTA on a shared counter + spin
//   barrier_reload    flag_barrier + every CTA re-reads a 25.6 KB activation vector (ld.global.cg) written by all CTAs
//   tagged_release    no barrier: producers st.release 45 floats each over a sentinel, consumers poll the data words
//   tagged_fence      same with plain stores + one __threadfence per producer thread
//   cluster_barrier   barrier.cluster arrive/wait, cluster sizes 2..16
//   l2_latency        dependent-load chain over a 1 MB L2-resident buffer: idle, and while every SM keeps K x 36 KB bulk
//                     copies (cp.async.bulk) in flight -- the queueing effect that made a 5-stage weight ring slower
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

namespace cg = cooperative_groups;

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
constexpr int XF = 5 * 1280;  // floats of the exchanged activation vector (5 rows x d_model 1280)
constexpr unsigned SENT = 0x7FC0DEADu;

__device__ __forceinline__ void st_release(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_relaxed4(const float* p) {
  float4 v;
  asm volatile("ld.relaxed.gpu.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_f(float* p, float v) { asm volatile("st.release.gpu.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ bool sent4(const float4& v) {
  return __float_as_uint(v.x) == SENT || __float_as_uint(v.y) == SENT || __float_as_uint(v.z) == SENT || __float_as_uint(v.w) == SENT;
}

__device__ __forceinline__ void flag_barrier(unsigned* flags, unsigned& epoch) {
  __syncthreads();
  ++epoch;
  if (threadIdx.x == 0) st_release(flags + blockIdx.x * 32, epoch);
  if (threadIdx.x < gridDim.x) {
    const unsigned* f = flags + threadIdx.x * 32;
    while (static_cast<int>(ld_acquire(f) - epoch) < 0) {
    }
  }
  __syncthreads();
}

__global__ void k_flag_barrier(unsigned* flags, unsigned epoch0) {
  unsigned epoch = epoch0;
  for (int i = 0; i < ITER; ++i) flag_barrier(flags, epoch);
}

__global__ void k_cg_sync() {
  cg::grid_group g = cg::this_grid();
  for (int i = 0; i < ITER; ++i) g.sync();
}

__global__ void k_atomic_barrier(unsigned* counter, unsigned base) {
  unsigned target = base;
  for (int i = 0; i < ITER; ++i) {
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(counter, 1u);
      while (static_cast<int>(*reinterpret_cast<volatile unsigned*>(counter) - target) < 0) {
      }
      __threadfence();
    }
    __syncthreads();
  }
}

// barrier + reload: CTA b owns floats [b*per, ...) of x; every round it rewrites them, crosses the barrier, reads all of x
__global__ void k_barrier_reload(unsigned* flags, unsigned epoch0, float* x, float* sink) {
  unsigned epoch = epoch0;
  const int per = (XF + gridDim.x - 1) / gridDim.x;
  float acc = 0.f;
  for (int i = 0; i < ITER; ++i) {
    for (int j = threadIdx.x; j < per; j += blockDim.x) {
      const int k = blockIdx.x * per + j;
      if (k < XF) x[k] = static_cast<float>(i + k);
    }
    flag_barrier(flags, epoch);
    for (int j = threadIdx.x; j < XF / 4; j += blockDim.x) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(x) + j);
      acc += v.x + v.y + v.z + v.w;
    }
    flag_barrier(flags, epoch);  // (write-after-read: the next round overwrites x)
  }
  if (acc == -1.f) *sink = acc;
}

// tagged exchange over 4 rotating buffers: round i writes buffer i%4, reads it back complete, resets buffer (i+2)%4
template <bool kRelease>
__global__ void k_tagged(float* bufs, float* sink) {
  const int per = (XF + gridDim.x - 1) / gridDim.x;
  float acc = 0.f;
  const float sv = __uint_as_float(SENT);
  for (int i = 0; i < ITER; ++i) {
    float* out = bufs + (i & 3) * XF;
    float* rst = bufs + ((i + 2) & 3) * XF;
    __syncthreads();
    for (int j = threadIdx.x; j < per; j += blockDim.x) {
      const int k = blockIdx.x * per + j;
      if (k < XF) {
        if (kRelease) {
          st_release_f(out + k, static_cast<float>(i + k));
        } else {
          __threadfence();
          out[k] = static_cast<float>(i + k);
        }
      }
    }
    // poll: every round re-requests all incomplete slices together
    float4 v[8];
    bool need[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int idx = j * blockDim.x + threadIdx.x;
      need[j] = idx < XF / 4;
      v[j] = need[j] ? ld_relaxed4(out + 4 * idx) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    bool again = true;
    while (again) {
      again = false;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        need[j] = need[j] && sent4(v[j]);
        again |= need[j];
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (need[j]) v[j] = ld_relaxed4(out + 4 * (j * blockDim.x + threadIdx.x));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += v[j].x + v[j].w;
    __syncthreads();  // the CTA has seen the whole round: every CTA has finished reading the buffer of round i-2... i.e. (i+2)%4
    for (int j = threadIdx.x; j < per; j += blockDim.x) {
      const int k = blockIdx.x * per + j;
      if (k < XF) rst[k] = sv;
    }
  }
  if (acc == -1.f) *sink = acc;
}

__global__ void k_cluster_barrier(float* sink) {
  float acc = 0.f;
  for (int i = 0; i < ITER; ++i) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    acc += 1.f;
  }
  if (acc == -1.f) *sink = acc;
}

// dependent-load chain (thread 0 of warp 0) while warp 1's lane 0 keeps `stages` bulk copies of 36 KB in flight
__global__ void k_l2_latency(const unsigned* chain, int hops, const uint8_t* stream, size_t stream_bytes, int stages,
                             unsigned long long* out_cycles, unsigned* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bars[8];
  __shared__ volatile int stop;
  constexpr int STAGE = 36864;
  if (threadIdx.x == 0) {
    stop = 0;
    for (int s = 0; s < 8; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(&bars[s]))));
  }
  __syncthreads();
  if (threadIdx.x == 32 && stages > 0) {
    size_t off = (static_cast<size_t>(blockIdx.x) * 7919 * STAGE) % (stream_bytes - STAGE);
    unsigned phase[8] = {0};
    int s = 0;
    bool primed[8] = {false};
    while (!stop) {
      const unsigned bar = static_cast<unsigned>(__cvta_generic_to_shared(&bars[s]));
      if (primed[s]) {
        unsigned done = 0;
        while (!done)
          asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(phase[s]) : "memory");
        phase[s] ^= 1u;
      }
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(STAGE) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       static_cast<unsigned>(__cvta_generic_to_shared(smem + s * STAGE))),
                   "l"(stream + off), "r"(STAGE), "r"(bar)
                   : "memory");
      primed[s] = true;
      off += static_cast<size_t>(gridDim.x) * STAGE;
      if (off + STAGE > stream_bytes) off = (static_cast<size_t>(blockIdx.x) * STAGE) % (stream_bytes - STAGE);
      s = (s + 1) % stages;
    }
    for (int t = 0; t < stages; ++t) {  // drain
      if (!primed[t]) continue;
      const unsigned bar = static_cast<unsigned>(__cvta_generic_to_shared(&bars[t]));
      unsigned done = 0;
      while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(phase[t]) : "memory");
    }
  }
  if (threadIdx.x == 0) {
    unsigned idx = blockIdx.x * 97u;
    const unsigned n = 262144;  // 1 MB of 4-byte links
    idx %= n;
    for (int w = 0; w < 64; ++w) idx = __ldcg(chain + idx);  // warm
    const long long t0 = clock64();
    for (int h = 0; h < hops; ++h) idx = __ldcg(chain + idx);
    const long long t1 = clock64();
    out_cycles[blockIdx.x] = static_cast<unsigned long long>(t1 - t0);
    if (idx == 0xFFFFFFFFu) *sink = idx;
    stop = 1;
  }
}

template <typename F>
float timed(F&& launch) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  launch();  // warm-up
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(a));
  launch();
  CK(cudaEventRecord(b));
  CK(cudaEventSynchronize(b));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, a, b));
  return ms;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, %d iterations per experiment\n", prop.name, sms, ITER);
  unsigned* flags;
  CK(cudaMalloc(&flags, 256 * 32 * sizeof(unsigned)));
  CK(cudaMemset(flags, 0, 256 * 32 * sizeof(unsigned)));
  float *x, *sink, *bufs;
  CK(cudaMalloc(&x, XF * sizeof(float)));
  CK(cudaMalloc(&sink, 64));
  CK(cudaMalloc(&bufs, 4 * XF * sizeof(float)));
  unsigned epoch = 0;
  auto coop = [&](const void* fn, void** args, size_t smem = 0) { CK(cudaLaunchCooperativeKernel(fn, dim3(sms), dim3(THREADS), args, smem, 0)); };

  {
    auto run = [&] {
      void* args[] = {&flags, &epoch};
      coop(reinterpret_cast<const void*>(k_flag_barrier), args);
      epoch += ITER;
    };
    printf("flag_barrier      %8.3f us per barrier\n", timed(run) * 1e3f / ITER);
  }
  {
    auto run = [&] {
      void* args[] = {nullptr};
      coop(reinterpret_cast<const void*>(k_cg_sync), args);
    };
    printf("cg_grid_sync      %8.3f us per barrier\n", timed(run) * 1e3f / ITER);
  }
  {
    unsigned* counter;
    CK(cudaMalloc(&counter, 128));
    CK(cudaMemset(counter, 0, 128));
    unsigned base = 0;
    auto run = [&] {
      void* args[] = {&counter, &base};
      coop(reinterpret_cast<const void*>(k_atomic_barrier), args);
      base += static_cast<unsigned>(ITER) * sms;
    };
    printf("atomic_barrier    %8.3f us per barrier\n", timed(run) * 1e3f / ITER);
  }
  {
    auto run = [&] {
      void* args[] = {&flags, &epoch, &x, &sink};
      coop(reinterpret_cast<const void*>(k_barrier_reload), args);
      epoch += 2 * ITER;
    };
    printf("barrier_reload    %8.3f us per round (write 25.6 KB slice-wise, barrier, everyone reads it, barrier)\n", timed(run) * 1e3f / ITER);
  }
  for (int rel = 1; rel >= 0; --rel) {
    std::vector<unsigned> init(4 * XF, SENT);
    auto run = [&] {
      CK(cudaMemcpy(bufs, init.data(), init.size() * 4, cudaMemcpyHostToDevice));
      void* args[] = {&bufs, &sink};
      coop(rel ? reinterpret_cast<const void*>(k_tagged<true>) : reinterpret_cast<const void*>(k_tagged<false>), args);
    };
    printf("tagged_%-10s %8.3f us per round (no barrier: sentinel-polled exchange of the same vector)\n", rel ? "release" : "fence",
           timed(run) * 1e3f / ITER);
  }
  for (int cs = 2; cs <= 16; cs *= 2) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((sms / cs) * cs);
    cfg.blockDim = dim3(THREADS);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    if (cs > 8) CK(cudaFuncSetAttribute(k_cluster_barrier, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaError_t probe = cudaLaunchKernelEx(&cfg, k_cluster_barrier, sink);
    if (probe != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
      printf("cluster_barrier   n/a for a cluster of %d (%s)\n", cs, cudaGetErrorString(probe));
      cudaGetLastError();
      continue;
    }
    auto run = [&] { CK(cudaLaunchKernelEx(&cfg, k_cluster_barrier, sink)); };
    printf("cluster_barrier   %8.3f us per barrier (cluster of %d)\n", timed(run) * 1e3f / ITER, cs);
  }
  {
    const unsigned n = 262144;
    std::vector<unsigned> chain(n);
    for (unsigned i = 0; i < n; ++i) chain[i] = (i * 40503u + 12345u) % n;  // full-period LCG step: a pseudo-random walk
    unsigned* d_chain;
    CK(cudaMalloc(&d_chain, n * 4));
    CK(cudaMemcpy(d_chain, chain.data(), n * 4, cudaMemcpyHostToDevice));
    const size_t stream_bytes = 1ull << 30;
    uint8_t* stream;
    CK(cudaMalloc(&stream, stream_bytes));
    CK(cudaMemset(stream, 1, stream_bytes));
    unsigned long long* cyc;
    CK(cudaMalloc(&cyc, 256 * 8));
    CK(cudaFuncSetAttribute(k_l2_latency, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * 36864));
    const int hops = 4096;
    for (int stages = 0; stages <= 5; ++stages) {
      unsigned* sk = reinterpret_cast<unsigned*>(sink);
      k_l2_latency<<<sms, 64, 5 * 36864>>>(d_chain, hops, stream, stream_bytes, stages, cyc, sk);
      CK(cudaDeviceSynchronize());
      std::vector<unsigned long long> h(sms);
      CK(cudaMemcpy(h.data(), cyc, sms * 8, cudaMemcpyDeviceToHost));
      double sum = 0;
      for (int i = 0; i < sms; ++i) sum += static_cast<double>(h[i]);
      printf("l2_latency        %8.1f cycles per dependent L2 load with %d x 36 KB bulk copies in flight per SM\n", sum / sms / hops, stages);
    }
  }
  return 0;
}
