// Whisper encoder pieces that are not plain GEMMs: conv1 stem, LayerNorm, non-causal self-attention.
//
// Replaces (reference side) the encoder half of ctranslate2.models.Whisper.generate
// (/root/reference/main.py:687-692).  Architecture per
// [HF] modeling_whisper.py:567-568,619-625 (conv stem), :361-415 (encoder layer).
#include <mutex>

#include "kernels.h"
#include "ptx.cuh"

namespace wisb {

void make_tmap_f16_2d(CUtensorMap* map, const void* ptr, long long cols, long long rows, long long ld, int box_cols,
                      int box_rows);

namespace {

// =====================================================================================================================
// conv1: Conv1d(80 -> d, k=3, pad=1) + GELU.  1.8 GFLOP per window (0.1 % of the encoder): fp32 FMA, smem tiled.
// Output row layout (fp16 [B, 3072, d]): row 0 = zeros (left pad), rows 1..3000 = frames, rows 3001.. = zeros, so that
// conv2 (k=3, stride 2, pad 1) is a plain GEMM whose A row t' is the 3*d contiguous values starting at row 2*t'.
// =====================================================================================================================
constexpr int C1_FT = 64;   // frames per CTA
constexpr int C1_CT = 64;   // output channels per CTA
constexpr int C1_K = 3 * N_MELS;
constexpr int C1_SMEM = (N_MELS * (C1_FT + 2) + C1_K * (C1_CT + 1)) * 4;

__global__ void __launch_bounds__(256)
conv1_gelu_kernel(const float* __restrict__ mel, const __half* __restrict__ w, const float* __restrict__ bias,
                  __half* __restrict__ h1, int d) {
  extern __shared__ float c1_smem[];
  float (*xm)[C1_FT + 2] = reinterpret_cast<float (*)[C1_FT + 2]>(c1_smem);
  float (*wt)[C1_CT + 1] = reinterpret_cast<float (*)[C1_CT + 1]>(c1_smem + N_MELS * (C1_FT + 2));
  const int b = blockIdx.z;
  const int f0 = blockIdx.x * C1_FT;
  const int c0 = blockIdx.y * C1_CT;
  const int tid = threadIdx.x;
  const float* melb = mel + static_cast<long long>(b) * N_MELS * N_FRAMES;
  for (int i = tid; i < N_MELS * (C1_FT + 2); i += 256) {
    const int ci = i / (C1_FT + 2), fl = i % (C1_FT + 2);
    const int f = f0 + fl - 1;
    xm[ci][fl] = (f >= 0 && f < N_FRAMES) ? melb[ci * N_FRAMES + f] : 0.f;
  }
  for (int i = tid; i < C1_CT * C1_K; i += 256) {
    const int co = i / C1_K, kk = i % C1_K;
    wt[kk][co] = __half2float(w[static_cast<long long>(c0 + co) * C1_K + kk]);
  }
  __syncthreads();
  const int tc = tid % 16;  // 4 channels each
  const int tf = tid / 16;  // 4 frames each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int kk = 0; kk < C1_K; ++kk) {
    const int k = kk / N_MELS, ci = kk - k * N_MELS;
    float a[4], bb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = xm[ci][tf * 4 + i + k];
#pragma unroll
    for (int j = 0; j < 4; ++j) bb[j] = wt[kk][tc * 4 + j];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int f = f0 + tf * 4 + i;
    if (f >= N_FRAMES) continue;
    __half* o = h1 + (static_cast<long long>(b) * H1_ROWS + 1 + f) * d + c0 + tc * 4;
    __half2 v0 = __floats2half2_rn(gelu_erf(acc[i][0] + bias[c0 + tc * 4 + 0]), gelu_erf(acc[i][1] + bias[c0 + tc * 4 + 1]));
    __half2 v1 = __floats2half2_rn(gelu_erf(acc[i][2] + bias[c0 + tc * 4 + 2]), gelu_erf(acc[i][3] + bias[c0 + tc * 4 + 3]));
    uint2 u;
    u.x = *reinterpret_cast<uint32_t*>(&v0);
    u.y = *reinterpret_cast<uint32_t*>(&v1);
    *reinterpret_cast<uint2*>(o) = u;
  }
}

// =====================================================================================================================
// LayerNorm: one warp per row, fp32 in, fp16 out (feeds the next GEMM's A operand). eps = 1e-5, biased variance.
// =====================================================================================================================
constexpr int LN_MAX_IT = 12;  // d <= 1536

__global__ void __launch_bounds__(256)
layernorm_f32_to_f16_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                            __half* __restrict__ y, int rows, int d) {
  pdl_launch_dependents();  // (programmatic dependent launch: the next kernel's prologue overlaps this one)
  pdl_wait();               // the producer of x has completed
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int iters = d / 128;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * d);
  float4 v[LN_MAX_IT];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_IT; ++i)
    if (i < iters) {
      v[i] = xr[i * 32 + lane];
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
  const float mean = warp_sum(s) / d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAX_IT; ++i)
    if (i < iters) {
      const float a0 = v[i].x - mean, a1 = v[i].y - mean, a2 = v[i].z - mean, a3 = v[i].w - mean;
      q += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
    }
  const float rstd = rsqrtf(warp_sum(q) / d + 1e-5f);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  const float4* b4 = reinterpret_cast<const float4*>(b);
  uint2* yr = reinterpret_cast<uint2*>(y + static_cast<long long>(row) * d);
#pragma unroll
  for (int i = 0; i < LN_MAX_IT; ++i)
    if (i < iters) {
      const float4 gg = __ldg(g4 + i * 32 + lane), bb = __ldg(b4 + i * 32 + lane);
      __half2 h0 = __floats2half2_rn((v[i].x - mean) * rstd * gg.x + bb.x, (v[i].y - mean) * rstd * gg.y + bb.y);
      __half2 h1 = __floats2half2_rn((v[i].z - mean) * rstd * gg.z + bb.z, (v[i].w - mean) * rstd * gg.w + bb.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&h0);
      u.y = *reinterpret_cast<uint32_t*>(&h1);
      yr[i * 32 + lane] = u;
    }
}

// =====================================================================================================================
// Encoder self-attention on tcgen05: one CTA per (128-query tile, head, window), 2 CTAs per SM.
//   S = Q K^T          : tcgen05.mma M=128 N=128 K=64 (4 instr), fp32 in TMEM cols [0,128)
//   softmax            : 128 threads, one query row each, tcgen05.ld from TMEM, exp2, P (fp16) -> swizzled smem
//   O_j = P V_j        : tcgen05.mma M=128 N=64 K=128 (8 instr), TMEM cols [128,192); rescaled/accumulated in registers
// Keys >= 1500 (padding rows of the 1536-row window layout) are masked.
// V operand: kVMN = true  -> V tile as loaded [kv][dh] (MN-major B descriptor)
//            kVMN = false -> pre-transposed Vt [dh][kv] written by the QKV GEMM epilogue (K-major B descriptor)
// =====================================================================================================================
constexpr int AT_THREADS = 192;
constexpr int AT_BM = 128, AT_BN = 128;
constexpr int AT_NB = T_ENC_PAD / AT_BN;  // 12 key blocks
constexpr int AT_TILE = 128 * 64 * 2;     // 16 KB
constexpr int AT_SMEM = AT_TILE /*Q*/ + 2 * AT_TILE /*K*/ + 2 * AT_TILE /*V*/ + 2 * AT_TILE /*P*/ + 128;

template <bool kVMN>
__global__ void __launch_bounds__(AT_THREADS, 2)
enc_attn_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                const __grid_constant__ CUtensorMap map_v, __half* __restrict__ ctx, int d, int H) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t base = smem_u32(smem);
  if ((base & 1023u) != 0) __trap();
  const uint32_t sQ = base;
  const uint32_t sK = base + AT_TILE;
  const uint32_t sV = base + 3 * AT_TILE;
  const uint32_t sP = base + 5 * AT_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * AT_TILE);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t q_full = bar0;
  auto kv_full = [&](int s) { return bar0 + 8u * (1 + s); };
  auto kv_empty = [&](int s) { return bar0 + 8u * (3 + s); };
  const uint32_t s_full = bar0 + 8u * 5, s_empty = bar0 + 8u * 6, p_full = bar0 + 8u * 7, o_full = bar0 + 8u * 8;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(bars + 9);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int row0 = b * T_ENC_PAD + qt * AT_BM;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_empty, 4);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&map_q);
    tma_prefetch_desc(&map_k);
    tma_prefetch_desc(&map_v);
  }
  if (warp == 1) {
    tmem_alloc<256>(smem_u32(const_cast<uint32_t*>(tmem_slot)));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_o = tmem_base + 128;
  pdl_launch_dependents();
  pdl_wait();  // barriers, TMEM and descriptors were set up under the QKV GEMM's tail; its output is visible from here on

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, AT_TILE);
      tma_load_2d(sQ, &map_q, q_full, h * HEAD_DIM, row0);
      for (int j = 0; j < AT_NB; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1u;
        mbar_wait(kv_empty(st), ph ^ 1u);
        mbar_arrive_expect_tx(kv_full(st), 2 * AT_TILE);
        tma_load_2d(sK + st * AT_TILE, &map_k, kv_full(st), d + h * HEAD_DIM, b * T_ENC_PAD + j * AT_BN);
        if (kVMN) {
          tma_load_2d(sV + st * AT_TILE, &map_v, kv_full(st), 2 * d + h * HEAD_DIM, b * T_ENC_PAD + j * AT_BN);
        } else {
          // Vt viewed as [B*H*64 rows, 1536 cols]; two 64x64 halves of the key block
          tma_load_2d(sV + st * AT_TILE, &map_v, kv_full(st), j * AT_BN, (b * H + h) * HEAD_DIM);
          tma_load_2d(sV + st * AT_TILE + AT_TILE / 2, &map_v, kv_full(st), j * AT_BN + 64, (b * H + h) * HEAD_DIM);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_f16(AT_BM, AT_BN, false, false);
      constexpr uint32_t idesc_o = make_idesc_f16(AT_BM, HEAD_DIM, false, kVMN);
      mbar_wait(q_full, 0);
      for (int j = 0; j < AT_NB; ++j) {
        const int st = j & 1;
        mbar_wait(kv_full(st), (j >> 1) & 1u);
        if (j > 0) mbar_wait(s_empty, (j - 1) & 1u);
        tc_fence_after();
        const uint64_t dq = make_desc_sw128(sQ, 1024);
        const uint64_t dk = make_desc_sw128(sK + st * AT_TILE, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_s, dq + 2u * k, dk + 2u * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(s_full);
        mbar_wait(p_full, j & 1u);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t dp = make_desc_sw128(sP + (k >> 2) * AT_TILE, 1024) + 2u * (k & 3);
          uint64_t dv;
          if (kVMN)
            dv = make_desc_sw128(sV + st * AT_TILE + k * 2048, 1024);
          else
            dv = make_desc_sw128(sV + st * AT_TILE + (k >> 2) * (AT_TILE / 2), 1024) + 2u * (k & 3);
          umma_f16_ss(tmem_o, dp, dv, idesc_o, k != 0 ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(kv_empty(st));
      }
    }
  } else {
    const int q = warp & 3;
    const int r = q * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const float c = 0.125f * 1.4426950408889634f;  // dh^-0.5 * log2(e)
    float m_run = -INFINITY, l_run = 0.f;
    float acc[HEAD_DIM];
#pragma unroll
    for (int i = 0; i < HEAD_DIM; ++i) acc[i] = 0.f;
    uint8_t* sp_gen = smem + 5 * AT_TILE;
    for (int j = 0; j < AT_NB; ++j) {
      mbar_wait(s_full, j & 1u);
      tc_fence_after();
      const int kv0 = j * AT_BN;
      const bool tail = (kv0 + AT_BN > T_ENC);
      float mx = -INFINITY;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_s + lane_off + ch * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float s = __uint_as_float(v[i]);
          if (tail && kv0 + ch * 32 + i >= T_ENC) s = -INFINITY;
          mx = fmaxf(mx, s);
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = exp2f((m_run - m_new) * c);
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(tmem_o + lane_off + ch * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) acc[ch * 32 + i] = (acc[ch * 32 + i] + __uint_as_float(v[i])) * alpha;
        }
      }
      float rowsum = 0.f;
      const float mc = m_new * c;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(tmem_s + lane_off + ch * 32, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float p0 = exp2f(fmaf(__uint_as_float(v[2 * i]), c, -mc));
          float p1 = exp2f(fmaf(__uint_as_float(v[2 * i + 1]), c, -mc));
          if (tail) {
            if (kv0 + ch * 32 + 2 * i >= T_ENC) p0 = 0.f;
            if (kv0 + ch * 32 + 2 * i + 1 >= T_ENC) p1 = 0.f;
          }
          rowsum += p0 + p1;
          __half2 hh = __floats2half2_rn(p0, p1);
          pk[i] = *reinterpret_cast<uint32_t*>(&hh);
        }
        // P is the K-major A operand [128 rows, 128 keys] as two 64-key halves of 128-byte swizzled rows
        uint8_t* half_base = sp_gen + (ch >> 1) * AT_TILE + r * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int chunk = (ch & 1) * 4 + g;  // 16-byte chunk index inside the 128-byte row
          uint4 u = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
          *reinterpret_cast<uint4*>(half_base + ((chunk ^ (r & 7)) << 4)) = u;
        }
      }
      l_run = l_run * alpha + rowsum;
      m_run = m_new;
      tc_fence_before();
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(s_empty);
        mbar_arrive(p_full);
      }
    }
    mbar_wait(o_full, (AT_NB - 1) & 1u);
    tc_fence_after();
    const float inv = 1.0f / l_run;
    __half* o = ctx + static_cast<long long>(row0 + r) * d + h * HEAD_DIM;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(tmem_o + lane_off + ch * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t w4[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int e = g * 8 + 2 * i;
          __half2 hh = __floats2half2_rn((acc[ch * 32 + e] + __uint_as_float(v[e])) * inv,
                                         (acc[ch * 32 + e + 1] + __uint_as_float(v[e + 1])) * inv);
          w4[i] = *reinterpret_cast<uint32_t*>(&hh);
        }
        *reinterpret_cast<uint4*>(o + ch * 32 + g * 8) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}

// SIMT cross-check (diagnostics / tests only): one thread per query row, online softmax in fp32
__global__ void __launch_bounds__(128)
enc_attn_ref_kernel(const __half* __restrict__ qkv, __half* __restrict__ ctx, int d) {
  const int b = blockIdx.z, h = blockIdx.y;
  const int t = blockIdx.x * 128 + threadIdx.x;
  const long long ld = 3LL * d;
  const __half* base = qkv + static_cast<long long>(b) * T_ENC_PAD * ld;
  float q[HEAD_DIM], acc[HEAD_DIM];
  for (int e = 0; e < HEAD_DIM; ++e) {
    q[e] = __half2float(base[t * ld + h * HEAD_DIM + e]);
    acc[e] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  for (int k = 0; k < T_ENC; ++k) {
    const __half* kr = base + k * ld + d + h * HEAD_DIM;
    float s = 0.f;
    for (int e = 0; e < HEAD_DIM; ++e) s = fmaf(q[e], __half2float(kr[e]), s);
    s *= 0.125f;
    const float mn = fmaxf(m, s);
    const float a = __expf(m - mn), p = __expf(s - mn);
    const __half* vr = base + k * ld + 2 * d + h * HEAD_DIM;
    for (int e = 0; e < HEAD_DIM; ++e) acc[e] = acc[e] * a + p * __half2float(vr[e]);
    l = l * a + p;
    m = mn;
  }
  __half* o = ctx + (static_cast<long long>(b) * T_ENC_PAD + t) * d + h * HEAD_DIM;
  for (int e = 0; e < HEAD_DIM; ++e) o[e] = __float2half_rn(acc[e] / l);
}

}  // namespace

void conv1_gelu_run(const float* mel, const __half* w, const float* bias, __half* h1, int B, int d, cudaStream_t stream) {
  WISB_REQUIRE(d % C1_CT == 0, "conv1: d_model must be a multiple of 64");
  static std::atomic<unsigned long long> once{0};
  once_per_device(once, [] {
    WISB_CUDA(cudaFuncSetAttribute(conv1_gelu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C1_SMEM));
  });
  dim3 grid(cdiv(N_FRAMES, C1_FT), d / C1_CT, B);
  conv1_gelu_kernel<<<grid, 256, C1_SMEM, stream>>>(mel, w, bias, h1, d);
  WISB_CUDA(cudaGetLastError());
}

void layernorm_f32_to_f16_run(const float* x, const float* g, const float* b, __half* y, int rows, int d,
                              cudaStream_t stream, bool pdl) {
  WISB_REQUIRE(d % 128 == 0 && d <= 128 * LN_MAX_IT, "layernorm: d_model must be a multiple of 128, <= 1536");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(cdiv(rows, 8));
  cfg.blockDim = dim3(256);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  WISB_CUDA(cudaLaunchKernelEx(&cfg, layernorm_f32_to_f16_kernel, x, g, b, y, rows, d));
}

void enc_attn_plan(AttnPlan& p, const __half* qkv, const __half* vt, __half* ctx, int B, int d, int H, bool v_mn_major) {
  p.B = B;
  p.d = d;
  p.H = H;
  p.ctx = ctx;
  p.v_mn_major = v_mn_major;
  const long long rows = static_cast<long long>(B) * T_ENC_PAD;
  make_tmap_f16_2d(&p.map_q, qkv, 3LL * d, rows, 3LL * d, 64, 128);
  make_tmap_f16_2d(&p.map_k, qkv, 3LL * d, rows, 3LL * d, 64, 128);
  if (v_mn_major)
    make_tmap_f16_2d(&p.map_v, qkv, 3LL * d, rows, 3LL * d, 64, 128);
  else
    make_tmap_f16_2d(&p.map_v, vt, T_ENC_PAD, static_cast<long long>(B) * H * HEAD_DIM, T_ENC_PAD, 64, 64);
}

void enc_attn_run(const AttnPlan& p, cudaStream_t stream) {
  static std::atomic<unsigned long long> once{0};
  once_per_device(once, [] {
    WISB_CUDA(cudaFuncSetAttribute(enc_attn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
    WISB_CUDA(cudaFuncSetAttribute(enc_attn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
  });
  dim3 grid(T_ENC_PAD / AT_BM, p.H, p.B);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(AT_THREADS);
  cfg.dynamicSmemBytes = AT_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = p.pdl ? 1 : 0;
  if (p.v_mn_major)
    WISB_CUDA(cudaLaunchKernelEx(&cfg, enc_attn_kernel<true>, p.map_q, p.map_k, p.map_v, p.ctx, p.d, p.H));
  else
    WISB_CUDA(cudaLaunchKernelEx(&cfg, enc_attn_kernel<false>, p.map_q, p.map_k, p.map_v, p.ctx, p.d, p.H));
}

void enc_attn_ref_run(const __half* qkv, __half* ctx, int B, int d, int H, cudaStream_t stream) {
  dim3 grid(T_ENC_PAD / 128, H, B);
  enc_attn_ref_kernel<<<grid, 128, 0, stream>>>(qkv, ctx, d);
  WISB_CUDA(cudaGetLastError());
}

}  // namespace wisb
