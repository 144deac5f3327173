// Decoder step kernels for a handful of rows (utterances x beams <= 16): the HBM-bandwidth-bound part of
// ctranslate2.models.Whisper.generate (/root/reference/main.py:687-692).
// Architecture per [HF] modeling_whisper.py:417-508 (decoder layer), :650-700 (embeddings), :966-971 (tied projection).
//
// Activations stay fp32 end to end here (weights fp16 -> fp32 on the fly, fp32 FMA): with <= 16 rows the step is bound by
// streaming the 1.8 GB of decoder weights, so the extra precision is free and keeps the arg-max decisions close to the
// fp32 oracle.  Every kernel reads the current position from DecState so a single captured CUDA graph serves all steps.
#include <cooperative_groups.h>

#include <mutex>

#include "decoder.cuh"

namespace cg = cooperative_groups;

namespace wisb {

namespace {

// =====================================================================================================================
// GEMV-like skinny GEMM: out[R, N] = f(LN?(x)[R, K] . W[N, K]^T + bias).  One CTA owns C consecutive output columns and
// the whole K range; every lane keeps its K-slices of all R rows in registers and streams C weight rows through them
// with 16-byte loads (C independent loads in flight per lane), so each weight byte is read exactly once.
// =====================================================================================================================
template <int NR, int C>
__global__ void __launch_bounds__(256)
gemv_kernel(const GemvArgs a) {
  __shared__ float s_mean[NR], s_rstd[NR];
  __shared__ float s_red[8][C][NR];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = blockDim.x >> 5;
  const int R = a.R, K = a.K;
  const int n0 = blockIdx.x * C;
  const bool ln = a.ln_g != nullptr;

  if (ln) {
    // row statistics (recomputed per CTA: R*K floats out of L2, cheaper than another launch)
    for (int r = warp; r < R; r += nwarps) {
      const float4* xr = reinterpret_cast<const float4*>(a.x + static_cast<long long>(r) * K);
      float s = 0.f;
      for (int i = lane; i < K / 4; i += 32) {
        const float4 v = xr[i];
        s += v.x + v.y + v.z + v.w;
      }
      const float mean = warp_sum(s) / K;
      float q = 0.f;
      for (int i = lane; i < K / 4; i += 32) {
        const float4 v = xr[i];
        const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
        q += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
      }
      const float var = warp_sum(q) / K;
      if (lane == 0) {
        s_mean[r] = mean;
        s_rstd[r] = rsqrtf(var + 1e-5f);
      }
    }
    __syncthreads();
  }

  float acc[C][NR];
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int r = 0; r < NR; ++r) acc[c][r] = 0.f;

  for (int kv = tid; kv < K / 8; kv += blockDim.x) {
    const int k0 = kv * 8;
    uint4 wv[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int n = n0 + c;
      wv[c] = (n < a.N) ? __ldg(reinterpret_cast<const uint4*>(a.w + static_cast<long long>(n) * K + k0)) : make_uint4(0, 0, 0, 0);
    }
    float g[8], bb[8];
    if (ln) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.ln_g + k0)), g1 = __ldg(reinterpret_cast<const float4*>(a.ln_g + k0 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.ln_b + k0)), b1 = __ldg(reinterpret_cast<const float4*>(a.ln_b + k0 + 4));
      g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
      bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
    }
    float xr[NR][8];
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      if (r < R) {
        const float4 x0 = *reinterpret_cast<const float4*>(a.x + static_cast<long long>(r) * K + k0);
        const float4 x1 = *reinterpret_cast<const float4*>(a.x + static_cast<long long>(r) * K + k0 + 4);
        xr[r][0] = x0.x; xr[r][1] = x0.y; xr[r][2] = x0.z; xr[r][3] = x0.w;
        xr[r][4] = x1.x; xr[r][5] = x1.y; xr[r][6] = x1.z; xr[r][7] = x1.w;
        if (ln) {
          const float mean = s_mean[r], rstd = s_rstd[r];
#pragma unroll
          for (int i = 0; i < 8; ++i) xr[r][i] = (xr[r][i] - mean) * rstd * g[i] + bb[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) xr[r][i] = 0.f;
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const __half2* h2 = reinterpret_cast<const __half2*>(&wv[c]);
      float wf[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __half22float2(h2[i]);
        wf[2 * i] = f.x;
        wf[2 * i + 1] = f.y;
      }
#pragma unroll
      for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[c][r] = fmaf(wf[i], xr[r][i], acc[c][r]);
    }
  }
  // reduce over lanes, then over warps
#pragma unroll
  for (int c = 0; c < C; ++c)
#pragma unroll
    for (int r = 0; r < NR; ++r) {
      const float v = warp_sum(acc[c][r]);
      if (lane == 0) s_red[warp][c][r] = v;
    }
  __syncthreads();
  if (tid < C * NR) {
    const int c = tid / NR, r = tid % NR;
    const int n = n0 + c;
    if (r < R && n < a.N) {
      float v = 0.f;
      for (int w = 0; w < nwarps; ++w) v += s_red[w][c][r];
      if (a.bias != nullptr) v += a.bias[n];
      switch (a.epi) {
        case GV_STORE:
          a.out[static_cast<long long>(r) * a.ldo + n] = v;
          break;
        case GV_RESID:
          a.out[static_cast<long long>(r) * a.ldo + n] += v;
          break;
        case GV_GELU:
          a.out[static_cast<long long>(r) * a.ldo + n] = gelu_erf(v);
          break;
        case GV_QKV: {
          const int d = a.d_model;
          if (n < d) {
            a.out[static_cast<long long>(r) * a.ldo + n] = v;
          } else {
            const int pos = a.st->pos;
            __half* cache = (n < 2 * d) ? a.kcache : a.vcache;
            const int e = (n < 2 * d) ? n - d : n - 2 * d;
            cache[(static_cast<long long>(r) * a.t_max + pos) * d + e] = __float2half_rn(v);
          }
          break;
        }
        default:
          break;
      }
    }
  }
}

__global__ void dec_embed_kernel(const int* __restrict__ tokens, const __half* __restrict__ tok_emb,
                                 const float* __restrict__ pos_emb, float* __restrict__ x, int d, const DecState* st) {
  const int r = blockIdx.x;
  const int pos = st->pos;
  const __half* e = tok_emb + static_cast<long long>(tokens[r]) * d;
  const float* p = pos_emb + static_cast<long long>(pos) * d;
  for (int i = threadIdx.x; i < d; i += blockDim.x) x[static_cast<long long>(r) * d + i] = __half2float(e[i]) + p[i];
}

// =====================================================================================================================
// self-attention over the cache: one warp per (row, head); keys t <= pos; position t of row r is stored in slot
// indir[r][t] for t < pos and in slot r for t == pos (written by this step's QKV kernel).
// =====================================================================================================================
constexpr int SA_TMAX = 448;

__global__ void __launch_bounds__(128)
dec_self_attn_kernel(const float* __restrict__ q, const __half* __restrict__ kcache, const __half* __restrict__ vcache,
                     const int* __restrict__ indir0, const int* __restrict__ indir1, const int* __restrict__ flip,
                     float* __restrict__ ctx, int d, int H, int t_max, const DecState* st) {
  __shared__ float s_p[4][SA_TMAX];
  __shared__ int s_slot[4][SA_TMAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x * 4 + warp;
  const int r = blockIdx.y;
  if (h >= H) return;
  const int pos = st->pos;
  const int* indir = (*flip ? indir1 : indir0) + static_cast<long long>(r) * t_max;
  const float* qr = q + static_cast<long long>(r) * d + h * HEAD_DIM;
  float qv[HEAD_DIM];
#pragma unroll
  for (int i = 0; i < HEAD_DIM / 4; ++i) {
    const float4 v = *reinterpret_cast<const float4*>(qr + 4 * i);
    qv[4 * i] = v.x; qv[4 * i + 1] = v.y; qv[4 * i + 2] = v.z; qv[4 * i + 3] = v.w;
  }
  float mx = -INFINITY;
  for (int t = lane; t <= pos; t += 32) {
    const int slot = (t == pos) ? r : indir[t];
    s_slot[warp][t] = slot;
    const uint4* kr = reinterpret_cast<const uint4*>(kcache + (static_cast<long long>(slot) * t_max + t) * d + h * HEAD_DIM);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 u = kr[i];
      const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h2[j]);
        s = fmaf(qv[8 * i + 2 * j], f.x, s);
        s = fmaf(qv[8 * i + 2 * j + 1], f.y, s);
      }
    }
    s *= 0.125f;
    s_p[warp][t] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int t = lane; t <= pos; t += 32) {
    const float p = __expf(s_p[warp][t] - mx);
    s_p[warp][t] = p;
    sum += p;
  }
  sum = warp_sum(sum);
  __syncwarp();
  float o0 = 0.f, o1 = 0.f;
  for (int t = 0; t <= pos; ++t) {
    const float p = s_p[warp][t];
    const __half2 v = *reinterpret_cast<const __half2*>(vcache + (static_cast<long long>(s_slot[warp][t]) * t_max + t) * d +
                                                        h * HEAD_DIM + 2 * lane);
    const float2 f = __half22float2(v);
    o0 = fmaf(p, f.x, o0);
    o1 = fmaf(p, f.y, o1);
  }
  const float inv = 1.0f / sum;
  float2* o = reinterpret_cast<float2*>(ctx + static_cast<long long>(r) * d + h * HEAD_DIM + 2 * lane);
  *o = make_float2(o0 * inv, o1 * inv);
}

// =====================================================================================================================
// cross-attention: one cluster of 8 CTAs per (head, utterance); CTA c owns keys [192 c, 192 c + 192) of the 1536-row
// padded window (keys >= 1500 masked).  Inside a CTA, groups of 8 lanes walk keys with an online softmax for all
// `beam` rows at once (K/V are read once for every beam), partials are merged through shared memory, and the 8 CTAs
// merge through distributed shared memory -- no extra kernel, no global scratch.
// =====================================================================================================================
constexpr int CA_CLUSTER = 8;
constexpr int CA_KEYS = T_ENC_PAD / CA_CLUSTER;  // 192
constexpr int CA_THREADS = 128;
constexpr int CA_GROUPS = CA_THREADS / 8;        // 16 groups of 8 lanes; 12 keys each

template <int NB>
__global__ void __cluster_dims__(1, 1, CA_CLUSTER) __launch_bounds__(CA_THREADS)
dec_cross_attn_kernel(const float* __restrict__ q, const __half* __restrict__ kmat, const __half* __restrict__ vmat,
                      float* __restrict__ ctx, int beam, int d, int H) {
  __shared__ float s_acc[CA_GROUPS][NB][HEAD_DIM];
  __shared__ float s_m[CA_GROUPS][NB], s_l[CA_GROUPS][NB];
  __shared__ float c_acc[NB][HEAD_DIM];  // this CTA's merged partial (read by the cluster leader through DSMEM)
  __shared__ float c_m[NB], c_l[NB];
  cg::cluster_group cluster = cg::this_cluster();
  const int h = blockIdx.x, u = blockIdx.y, cta = blockIdx.z;
  const int tid = threadIdx.x;
  const int grp = tid >> 3, gl = tid & 7;  // lane gl of group grp owns dims [8 gl, 8 gl + 8)
  const long long head_off = (static_cast<long long>(u) * H + h) * T_ENC_PAD * HEAD_DIM;
  const __half* kb = kmat + head_off;
  const __half* vb = vmat + head_off;

  float qv[NB][8];
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    if (k < beam) {
      const float* qr = q + static_cast<long long>(u * beam + k) * d + h * HEAD_DIM + gl * 8;
      const float4 a0 = *reinterpret_cast<const float4*>(qr), a1 = *reinterpret_cast<const float4*>(qr + 4);
      qv[k][0] = a0.x; qv[k][1] = a0.y; qv[k][2] = a0.z; qv[k][3] = a0.w;
      qv[k][4] = a1.x; qv[k][5] = a1.y; qv[k][6] = a1.z; qv[k][7] = a1.w;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) qv[k][i] = 0.f;
    }
  }
  float m[NB], l[NB], acc[NB][8];
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    m[k] = -INFINITY;
    l[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
  }
  const unsigned gmask = 0xFFu << (tid & 24);  // the 8 lanes of this group (shuffles stay inside it)
  const int t_begin = cta * CA_KEYS;
  for (int t = t_begin + grp; t < t_begin + CA_KEYS; t += CA_GROUPS) {
    if (t >= T_ENC) break;  // uniform inside the 8-lane group
    const uint4 ku = __ldg(reinterpret_cast<const uint4*>(kb + static_cast<long long>(t) * HEAD_DIM + gl * 8));
    const uint4 vu = __ldg(reinterpret_cast<const uint4*>(vb + static_cast<long long>(t) * HEAD_DIM + gl * 8));
    float kf[8], vf[8];
    {
      const __half2* k2 = reinterpret_cast<const __half2*>(&ku);
      const __half2* v2 = reinterpret_cast<const __half2*>(&vu);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 a = __half22float2(k2[i]), b = __half22float2(v2[i]);
        kf[2 * i] = a.x; kf[2 * i + 1] = a.y;
        vf[2 * i] = b.x; vf[2 * i + 1] = b.y;
      }
    }
#pragma unroll
    for (int k = 0; k < NB; ++k) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s = fmaf(qv[k][i], kf[i], s);
      s += __shfl_xor_sync(gmask, s, 1);
      s += __shfl_xor_sync(gmask, s, 2);
      s += __shfl_xor_sync(gmask, s, 4);
      s *= 0.125f;
      const float mn = fmaxf(m[k], s);
      const float al = __expf(m[k] - mn), p = __expf(s - mn);
      l[k] = l[k] * al + p;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[k][i] = fmaf(acc[k][i], al, p * vf[i]);
      m[k] = mn;
    }
  }
#pragma unroll
  for (int k = 0; k < NB; ++k) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s_acc[grp][k][gl * 8 + i] = acc[k][i];
    if (gl == 0) {
      s_m[grp][k] = m[k];
      s_l[grp][k] = l[k];
    }
  }
  __syncthreads();
  // merge the 16 groups: thread (k, e) pairs
  for (int idx = tid; idx < NB * HEAD_DIM; idx += CA_THREADS) {
    const int k = idx / HEAD_DIM, e = idx % HEAD_DIM;
    float mm = -INFINITY;
    for (int g = 0; g < CA_GROUPS; ++g) mm = fmaxf(mm, s_m[g][k]);
    float a = 0.f, ll = 0.f;
    for (int g = 0; g < CA_GROUPS; ++g) {
      const float w = (s_m[g][k] == -INFINITY) ? 0.f : __expf(s_m[g][k] - mm);
      a = fmaf(w, s_acc[g][k][e], a);
      ll = fmaf(w, s_l[g][k], ll);
    }
    c_acc[k][e] = a;
    if (e == 0) {
      c_m[k] = mm;
      c_l[k] = ll;
    }
  }
  cluster.sync();
  if (cta == 0) {
    for (int idx = tid; idx < NB * HEAD_DIM; idx += CA_THREADS) {
      const int k = idx / HEAD_DIM, e = idx % HEAD_DIM;
      if (k >= beam) continue;
      float mm = -INFINITY;
      for (int c = 0; c < CA_CLUSTER; ++c) mm = fmaxf(mm, *cluster.map_shared_rank(&c_m[k], c));
      float a = 0.f, ll = 0.f;
      for (int c = 0; c < CA_CLUSTER; ++c) {
        const float mc = *cluster.map_shared_rank(&c_m[k], c);
        const float w = (mc == -INFINITY) ? 0.f : __expf(mc - mm);
        a = fmaf(w, *cluster.map_shared_rank(&c_acc[k][e], c), a);
        ll = fmaf(w, *cluster.map_shared_rank(&c_l[k], c), ll);
      }
      ctx[static_cast<long long>(u * beam + k) * d + h * HEAD_DIM + e] = a / ll;
    }
  }
  cluster.sync();  // keep every CTA's shared memory alive until the leader has read it
}

template <int NR, int C>
void gemv_launch(const GemvArgs& a, cudaStream_t stream) {
  int threads = round_up(a.K / 8 < 160 ? a.K / 8 : 160, 32);
  if (threads < C * NR) threads = round_up(C * NR, 32);
  if (threads > 256) threads = 256;
  gemv_kernel<NR, C><<<cdiv(a.N, C), threads, 0, stream>>>(a);
  WISB_CUDA(cudaGetLastError());
}

}  // namespace

void gemv_run(const GemvArgs& a, cudaStream_t stream) {
  WISB_REQUIRE(a.K % 8 == 0 && a.R >= 1 && a.R <= DEC_MAX_ROWS, "gemv: bad shape");
  if (a.R <= 4)
    gemv_launch<4, 8>(a, stream);
  else if (a.R <= 8)
    gemv_launch<8, 8>(a, stream);
  else
    gemv_launch<16, 4>(a, stream);
}

void dec_embed_run(const int* tokens, const __half* tok_emb, const float* pos_emb, float* x, int R, int d,
                   const DecState* st, cudaStream_t stream) {
  dec_embed_kernel<<<R, 256, 0, stream>>>(tokens, tok_emb, pos_emb, x, d, st);
  WISB_CUDA(cudaGetLastError());
}

void dec_self_attn_run(const float* q, const __half* kcache, const __half* vcache, const int* indir0, const int* indir1,
                       const int* flip, float* ctx, int R, int d, int H, int t_max, const DecState* st,
                       cudaStream_t stream) {
  WISB_REQUIRE(t_max <= SA_TMAX, "self-attention: t_max > 448");
  dim3 grid(cdiv(H, 4), R);
  dec_self_attn_kernel<<<grid, 128, 0, stream>>>(q, kcache, vcache, indir0, indir1, flip, ctx, d, H, t_max, st);
  WISB_CUDA(cudaGetLastError());
}

void dec_cross_attn_run(const float* q, const __half* k, const __half* v, float* ctx, int n_utt, int beam, int d, int H,
                        cudaStream_t stream) {
  WISB_REQUIRE(beam >= 1 && beam <= MAX_BEAM, "cross-attention: beam out of range");
  dim3 grid(H, n_utt, CA_CLUSTER);
  if (beam == 1)
    dec_cross_attn_kernel<1><<<grid, CA_THREADS, 0, stream>>>(q, k, v, ctx, beam, d, H);
  else if (beam <= 5)
    dec_cross_attn_kernel<5><<<grid, CA_THREADS, 0, stream>>>(q, k, v, ctx, beam, d, H);
  else
    dec_cross_attn_kernel<MAX_BEAM><<<grid, CA_THREADS, 0, stream>>>(q, k, v, ctx, beam, d, H);
  WISB_CUDA(cudaGetLastError());
}

}  // namespace wisb
