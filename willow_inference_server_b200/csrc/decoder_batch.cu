// Batched decoder pass: ONE pass over the decoder weights serves every row (utterance x beam) of a batch of utterances.
//
// Reference semantics: the decoder step of ctranslate2.models.Whisper.generate for a batch of feature windows
// (/root/reference/main.py:676-693 feeds several windows per call; SURVEY.md section 8a row A10, section 7 step 5:
// "decoder with M = B x beam rows").  Architecture per [HF] modeling_whisper.py:417-508.
//
// Up to 8 rows the persistent SIMT pass (decoder_mega.cu) is the latency path.  Beyond that the pass is a chain of
//   * tcgen05 GEMMs (gemm_tc.cu): rows are the M dimension (padded to 128-row tiles), the weight matrix streams through
//     the TMA ring exactly once per pass whatever the number of rows; narrow tiles (BN = 64) and split-K keep >= ~100 CTAs
//     streaming even when the weight matrix has only d_model output columns;
//   * the small kernels below: embedding + LayerNorm, split-K reduction + bias + residual + LayerNorm (one kernel),
//     self-attention over the beam-indirected cache, cross-attention that reads each utterance's K/V once for all beams.
// The whole chain (+ the search kernels) is captured in one CUDA graph per batch shape by the engine.
// Activations feeding a GEMM are fp16 (tensor-core operands), the residual stream and all reductions are fp32.
#include <cooperative_groups.h>

#include "decoder.cuh"
#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace wisb {

namespace {

constexpr int BD_LN_THREADS = 128;
constexpr int BD_LN_MAX = 12;  // d_model <= 1536

__device__ __forceinline__ float block_sum_128(float v, float* s_red) {
  v = warp_sum(v);
  __syncthreads();  // s_red may still be read from a previous call
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  return s_red[0] + s_red[1] + s_red[2] + s_red[3];
}

// LayerNorm of the block's row held in registers (v[i] = x[tid + 128 i]) -> fp16; two-pass statistics in fp32
__device__ __forceinline__ void row_layernorm_store(const float (&v)[BD_LN_MAX], int iters, int d, const float* __restrict__ g,
                                                    const float* __restrict__ b, __half* __restrict__ out, float* s_red) {
  const int tid = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < BD_LN_MAX; ++i)
    if (i < iters) s += v[i];
  const float mean = block_sum_128(s, s_red) / d;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < BD_LN_MAX; ++i)
    if (i < iters) {
      const float a = v[i] - mean;
      q = fmaf(a, a, q);
    }
  const float rstd = rsqrtf(block_sum_128(q, s_red) / d + 1e-5f);
#pragma unroll
  for (int i = 0; i < BD_LN_MAX; ++i)
    if (i < iters) {
      const int c = tid + i * BD_LN_THREADS;
      out[c] = __float2half_rn((v[i] - mean) * rstd * __ldg(g + c) + __ldg(b + c));
    }
}

// x[r] = tok_emb[token[r]] + pos_emb[row_pos[r]];  xn[r] = LN(x[r])   (first LayerNorm of decoder layer 0)
__global__ void __launch_bounds__(BD_LN_THREADS)
bd_embed_ln_kernel(const int* __restrict__ tokens, const int* __restrict__ row_pos, const __half* __restrict__ tok_emb,
                   const float* __restrict__ pos_emb, const float* __restrict__ g, const float* __restrict__ b,
                   float* __restrict__ x, __half* __restrict__ xn, int d) {
  __shared__ float s_red[4];
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x, tid = threadIdx.x;
  const int iters = d / BD_LN_THREADS;
  const __half* e = tok_emb + static_cast<long long>(tokens[r]) * d;
  const float* p = pos_emb + static_cast<long long>(row_pos[r]) * d;
  float v[BD_LN_MAX];
#pragma unroll
  for (int i = 0; i < BD_LN_MAX; ++i)
    if (i < iters) {
      const int c = tid + i * BD_LN_THREADS;
      v[i] = __half2float(e[c]) + p[c];
      x[static_cast<long long>(r) * d + c] = v[i];
    }
  row_layernorm_store(v, iters, d, g, b, xn + static_cast<long long>(r) * d, s_red);
}

// x[r] += bias + sum_s partial[s][r]  (split-K slabs of the preceding GEMM, fixed summation order);  xn[r] = LN(x[r])
__global__ void __launch_bounds__(BD_LN_THREADS)
bd_resid_ln_kernel(float* __restrict__ x, const float* __restrict__ partial, int n_splits, long long split_stride,
                   const float* __restrict__ bias, const float* __restrict__ g, const float* __restrict__ b,
                   __half* __restrict__ xn, int d) {
  __shared__ float s_red[4];
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.x, tid = threadIdx.x;
  const int iters = d / BD_LN_THREADS;
  float v[BD_LN_MAX];
#pragma unroll
  for (int i = 0; i < BD_LN_MAX; ++i)
    if (i < iters) {
      const int c = tid + i * BD_LN_THREADS;
      const long long at = static_cast<long long>(r) * d + c;
      float acc = __ldg(bias + c);
      for (int s = 0; s < n_splits; ++s) acc += partial[s * split_stride + at];
      v[i] = x[at] + acc;
      x[at] = v[i];
    }
  row_layernorm_store(v, iters, d, g, b, xn + static_cast<long long>(r) * d, s_red);
}

// =====================================================================================================================
// self-attention over the cache: one warp per (row, head).  Position t < pos of row r lives in cache slot indir[r][t]
// (beam reordering by indirection, search.cu), position pos in the row's own slot (written by this pass's QKV GEMM).
// Prefill rows (prompt positions of one utterance, all in the utterance's first slot) attend their own slot only.
// =====================================================================================================================
constexpr int BD_SA_TMAX = 448;

__global__ void __launch_bounds__(128)
bd_self_attn_kernel(const float* __restrict__ q, const __half* __restrict__ kcache, const __half* __restrict__ vcache,
                    const int* __restrict__ row_pos, const int* __restrict__ row_slot, const int* __restrict__ indir0,
                    const int* __restrict__ indir1, const int* __restrict__ flip, const int* __restrict__ done,
                    __half* __restrict__ ctx, int d, int H, int t_cap, int t_ind, int rows_per_utt, int prefill) {
  __shared__ float s_p[4][BD_SA_TMAX];
  __shared__ int s_slot[4][BD_SA_TMAX];
  pdl_launch_dependents();
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x * 4 + warp;
  const int r = blockIdx.y;
  if (h >= H) return;
  if (done != nullptr && done[r / rows_per_utt]) return;  // finished utterance: its rows are dead weight
  const int pos = row_pos[r];
  const int own = row_slot[r];
  const int* indir = (*flip ? indir1 : indir0) + static_cast<long long>(r) * t_ind;
  const float* qr = q + static_cast<long long>(r) * d + h * HEAD_DIM;
  float qv[HEAD_DIM];
#pragma unroll
  for (int i = 0; i < HEAD_DIM / 4; ++i) {
    const float4 v = *reinterpret_cast<const float4*>(qr + 4 * i);
    qv[4 * i] = v.x; qv[4 * i + 1] = v.y; qv[4 * i + 2] = v.z; qv[4 * i + 3] = v.w;
  }
  float mx = -INFINITY;
  for (int t = lane; t <= pos; t += 32) {
    const int slot = (prefill || t == pos) ? own : indir[t];
    s_slot[warp][t] = slot;
    const uint4* kr = reinterpret_cast<const uint4*>(kcache + (static_cast<long long>(slot) * t_cap + t) * d + h * HEAD_DIM);
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint4 u = kr[i];
      const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h2[j]);
        s0 = fmaf(qv[8 * i + 2 * j], f.x, s0);
        s1 = fmaf(qv[8 * i + 2 * j + 1], f.y, s1);
      }
    }
    const float s = (s0 + s1) * 0.125f;
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
    const __half2 v = *reinterpret_cast<const __half2*>(vcache + (static_cast<long long>(s_slot[warp][t]) * t_cap + t) * d +
                                                        h * HEAD_DIM + 2 * lane);
    const float2 f = __half22float2(v);
    o0 = fmaf(p, f.x, o0);
    o1 = fmaf(p, f.y, o1);
  }
  const float inv = 1.0f / sum;
  *reinterpret_cast<__half2*>(ctx + static_cast<long long>(r) * d + h * HEAD_DIM + 2 * lane) = __floats2half2_rn(o0 * inv, o1 * inv);
}

// =====================================================================================================================
// cross-attention: one cluster of 8 CTAs per (head, utterance); CTA c owns keys [192 c, 192 c + 192) of the 1536-row
// padded window (keys >= 1500 are skipped).  Inside a CTA, groups of 8 lanes walk keys with an online softmax for all
// rows of the utterance at once -- the utterance's K/V (251.7 MB over the layers at large-v2) are read ONCE per pass for
// every beam -- partials are merged through shared memory, and the 8 CTAs merge through distributed shared memory.
// Finished utterances are skipped: their K/V are not read at all.
// =====================================================================================================================
constexpr int BD_CA_CLUSTER = 8;
constexpr int BD_CA_KEYS = T_ENC_PAD / BD_CA_CLUSTER;  // 192
constexpr int BD_CA_THREADS = 128;
constexpr int BD_CA_GROUPS = BD_CA_THREADS / 8;        // 16 groups of 8 lanes; 12 keys each

template <int NB>
__global__ void __cluster_dims__(1, 1, BD_CA_CLUSTER) __launch_bounds__(BD_CA_THREADS)
bd_cross_attn_kernel(const float* __restrict__ q, const __half* __restrict__ kmat, const __half* __restrict__ vmat,
                     const int* __restrict__ done, __half* __restrict__ ctx, int rows_per_utt, int d, int H) {
  // dynamic smem: [K tile 192 x 64 fp16 | V tile 192 x 64 fp16 | per-group partial accumulators]
  extern __shared__ __align__(128) uint8_t ca_smem[];
  __half* sK = reinterpret_cast<__half*>(ca_smem);
  __half* sV = sK + BD_CA_KEYS * HEAD_DIM;
  float (*s_acc)[NB][HEAD_DIM] = reinterpret_cast<float (*)[NB][HEAD_DIM]>(ca_smem + 2 * BD_CA_KEYS * HEAD_DIM * 2);
  __shared__ float s_m[BD_CA_GROUPS][NB], s_l[BD_CA_GROUPS][NB];
  __shared__ float c_acc[NB][HEAD_DIM];  // this CTA's merged partial (read by the cluster leader through DSMEM)
  __shared__ float c_m[NB], c_l[NB];
  __shared__ uint64_t s_bar;
  cg::cluster_group cluster = cg::this_cluster();
  const int h = blockIdx.x, u = blockIdx.y, cta = blockIdx.z;
  const int tid = threadIdx.x;
  pdl_launch_dependents();
  const int grp = tid >> 3, gl = tid & 7;  // lane gl of group grp owns dims [8 gl, 8 gl + 8)
  const long long head_off = (static_cast<long long>(u) * H + h) * T_ENC_PAD * HEAD_DIM;
  const int t_begin = cta * BD_CA_KEYS;
  const uint32_t bar = smem_u32(&s_bar);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  pdl_wait();  // q (and the `done` flags) come from the previous kernels of the chain
  if (done != nullptr && done[u]) return;  // uniform over the cluster: no CTA of it reaches cluster.sync()
  // one elected thread streams this CTA's 192 keys and values (2 x 24 KB, contiguous) into shared memory with TMA
  if (tid == 0) {
    mbar_arrive_expect_tx(bar, 2u * BD_CA_KEYS * HEAD_DIM * 2u);
    bulk_load_1d(smem_u32(sK), kmat + head_off + static_cast<long long>(t_begin) * HEAD_DIM, BD_CA_KEYS * HEAD_DIM * 2, bar);
    bulk_load_1d(smem_u32(sV), vmat + head_off + static_cast<long long>(t_begin) * HEAD_DIM, BD_CA_KEYS * HEAD_DIM * 2, bar);
  }
  float qv[NB][8];
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    if (k < rows_per_utt) {
      const float* qr = q + static_cast<long long>(u * rows_per_utt + k) * d + h * HEAD_DIM + gl * 8;
      const float4 a0 = *reinterpret_cast<const float4*>(qr), a1 = *reinterpret_cast<const float4*>(qr + 4);
      qv[k][0] = a0.x * 0.125f; qv[k][1] = a0.y * 0.125f; qv[k][2] = a0.z * 0.125f; qv[k][3] = a0.w * 0.125f;
      qv[k][4] = a1.x * 0.125f; qv[k][5] = a1.y * 0.125f; qv[k][6] = a1.z * 0.125f; qv[k][7] = a1.w * 0.125f;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) qv[k][i] = 0.f;
    }
  }
  __syncthreads();  // barrier init visible
  mbar_wait(bar, 0);
  float m[NB], l[NB], acc[NB][8];
#pragma unroll
  for (int k = 0; k < NB; ++k) {
    m[k] = -INFINITY;
    l[k] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
  }
  const unsigned gmask = 0xFFu << (tid & 24);  // the 8 lanes of this group (shuffles stay inside it)
#pragma unroll 1
  for (int tl = grp; tl < BD_CA_KEYS; tl += BD_CA_GROUPS) {
    if (t_begin + tl >= T_ENC) break;  // uniform inside the 8-lane group
    const uint4 ku = *reinterpret_cast<const uint4*>(sK + tl * HEAD_DIM + gl * 8);
    const uint4 vu = *reinterpret_cast<const uint4*>(sV + tl * HEAD_DIM + gl * 8);
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
  for (int idx = tid; idx < NB * HEAD_DIM; idx += BD_CA_THREADS) {
    const int k = idx / HEAD_DIM, e = idx % HEAD_DIM;
    float mm = -INFINITY;
    for (int g = 0; g < BD_CA_GROUPS; ++g) mm = fmaxf(mm, s_m[g][k]);
    float a = 0.f, ll = 0.f;
    for (int g = 0; g < BD_CA_GROUPS; ++g) {
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
    for (int idx = tid; idx < NB * HEAD_DIM; idx += BD_CA_THREADS) {
      const int k = idx / HEAD_DIM, e = idx % HEAD_DIM;
      if (k >= rows_per_utt) continue;
      float mm = -INFINITY;
      for (int c = 0; c < BD_CA_CLUSTER; ++c) mm = fmaxf(mm, *cluster.map_shared_rank(&c_m[k], c));
      float a = 0.f, ll = 0.f;
      for (int c = 0; c < BD_CA_CLUSTER; ++c) {
        const float mc = *cluster.map_shared_rank(&c_m[k], c);
        const float w = (mc == -INFINITY) ? 0.f : __expf(mc - mm);
        a = fmaf(w, *cluster.map_shared_rank(&c_acc[k][e], c), a);
        ll = fmaf(w, *cluster.map_shared_rank(&c_l[k], c), ll);
      }
      ctx[static_cast<long long>(u * rows_per_utt + k) * d + h * HEAD_DIM + e] = __float2half_rn(a / ll);
    }
  }
  cluster.sync();  // keep every CTA's shared memory alive until the leader has read it
}

template <typename Kern, typename... Args>
void bd_launch(Kern kern, dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = pdl ? 1 : 0;
  WISB_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
}

int bd_ca_smem(int nb) { return 2 * BD_CA_KEYS * HEAD_DIM * 2 + BD_CA_GROUPS * nb * HEAD_DIM * 4; }

void cross_attn_launch(const BatchArgs& a, const BatchLayer& ly, cudaStream_t s) {
  static std::atomic<unsigned long long> once{0};
  once_per_device(once, [&] {
    WISB_CUDA(cudaFuncSetAttribute(bd_cross_attn_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bd_ca_smem(1)));
    WISB_CUDA(cudaFuncSetAttribute(bd_cross_attn_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, bd_ca_smem(5)));
    WISB_CUDA(cudaFuncSetAttribute(bd_cross_attn_kernel<MAX_BEAM>, cudaFuncAttributeMaxDynamicSharedMemorySize, bd_ca_smem(MAX_BEAM)));
  });
  dim3 grid(a.H, a.n_utt, BD_CA_CLUSTER);
  const int rpu = a.rows_per_utt;
  if (rpu == 1)
    bd_launch(bd_cross_attn_kernel<1>, grid, dim3(BD_CA_THREADS), bd_ca_smem(1), s, a.pdl != 0, a.q, ly.ck, ly.cv, a.done, a.ctx, rpu, a.d, a.H);
  else if (rpu <= 5)
    bd_launch(bd_cross_attn_kernel<5>, grid, dim3(BD_CA_THREADS), bd_ca_smem(5), s, a.pdl != 0, a.q, ly.ck, ly.cv, a.done, a.ctx, rpu, a.d, a.H);
  else
    bd_launch(bd_cross_attn_kernel<MAX_BEAM>, grid, dim3(BD_CA_THREADS), bd_ca_smem(MAX_BEAM), s, a.pdl != 0, a.q, ly.ck, ly.cv, a.done, a.ctx, rpu, a.d, a.H);
}

void run_gemm_rows(const GemmPlan& plan, int rows, int pdl, cudaStream_t s) {
  GemmPlan p = plan;  // the plan is built for the buffer capacity; this pass uses the first `rows` rows
  p.M = round_up(rows, 128);
  if (p.M > plan.M) p.M = plan.M;
  p.epi.m_valid = rows;
  p.pdl = pdl;
  gemm_run(p, s);
}

}  // namespace

int batch_pass_run(const BatchArgs& a, const BatchLayer* layers, int n_layers, cudaStream_t s) {
  WISB_REQUIRE(a.R >= 1 && a.R == a.n_utt * a.rows_per_utt, "batched decoder pass: rows must be utterances x rows per utterance");
  WISB_REQUIRE(a.rows_per_utt >= 1 && a.rows_per_utt <= MAX_BEAM, "batched decoder pass: 1..8 rows per utterance");
  WISB_REQUIRE(a.d % BD_LN_THREADS == 0 && a.d <= BD_LN_THREADS * BD_LN_MAX, "batched decoder pass: d_model multiple of 128, <= 1536");
  WISB_REQUIRE(a.t_ind <= BD_SA_TMAX, "batched decoder pass: more than 448 text positions");
  const bool pdl = a.pdl != 0;
  int n = 0;
  bd_launch(bd_embed_ln_kernel, dim3(a.R), dim3(BD_LN_THREADS), 0, s, pdl, a.tokens, a.row_pos, a.tok_emb, a.pos_emb,
            layers[0].ln1g, layers[0].ln1b, a.x, a.xn, a.d);
  ++n;
  for (int i = 0; i < n_layers; ++i) {
    const BatchLayer& ly = layers[i];
    run_gemm_rows(ly.qkv, a.R, a.pdl, s);
    bd_launch(bd_self_attn_kernel, dim3(cdiv(a.H, 4), a.R), dim3(128), 0, s, pdl, a.q, ly.kcache, ly.vcache, a.row_pos,
              a.row_slot, a.indir0, a.indir1, a.flip, a.done, a.ctx, a.d, a.H, a.t_cap, a.t_ind, a.rows_per_utt, a.prefill);
    run_gemm_rows(ly.o, a.R, a.pdl, s);
    bd_launch(bd_resid_ln_kernel, dim3(a.R), dim3(BD_LN_THREADS), 0, s, pdl, a.x, a.part, ly.o.k_splits, a.part_stride,
              ly.ob, ly.ln2g, ly.ln2b, a.xn, a.d);
    run_gemm_rows(ly.cq, a.R, a.pdl, s);
    cross_attn_launch(a, ly, s);
    run_gemm_rows(ly.co, a.R, a.pdl, s);
    bd_launch(bd_resid_ln_kernel, dim3(a.R), dim3(BD_LN_THREADS), 0, s, pdl, a.x, a.part, ly.co.k_splits, a.part_stride,
              ly.cob, ly.ln3g, ly.ln3b, a.xn, a.d);
    run_gemm_rows(ly.fc1, a.R, a.pdl, s);
    run_gemm_rows(ly.fc2, a.R, a.pdl, s);
    bd_launch(bd_resid_ln_kernel, dim3(a.R), dim3(BD_LN_THREADS), 0, s, pdl, a.x, a.part, ly.fc2.k_splits, a.part_stride,
              ly.fc2b, ly.next_g, ly.next_b, a.xn, a.d);
    n += 11;
  }
  if (a.with_logits) {
    run_gemm_rows(*a.vocab, a.R, a.pdl, s);
    ++n;
  }
  return n;
}

}  // namespace wisb
