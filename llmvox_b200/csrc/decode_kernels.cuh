// Decode-step kernels other than the GEMMs: input assembly (a2-a4), LayerNorm (a8), flash-decode attention
// over the paged KV cache (a6), the sampler (a9).  Reference lines are cited per kernel.
#pragma once
#include "common.cuh"

namespace lvx {

// Per-engine device-side session state (all indexed by slot).
struct SessionState {
  int* ctx_len;      // [S] codes decoded so far == KV tokens held == next position
  int* text_len;     // [S]
  int* text_ids;     // [S, max_context]
  int* codes;        // [S, max_context] code history
  int* page_table;   // [S, max_pages]
  int* eoa_pos;      // [S] position of the first end-of-audio code (eoa_id) of the sentence, -1 = none yet
  int max_context, max_pages, eoa_id;
};

// ---------------------------------------------------------------------------------------------------
// streaming_server.py:313-334 + src/model.py:206-212: gather text row (256) and previous-code row (512,
// zeros at step 0), L2-normalise with eps 1e-8 (F.normalize divides by max(norm, eps)), add wpe[t].
// One CTA (n_embd / 4 threads, one float4 each) per session.  Writes the fp32 residual stream x[b, 768].
// ---------------------------------------------------------------------------------------------------
__global__ void assemble_input_kernel(const int* __restrict__ slots, SessionState st,
                                                             const float* __restrict__ text_table,
                                                             const float* __restrict__ codebook,
                                                             const float* __restrict__ wpe, int text_dim, int code_dim,
                                                             int pad_id, int step_offset, float* __restrict__ x) {
  __shared__ float red[32];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, slot = slots[b];
  const int t = st.ctx_len[slot] + step_offset;
  const int C = text_dim + code_dim;
  const int c = threadIdx.x * 4;
  int tid = pad_id;
  if (t < st.text_len[slot]) tid = st.text_ids[(size_t)slot * st.max_context + t];
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (c < C) {
    if (c < text_dim) {
      v = load4(text_table + (size_t)tid * text_dim + c);
    } else if (t > 0) {
      const int prev = st.codes[(size_t)slot * st.max_context + t - 1];
      v = load4(codebook + (size_t)prev * code_dim + (c - text_dim));
    }
  }
  const float ss = block_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w, red);
  const float denom = fmaxf(sqrtf(ss), 1e-8f);
  if (c < C) {
    const float4 pe = load4(wpe + (size_t)t * C + c);
    store4(x + (size_t)b * C + c, make_float4(v.x / denom + pe.x, v.y / denom + pe.y, v.z / denom + pe.z,
                                               v.w / denom + pe.w));
  }
}

// Drop-in path (`model(emb, kvcache)`): x = emb + wpe[pos]  (src/model.py:206-212).
__global__ void add_wpe_kernel(const float* __restrict__ emb, const int* __restrict__ pos,
                               const float* __restrict__ wpe, int C, float* __restrict__ x) {
  const int b = blockIdx.x;
  for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
    const float4 e = load4(emb + (size_t)b * C + c), w = load4(wpe + (size_t)pos[b] * C + c);
    store4(x + (size_t)b * C + c, make_float4(e.x + w.x, e.y + w.y, e.z + w.z, e.w + w.w));
  }
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm over the last dim (src/model.py:29-38, eps 1e-5; vocoder LayerNorms use eps 1e-6).  One warp
// per row, C = 768 held in registers (6 float4 per lane), two-pass mean / biased variance like ATen.
// Optional affine weight / bias (AdaLayerNorm passes scale / shift rows; modules.py:81-86).
// Rows with row_chunk[r] < 0 (padding rows of the vocoder layout) are skipped.
// ---------------------------------------------------------------------------------------------------
template <typename TOut, int C>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, int rows,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        float eps, const int* __restrict__ row_chunk,
                                                        TOut* __restrict__ out) {
  constexpr int V = C / 128;  // float4 per lane
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  if (row_chunk && row_chunk[row] < 0) return;
  const float* xr = x + (size_t)row * C;
  float4 v[V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i] = load4(xr + (lane + 32 * i) * 4);
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += a * a + b * b + c * c + d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
  TOut* o = out + (size_t)row * C;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = (lane + 32 * i) * 4;
    float4 r = make_float4((v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd,
                           (v[i].w - mean) * rstd);
    if (w) {
      const float4 ww = load4(w + c);
      r.x *= ww.x; r.y *= ww.y; r.z *= ww.z; r.w *= ww.w;
    }
    if (bias) {
      const float4 bb = load4(bias + c);
      r.x += bb.x; r.y += bb.y; r.z += bb.z; r.w += bb.w;
    }
    store4(o + c, r);
  }
}

// ---------------------------------------------------------------------------------------------------
// Exact mode (LVX_PRECISION_EXACT) operand builders.  A GEMM operand row of width W is stored as 2 W bf16 values
// [hi(0..W) | lo(0..W)], hi = bf16(v), lo = bf16(v - hi): against the weight row laid out twice along K the fp32
// accumulator receives w.hi + w.lo = w.v up to 2^-18 |v| per element -- fp32-class activations on bf16 tensor cores.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_hi_lo(float v, bf16& hi, bf16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
// LayerNorm without affine terms (the weight is folded into the GEMM that follows) -> hi | lo.  One warp per row.
template <int C>
__global__ void __launch_bounds__(256) ln_split_kernel(const float* __restrict__ x, int rows, float eps, bf16* __restrict__ out) {
  constexpr int V = C / 128;
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * C;
  float4 v[V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    v[i] = load4(xr + (lane + 32 * i) * 4);
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += a * a + b * b + c * c + d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
  bf16* o = out + (size_t)row * 2 * C;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float r[4] = {(v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd, (v[i].w - mean) * rstd};
    bf16 hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_hi_lo(r[j], hi[j], lo[j]);
    *reinterpret_cast<uint2*>(o + c) = *reinterpret_cast<uint2*>(hi);
    *reinterpret_cast<uint2*>(o + C + c) = *reinterpret_cast<uint2*>(lo);
  }
}
// out[r, 0:W) | out[r, W:2W) = hi | lo of act(in[r, :]); act = ACT_NONE or ACT_GELU_TANH (src/model.py:21-26, exact tanhf)
__global__ void __launch_bounds__(256) split2_kernel(const float* __restrict__ in, int rows, int W, int act, bf16* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t n4 = (size_t)rows * W / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = (i * 4) / W, c = (i * 4) % W;
    const float4 v4 = load4(in + i * 4);
    float v[4] = {v4.x, v4.y, v4.z, v4.w};
    bf16 hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (act == ACT_GELU_TANH) v[j] = gelu_tanh(v[j]);
      split_hi_lo(v[j], hi[j], lo[j]);
    }
    bf16* o = out + r * 2 * W;
    *reinterpret_cast<uint2*>(o + c) = *reinterpret_cast<uint2*>(hi);
    *reinterpret_cast<uint2*>(o + W + c) = *reinterpret_cast<uint2*>(lo);
  }
}
// (N, K) fp32 weight, optionally scaled per column (the LayerNorm weight in front of the GEMM), rounded to bf16 and
// laid out twice along K: out[n, 0:K) = out[n, K:2K) = bf16(W[n, :] * scale)
__global__ void fold_dup_kernel(const float* __restrict__ Wt, const float* __restrict__ scale, int N, int K, int ld,
                                bf16* __restrict__ out) {
  const size_t n = (size_t)N * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / K, c = i % K;
    const bf16 v = __float2bfloat16_rn(Wt[r * ld + c] * (scale ? scale[c] : 1.0f));
    out[r * 2 * K + c] = v;
    out[r * 2 * K + K + c] = v;
  }
}

// ---------------------------------------------------------------------------------------------------
// Flash-decode attention over the paged KV cache (src/model.py:68-98 with is_train=False: one query row
// against [cache ; new row], no mask, scale 1/sqrt(head_dim)).  One CTA per (session, head); 4 warps x 4
// groups of 8 lanes; a group owns one cached token at a time and each lane HD/8 of its dims.  Online
// softmax per group, merged through shared memory.  The new token's K/V come from the qkv GEMM output and
// are appended to the cache here (the O(1) replacement of the reference's torch.cat, model.py:74-77).
//
// KV pool layout: [layer][k|v][page][head][tile of page_tokens x HD elements], element type TKV.  Inside a tile the
// bf16 pool is token-major ([token][dim]).  The fp32 pool (fp32 and exact precisions) is LANE-MAJOR per 8-token half page:
// [half][k = 0..5][lane = 8 part + t8][4 floats] with token = 8 half + t8 and dim = 24 part + 4 k + e, so that the six
// fully coalesced 512-byte loads a warp of the cluster kernel issues per half page land directly in the lane that
// multiplies them (lane (t8, part) owns 24 dims of token t8: no shared-memory staging, cluster_decode.cuh).
// ---------------------------------------------------------------------------------------------------
template <typename TKV, int HD>
__device__ __forceinline__ uint32_t kv_tile_offset(int t, int d) {
  if constexpr (sizeof(TKV) == 4) {
    static_assert(HD == 96, "lane-major fp32 KV tiles are laid out for head_dim 96");
    const int part = d / 24, r = d - 24 * part;
    return (uint32_t)((t >> 3) * (8 * HD) + (r >> 2) * 128 + ((part << 3) + (t & 7)) * 4 + (r & 3));
  } else {
    return (uint32_t)(t * HD + d);
  }
}
template <typename TKV, typename TOut, int HD>
__global__ void __launch_bounds__(128, 8) decode_attention_kernel(const float* __restrict__ qkv, TKV* __restrict__ kv,
                                                               const int* __restrict__ slots, SessionState st,
                                                               const int* __restrict__ pos_override, int layer,
                                                               int n_head, int page_tokens, long long pool_pages,
                                                               int step_offset, TOut* __restrict__ y) {
  constexpr int DPL = HD / 8;  // dims per lane
  static_assert(DPL % 4 == 0, "head_dim must be a multiple of 32");
  constexpr int NG = 16;       // token groups per CTA
  __shared__ float sm_m[NG], sm_l[NG];
  __shared__ __align__(16) float sm_acc[NG][HD];
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, h = blockIdx.y, slot = slots[b];
  const int C = n_head * HD;
  const int T = pos_override ? pos_override[b] : st.ctx_len[slot] + step_offset;  // cached tokens; new token index
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = warp * 4 + (lane >> 3), sub = lane & 7;
  const float scale = rsqrtf((float)HD);

  const float* qrow = qkv + (size_t)b * 3 * C + h * HD + sub * DPL;
  float q[DPL];
#pragma unroll
  for (int i = 0; i < DPL; i += 4) {
    const float4 a = load4(qrow + i);
    q[i] = a.x; q[i + 1] = a.y; q[i + 2] = a.z; q[i + 3] = a.w;
  }
  // the new token's k / v (rounded as they will be read back from the cache) are NOT kept in registers across the token
  // loop: they are re-read from the qkv row (L2) where they are needed.  79 -> 56 registers = 8 instead of 6 CTAs per SM;
  // the loop is latency-bound (60 % long_scoreboard stalls at 2048 sessions: ncu), so occupancy is bandwidth.
  auto new_kv = [&](float (&kn)[DPL], float (&vn)[DPL]) {
#pragma unroll
    for (int i = 0; i < DPL; i += 4) {
      const float4 k4 = load4(qrow + C + i), v4 = load4(qrow + 2 * C + i);
      kn[i] = round_to<TKV>(k4.x); kn[i + 1] = round_to<TKV>(k4.y); kn[i + 2] = round_to<TKV>(k4.z); kn[i + 3] = round_to<TKV>(k4.w);
      vn[i] = round_to<TKV>(v4.x); vn[i + 1] = round_to<TKV>(v4.y); vn[i + 2] = round_to<TKV>(v4.z); vn[i + 3] = round_to<TKV>(v4.w);
    }
  };
  const size_t head_stride = (size_t)page_tokens * HD;
  const size_t page_stride = (size_t)n_head * head_stride;
  TKV* kbase = kv + ((size_t)(layer * 2 + 0) * pool_pages) * page_stride + (size_t)h * head_stride;
  TKV* vbase = kv + ((size_t)(layer * 2 + 1) * pool_pages) * page_stride + (size_t)h * head_stride;
  const int* pt = st.page_table + (size_t)slot * st.max_pages;

  // append the new token (group 0 holds all HD dims across its 8 lanes)
  if (g == 0) {
    float kn[DPL], vn[DPL];
    new_kv(kn, vn);
    const int page = pt[T / page_tokens], off = T % page_tokens;
    TKV* kd = kbase + (size_t)page * page_stride;
    TKV* vd = vbase + (size_t)page * page_stride;
#pragma unroll
    for (int i = 0; i < DPL; i += 4) {
      const uint32_t o = kv_tile_offset<TKV, HD>(off, sub * DPL + i);
      store4(kd + o, make_float4(kn[i], kn[i + 1], kn[i + 2], kn[i + 3]));
      store4(vd + o, make_float4(vn[i], vn[i + 1], vn[i + 2], vn[i + 3]));
    }
  }

  float m = -INFINITY, l = 0.f, acc[DPL];
#pragma unroll
  for (int i = 0; i < DPL; ++i) acc[i] = 0.f;

  // the 8-lane groups of a warp run different trip counts: shuffles are confined to the group's own lanes
  const unsigned gmask = 0xFFu << (lane & 24);
  auto absorb = [&](const float* kk, const float* vv) {
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < DPL; ++i) d = fmaf(q[i], kk[i], d);
    d += __shfl_xor_sync(gmask, d, 4);
    d += __shfl_xor_sync(gmask, d, 2);
    d += __shfl_xor_sync(gmask, d, 1);
    d *= scale;
    const float mn = fmaxf(m, d);
    const float corr = expf(m - mn), pr = expf(d - mn);
    l = l * corr + pr;
#pragma unroll
    for (int i = 0; i < DPL; ++i) acc[i] = acc[i] * corr + pr * vv[i];
    m = mn;
  };

  for (int tok = g; tok < T; tok += NG) {   // (unrolled by two: 381 vs 369 us per launch at 2048 sessions, T = 300)
    const int page = pt[tok / page_tokens], off = tok % page_tokens;
    const TKV* kp = kbase + (size_t)page * page_stride;
    const TKV* vp = vbase + (size_t)page * page_stride;
    float kk[DPL], vv[DPL];
#pragma unroll
    for (int i = 0; i < DPL; i += 4) {
      const uint32_t o = kv_tile_offset<TKV, HD>(off, sub * DPL + i);
      const float4 a = load4(kp + o), c = load4(vp + o);
      kk[i] = a.x; kk[i + 1] = a.y; kk[i + 2] = a.z; kk[i + 3] = a.w;
      vv[i] = c.x; vv[i + 1] = c.y; vv[i + 2] = c.z; vv[i + 3] = c.w;
    }
    absorb(kk, vv);
  }
  if (g == (T % NG)) {   // the new token, taken by the group that would own index T
    float kn[DPL], vn[DPL];
    new_kv(kn, vn);
    absorb(kn, vn);
  }

  if (sub == 0) { sm_m[g] = m; sm_l[g] = l; }
#pragma unroll
  for (int i = 0; i < DPL; ++i) sm_acc[g][sub * DPL + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < HD) {
    float M = -INFINITY;
#pragma unroll
    for (int i = 0; i < NG; ++i) M = fmaxf(M, sm_m[i]);
    float L = 0.f, o = 0.f;
#pragma unroll
    for (int i = 0; i < NG; ++i) {
      const float w = (sm_m[i] == -INFINITY) ? 0.f : expf(sm_m[i] - M);
      L += w * sm_l[i];
      o += w * sm_acc[i][threadIdx.x];
    }
    store1(y + (size_t)b * C + h * HD + threadIdx.x, o / L);
  }
}

// ---------------------------------------------------------------------------------------------------
// Sampler (a9).  One CTA (256 threads) per session over V logits (V <= 16 * 256 * ... generic loop).
// Greedy: argmax, lowest index wins ties (== torch.argmax of softmax, streaming_server.py:342-346).
// Sampled: src/model.py:397-406 -- /temperature, keep >= k-th largest (exact radix select, ties kept),
// softmax, inverse-CDF draw in index order.  Appends the code to the session history, bumps ctx_len.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t float_order_key(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // larger float -> larger key
}

__device__ __forceinline__ void philox_round(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0,
                                             uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0, hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
  const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}
// Philox4x32-10; returns a uniform in [0, 1) with 24 bits.
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint32_t slot, uint32_t step) {
  uint32_t c0 = slot, c1 = step, c2 = 0, c3 = 0;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c0, c1, c2, c3, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return (float)(c0 >> 8) * (1.0f / 16777216.0f);
}

struct SamplerArgs {
  int greedy, top_k;
  float temperature;
  uint64_t seed;
  const float* uniform;     // optional [n]
  const int* forced;        // optional [n]: value stored in the history instead of the pick
  int* out_codes;           // optional [n]
  int record;               // append to history + bump ctx_len
};

// Scratch of one sampling problem (256 cooperating threads).
struct SamplerScratch {
  float red[32];
  int redi[32];
  unsigned int hist[256];
  float seg_sum[256];
  unsigned int sel_prefix, sel_remaining;
  int pick;
};
// One pick over vals[0, V) (shared memory, V logits of one session; overwritten) by 256 threads (tid in [0, 256)) that
// meet at `sync()`: the block barrier of sampler_kernel, or the worker-warps barrier of the cluster-resident kernel.
// Greedy: argmax, lowest index wins ties.  Sampled (src/model.py:397-406): / temperature, keep >= k-th largest (exact
// 4-pass radix select, ties kept), softmax, inverse CDF in index order against the uniform u.
template <typename Sync>
__device__ __forceinline__ int sample_pick(float* vals, int V, bool greedy, int top_k, float temperature, float u, SamplerScratch& S,
                                           int tid, Sync sync) {
  if (greedy || !(temperature > 0.f)) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < V; i += 256) {
      const float v = vals[i];
      if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { S.red[tid >> 5] = best; S.redi[tid >> 5] = bi; }
    sync();
    if (tid < 32) {
      best = tid < 8 ? S.red[tid] : -INFINITY;
      bi = tid < 8 ? S.redi[tid] : 0x7fffffff;
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (tid == 0) S.pick = (bi == 0x7fffffff) ? 0 : bi;
    }
    sync();
    return S.pick;
  }
  // logits / temperature (model.py:398)
  for (int i = tid; i < V; i += 256) vals[i] = vals[i] / temperature;
  sync();
  float thresh = -INFINITY;
  if (top_k > 0 && top_k < V) {
    // exact k-th largest by 4-pass MSB radix select over order-preserving keys
    if (tid == 0) { S.sel_prefix = 0; S.sel_remaining = (unsigned)top_k; }
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      S.hist[tid] = 0;
      sync();
      const unsigned prefix = S.sel_prefix;
      const unsigned himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
      for (int i = tid; i < V; i += 256) {
        const unsigned key = float_order_key(vals[i]);
        if ((key & himask) == (prefix & himask)) atomicAdd(&S.hist[(key >> shift) & 255u], 1u);
      }
      sync();
      if (tid == 0) {
        unsigned rem = S.sel_remaining;
        int d = 255;
        for (; d > 0; --d) {
          if (S.hist[d] >= rem) break;
          rem -= S.hist[d];
        }
        S.sel_prefix = prefix | ((unsigned)d << shift);
        S.sel_remaining = rem;
      }
      sync();
    }
    const unsigned key = S.sel_prefix;
    const unsigned uu = (key & 0x80000000u) ? (key & 0x7fffffffu) : ~key;
    thresh = __uint_as_float(uu);
  }
  // softmax over survivors (model.py:401-404), then inverse CDF in index order
  float mx = -INFINITY;
  for (int i = tid; i < V; i += 256) mx = fmaxf(mx, vals[i]);
  mx = warp_max(mx);
  if ((tid & 31) == 0) S.red[tid >> 5] = mx;
  sync();
  mx = S.red[0];
#pragma unroll
  for (int w = 1; w < 8; ++w) mx = fmaxf(mx, S.red[w]);
  // thread owns the contiguous segment [tid*seg, (tid+1)*seg)
  const int seg = (V + 255) / 256;
  float ssum = 0.f;
  for (int i = tid * seg; i < min(V, (tid + 1) * seg); ++i) {
    const float v = vals[i];
    const float e = (v >= thresh) ? expf(v - mx) : 0.f;
    vals[i] = e;
    ssum += e;
  }
  S.seg_sum[tid] = ssum;
  sync();
  if (tid == 0) {
    float total = 0.f;
    for (int i = 0; i < 256; ++i) total += S.seg_sum[i];
    const float target = u * total;
    float run = 0.f;
    int found = -1, last = 0;
    for (int s = 0; s < 256 && found < 0; ++s) {
      if (run + S.seg_sum[s] > target) {
        for (int i = s * seg; i < min(V, (s + 1) * seg); ++i) {
          if (vals[i] > 0.f) last = i;
          run += vals[i];
          if (run > target) { found = i; break; }
        }
        if (found < 0) continue;  // rounding: fall through to the next segment
      } else {
        run += S.seg_sum[s];
        if (S.seg_sum[s] > 0.f) {
          for (int i = min(V, (s + 1) * seg) - 1; i >= s * seg; --i)
            if (vals[i] > 0.f) { last = i; break; }
        }
      }
    }
    S.pick = found >= 0 ? found : last;
  }
  sync();
  return S.pick;
}

struct BlockSync {
  __device__ __forceinline__ void operator()() const { __syncthreads(); }
};

template <int MAXV>
__global__ void __launch_bounds__(256) sampler_kernel(const float* __restrict__ logits, int V,
                                                      const int* __restrict__ slots, SessionState st,
                                                      SamplerArgs a) {
  __shared__ float vals[MAXV];
  __shared__ SamplerScratch S;
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.x, slot = slots[b], tid = threadIdx.x;
  const float* lg = logits + (size_t)b * V;
  for (int i = tid; i < V; i += 256) vals[i] = lg[i];
  __syncthreads();
  const bool greedy = a.greedy != 0;
  float u = 0.f;
  if (!greedy) u = a.uniform ? a.uniform[b] : philox_uniform(a.seed, (uint32_t)slot, (uint32_t)st.ctx_len[slot]);
  const int pick = sample_pick(vals, V, greedy, a.top_k, a.temperature, u, S, tid, BlockSync());

  if (tid == 0) {
    if (a.out_codes) a.out_codes[b] = pick;
    if (a.record) {
      const int t = st.ctx_len[slot];
      const int rec = a.forced ? a.forced[b] : pick;
      st.codes[(size_t)slot * st.max_context + t] = rec;
      st.ctx_len[slot] = t + 1;
      if (rec == st.eoa_id && st.eoa_pos[slot] < 0) st.eoa_pos[slot] = t;   // streaming_server.py:379, 397
    }
  }
}

// out[i] = (eoa_pos[slot_i], ctx_len[slot_i]): the only per-round facts the host's chunk scheduler needs (:357-422)
__global__ void gather_progress_kernel(const int* __restrict__ slots, int n, SessionState st, int2* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_int2(st.eoa_pos[slots[i]], st.ctx_len[slots[i]]);
}

// history gather: out[i, j] = codes[slot_i, start + j]
__global__ void gather_codes_kernel(const int* __restrict__ slots, SessionState st, int start, int count,
                                    int* __restrict__ out) {
  const int b = blockIdx.x, slot = slots[b];
  for (int j = threadIdx.x; j < count; j += blockDim.x)
    out[(size_t)b * count + j] = st.codes[(size_t)slot * st.max_context + start + j];
}

__global__ void scatter_text_kernel(const int* __restrict__ slots, const int* __restrict__ offsets,
                                    const int* __restrict__ ids, SessionState st) {
  const int b = blockIdx.x, slot = slots[b];
  const int n = offsets[b + 1] - offsets[b], base = st.text_len[slot];
  for (int j = threadIdx.x; j < n; j += blockDim.x)
    st.text_ids[(size_t)slot * st.max_context + base + j] = ids[offsets[b] + j];
  __syncthreads();
  if (threadIdx.x == 0) st.text_len[slot] = base + n;
}

__global__ void reset_sessions_kernel(const int* __restrict__ slots, int n, SessionState st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    st.ctx_len[slots[i]] = 0;
    st.text_len[slots[i]] = 0;
    st.eoa_pos[slots[i]] = -1;
  }
}

// generic row gather: out[i, :] = table[idx[i], :]   (codes_to_features / llm_model drop-ins, a2 / a3)
template <typename TOut>
__global__ void gather_rows_kernel(const int* __restrict__ idx, const float* __restrict__ table, int width, int n,
                                   TOut* __restrict__ out, int ld_out, const int* __restrict__ dst_rows) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const int r = dst_rows ? dst_rows[i] : i;
  const float* src = table + (size_t)idx[i] * width;
  for (int c = threadIdx.x * 4; c < width; c += blockDim.x * 4) store4(out + (size_t)r * ld_out + c, load4(src + c));
}

}  // namespace lvx
