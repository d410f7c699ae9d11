// WavTokenizer-decoder kernels other than the GEMMs (a3, a11-a13).  Activations are channels-last
// (rows = frames, 768 contiguous channels) in a PADDED RAGGED layout: chunk i owns rows
// [row0_i, row0_i + L_i) with row0_i a multiple of 8; at least ROW_PAD zero rows precede the first chunk and follow every
// chunk, so the k=3 / k=7
// convolutions (as multi-tap GEMMs) read zeros at chunk edges exactly like the reference's per-chunk
// zero padding.  row_chunk[r] = chunk index of row r, or -1 for a padding row.
#pragma once
#include "common.cuh"

namespace lvx {

constexpr int ROW_PAD = 3;

struct ChunkInfo {
  int row0;  // first row of the chunk in the padded layout
  int len;   // frames
  int out0;  // first frame in the packed (unpadded) order == h_cu[i]
  int s_off; // unused (scores live in the launch group's [padded rows][sld] matrices)
};

// ---------------------------------------------------------------------------------------------------
// GroupNorm statistics (WavTokenizer/decoder/models.py:15-16: 32 groups, eps 1e-6): mean / rstd over
// (C/32 channels x L frames) per (chunk, group).  One CTA per (group, chunk); one shifted pass.
// ---------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) groupnorm_stats_kernel(const float* __restrict__ x,
                                                              const ChunkInfo* __restrict__ chunks, float eps,
                                                              float2* __restrict__ stats) {
  constexpr int G = 32, CPG = C / G, V = CPG / 4;
  __shared__ float red[32];
  const int g = blockIdx.x, ch = blockIdx.y;
  const ChunkInfo ci = chunks[ch];
  const int items = ci.len * V;
  const float* base = x + (size_t)ci.row0 * C + g * CPG;
  // ONE pass over the strip: sums of (x - K) and (x - K)^2 with K = the strip's first element (a value of the same
  // distribution, so |mean - K| is a few standard deviations at most and the cancellation in q - s^2 / n costs a few ulps;
  // the plain two-pass form read the strip twice, the second time from L2)
  const float K = base[0];
  float s = 0.f, q = 0.f;
  for (int i = threadIdx.x; i < items; i += 256) {
    const float4 v = load4(base + (size_t)(i / V) * C + (i % V) * 4);
    const float a = v.x - K, b = v.y - K, c = v.z - K, d = v.w - K;
    s += (a + b) + (c + d);
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float n = (float)(ci.len * CPG);
  const float ds = block_sum(s, red) / n;
  const float mean = K + ds;
  const float var = fmaxf(block_sum(q, red) / n - ds * ds, 0.f);
  if (threadIdx.x == 0) stats[(size_t)ch * G + g] = make_float2(mean, 1.0f / sqrtf(var + eps));
}

// GroupNorm apply (+ optional swish, models.py:10-12) -> GEMM operand type.  One warp per row (6 float4 per lane, no
// index divisions; swish through the fast exponential: relative error 2^-21).
template <typename TOut, int C>
__global__ void __launch_bounds__(256) groupnorm_apply_kernel(const float* __restrict__ x, int rows,
                                                              const int* __restrict__ row_chunk,
                                                              const float2* __restrict__ stats,
                                                              const float* __restrict__ w, const float* __restrict__ b,
                                                              int swish, TOut* __restrict__ out) {
  constexpr int G = 32, CPG = C / G, V = C / 128;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int ch = row_chunk[row];
  if (ch < 0) {  // padding row: the k=3 conv that follows must read zeros here
#pragma unroll
    for (int i = 0; i < V; ++i) store4(out + (size_t)row * C + (lane + 32 * i) * 4, make_float4(0.f, 0.f, 0.f, 0.f));
    return;
  }
  float4 v[V];
#pragma unroll
  for (int i = 0; i < V; ++i) v[i] = load4(x + (size_t)row * C + (lane + 32 * i) * 4);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float2 st = stats[(size_t)ch * G + c / CPG];
    const float4 ww = load4(w + c), bb = load4(b + c);
    float r[4] = {(v[i].x - st.x) * st.y * ww.x + bb.x, (v[i].y - st.x) * st.y * ww.y + bb.y,
                  (v[i].z - st.x) * st.y * ww.z + bb.z, (v[i].w - st.x) * st.y * ww.w + bb.w};
    if (swish) {
#pragma unroll
      for (int k = 0; k < 4; ++k) r[k] = r[k] / (1.0f + __expf(-r[k]));
    }
    store4(out + (size_t)row * C + c, make_float4(r[0], r[1], r[2], r[3]));
  }
}

// ---------------------------------------------------------------------------------------------------
// pos_net tail + backbone.norm (models.py:213,226-228): GroupNorm apply (affine) then AdaLayerNorm
// (LN eps 1e-6 no affine, * scale[bw] + shift[bw]; modules.py:81-86) -> new fp32 residual stream.
// One warp per row.
// ---------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) groupnorm_adaln_kernel(const float* __restrict__ x, int rows,
                                                              const int* __restrict__ row_chunk,
                                                              const float2* __restrict__ stats,
                                                              const float* __restrict__ gw, const float* __restrict__ gb,
                                                              const float* __restrict__ scale,
                                                              const float* __restrict__ shift, float eps,
                                                              float* __restrict__ out_gn, float* __restrict__ out) {
  constexpr int G = 32, CPG = C / G, V = C / 128;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int ch = row_chunk[row];
  if (ch < 0) return;
  float4 v[V];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float2 st = stats[(size_t)ch * G + c / CPG];
    const float4 a = load4(x + (size_t)row * C + c), ww = load4(gw + c), bb = load4(gb + c);
    v[i] = make_float4((a.x - st.x) * st.y * ww.x + bb.x, (a.y - st.x) * st.y * ww.y + bb.y,
                       (a.z - st.x) * st.y * ww.z + bb.z, (a.w - st.x) * st.y * ww.w + bb.w);
    if (out_gn) store4(out_gn + (size_t)row * C + c, v[i]);
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  const float mean = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += a * a + b * b + c * c + d * d;
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / C) + eps);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float4 sc = load4(scale + c), sh = load4(shift + c);
    store4(out + (size_t)row * C + c,
           make_float4((v[i].x - mean) * rstd * sc.x + sh.x, (v[i].y - mean) * rstd * sc.y + sh.y,
                       (v[i].z - mean) * rstd * sc.z + sh.z, (v[i].w - mean) * rstd * sc.w + sh.w));
  }
}

// ---------------------------------------------------------------------------------------------------
// ConvNeXt front half (modules.py:45-51): depthwise Conv1d k=7 pad 3 (zero padding at CHUNK edges) +
// bias, then AdaLayerNorm.  One warp per frame; the 7 neighbour rows come through L1/L2.
// dw weights are stored tap-major [7][C].
// ---------------------------------------------------------------------------------------------------
template <typename TOut, int C>
__global__ void __launch_bounds__(256) dwconv_adaln_kernel(const float* __restrict__ x, int rows,
                                                           const int* __restrict__ row_chunk,
                                                           const ChunkInfo* __restrict__ chunks,
                                                           const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                                                           const float* __restrict__ scale,
                                                           const float* __restrict__ shift, float eps,
                                                           TOut* __restrict__ out) {
  constexpr int V = C / 128;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int ch = row_chunk[row];
  if (ch < 0) return;
  const ChunkInfo ci = chunks[ch];
  float4 v[V];
#pragma unroll
  for (int i = 0; i < V; ++i) v[i] = load4(dw_b + (lane + 32 * i) * 4);
#pragma unroll
  for (int t = 0; t < 7; ++t) {
    const int r = row + t - 3;
    if (r < ci.row0 || r >= ci.row0 + ci.len) continue;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = (lane + 32 * i) * 4;
      const float4 a = load4(x + (size_t)r * C + c), w = load4(dw_w + t * C + c);
      v[i].x = fmaf(a.x, w.x, v[i].x);
      v[i].y = fmaf(a.y, w.y, v[i].y);
      v[i].z = fmaf(a.z, w.z, v[i].z);
      v[i].w = fmaf(a.w, w.w, v[i].w);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) s += v[i].x + v[i].y + v[i].z + v[i].w;
  const float mean = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += a * a + b * b + c * c + d * d;
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / C) + eps);
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float4 sc = load4(scale + c), sh = load4(shift + c);
    store4(out + (size_t)row * C + c,
           make_float4((v[i].x - mean) * rstd * sc.x + sh.x, (v[i].y - mean) * rstd * sc.y + sh.y,
                       (v[i].z - mean) * rstd * sc.z + sh.z, (v[i].w - mean) * rstd * sc.w + sh.w));
  }
}

// ---------------------------------------------------------------------------------------------------
// Row softmax of the pos_net attention scores (models.py:117-118; the C^-0.5 scale is the GEMM's alpha).
// Scores of padded row r: S[r * sld .. + L) fp32 (sld = the launch group's longest chunk rounded up to 64); P gets the
// same layout, with columns [L, L rounded up to 64) written as zeros so the P.V GEMM's last k-block may read them.  One
// warp per row.  Output type = GEMM operand type, in place for fp32.
// ---------------------------------------------------------------------------------------------------
// bf16x3 operand builders of the iSTFT head (the head and iDFT GEMMs keep ~fp32 accuracy on bf16 tensor cores by laying
// hi | lo | hi thirds side by side along K): out[r, 0:seg) = hi, [seg, 2 seg) = lo, [2 seg, 3 seg) = hi.
//   layernorm_split3_kernel : backbone final LayerNorm (models.py:229) straight into the head GEMM's operand -- one warp per
//                             row; the fp32 copy the separate split kernel re-read is gone
//   head_act_split3_kernel  : ISTFTHead activation (heads.py:57-63: mag = min(exp(.), 100), S = mag (cos p + i sin p)) straight
//                             into the iDFT GEMM's operand -- one thread per frequency bin computes exp and sincos ONCE for the
//                             real and the imaginary part (the per-output-element kernel computed exp twice and went
//                             through an fp32 spectrum: 0.46 + 0.23 ms per 61,440 frames)
// Padding rows become zeros (the GEMMs mask them on the way out).
// ---------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(256) layernorm_split3_kernel(const float* __restrict__ x, int rows, const float* __restrict__ w,
                                                               const float* __restrict__ bias, float eps,
                                                               const int* __restrict__ row_chunk, bf16* __restrict__ out) {
  constexpr int V = C / 128;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  bf16* o = out + (size_t)row * (3 * C);
  if (row_chunk[row] < 0) {
    for (int c = lane * 8; c < 3 * C; c += 256) *reinterpret_cast<uint4*>(o + c) = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
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
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int c = (lane + 32 * i) * 4;
    const float4 ww = load4(w + c), bb = load4(bias + c);
    const float r[4] = {(v[i].x - mean) * rstd * ww.x + bb.x, (v[i].y - mean) * rstd * ww.y + bb.y,
                        (v[i].z - mean) * rstd * ww.z + bb.z, (v[i].w - mean) * rstd * ww.w + bb.w};
    __nv_bfloat162 h[2], l[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      h[k] = __floats2bfloat162_rn(r[2 * k], r[2 * k + 1]);
      const float2 hf = __bfloat1622float2(h[k]);
      l[k] = __floats2bfloat162_rn(r[2 * k] - hf.x, r[2 * k + 1] - hf.y);
    }
    const uint2 hu = make_uint2(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]));
    const uint2 lu = make_uint2(*reinterpret_cast<uint32_t*>(&l[0]), *reinterpret_cast<uint32_t*>(&l[1]));
    *reinterpret_cast<uint2*>(o + c) = hu;
    *reinterpret_cast<uint2*>(o + C + c) = lu;
    *reinterpret_cast<uint2*>(o + 2 * C + c) = hu;
  }
}
// exp / sincos of the operand builder below.  Its output is a hi | lo bf16 pair (16 mantissa bits, 2^-17), so the accurate
// libdevice routines (149 instructions per bin with them: the kernel was issue-bound at 57 % of the issue slots) buy
// nothing: ex2.approx on x log2(e) (relative error ~1e-6 for |x| < 5; larger magnitudes clip at 100) and sin / cos.approx after a
// two-term Cody-Waite reduction to [-pi, pi] (absolute error 2^-21.4); phases beyond 8192 rad take the libdevice path.
__device__ __forceinline__ void sincos_reduced(float p, float* sn, float* cs) {
  if (fabsf(p) > 8192.0f) {
    sincosf(p, sn, cs);
    return;
  }
  const float k = rintf(p * 0.15915494309189535f);
  float r = fmaf(k, -6.2831854820251465f, p);   // fp32(2 pi) ...
  r = fmaf(k, 1.7484555e-7f, r);                // ... minus the 1.7484555e-7 it exceeds 2 pi by
  *sn = __sinf(r);
  *cs = __cosf(r);
}
// One WARP per row, 8 rows per CTA: a lane loads the magnitude / phase pairs of 8 bins before it touches any (16 loads in flight
// per lane -- with one bin per thread the kernel sat at 2.5 TB/s whatever its arithmetic and store pattern were: too few bytes in
// flight per SM), the hi and lo halves of the row's operand are built in the warp's slice of shared memory and go out as whole
// 16-byte pieces.  (One 2-byte store per element, with the imaginary parts starting at the odd column 641, also left every warp
// store a 64-byte run across three partly written sectors: 405 MB of DRAM reads for 315 MB of input under ncu.)
__global__ void __launch_bounds__(256) head_act_split3_kernel(const float* __restrict__ raw, int ld_raw, int rows,
                                                              const int* __restrict__ row_chunk, int bins, bf16* __restrict__ out,
                                                              int seg) {
  extern __shared__ __align__(16) bf16 has_rows[];   // per warp: hi[seg] | lo[seg]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;
  if (row >= rows) return;
  bf16* hi = has_rows + (size_t)warp * 2 * seg;
  bf16* lo = hi + seg;
  uint4* o4 = reinterpret_cast<uint4*>(out + (size_t)row * (3 * seg));
  const int n4 = seg / 8;
  if (row_chunk[row] < 0) {
    for (int i = lane; i < 3 * n4; i += 32) o4[i] = make_uint4(0u, 0u, 0u, 0u);
    return;
  }
  const float* r = raw + (size_t)row * ld_raw;
  for (int k0 = 0; k0 < bins; k0 += 256) {
    float m[8], ph[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + 32 * j + lane;
      m[j] = k < bins ? r[k] : 0.f;
      ph[j] = k < bins ? r[bins + k] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + 32 * j + lane;
      if (k >= bins) continue;
      const float mag = fminf(__expf(m[j]), 100.0f);
      float sn, cs;
      sincos_reduced(ph[j], &sn, &cs);
      const float re = mag * cs, im = mag * sn;
      const bf16 rh = __float2bfloat16_rn(re), ih = __float2bfloat16_rn(im);
      hi[k] = rh;
      lo[k] = __float2bfloat16_rn(re - __bfloat162float(rh));
      hi[bins + k] = ih;
      lo[bins + k] = __float2bfloat16_rn(im - __bfloat162float(ih));
    }
  }
  for (int k = 2 * bins + lane; k < seg; k += 32) hi[k] = lo[k] = __float2bfloat16_rn(0.f);   // alignment columns
  __syncwarp();
  const uint4* h4 = reinterpret_cast<const uint4*>(hi);
  const uint4* l4 = reinterpret_cast<const uint4*>(lo);
  for (int i = lane; i < n4; i += 32) {
    const uint4 h = h4[i];
    o4[i] = h;
    o4[n4 + i] = l4[i];
    o4[2 * n4 + i] = h;
  }
}

// ---------------------------------------------------------------------------------------------------
// V^T operand of the pos_net attention's P V GEMM: Vt[c][r] = V[r][c], V = columns [v0, v0 + C) of the fused q|k|v
// output (bf16, row stride ld).  64 x 64 tiles through shared memory; padding rows become zeros (the grouped P V GEMM
// reads its K tail up to the next multiple of 64 rows, against P columns that are zero there).  Replaces a second
// projection GEMM with a transposed store (0.57 ms per 61,440 frames at 128 TFLOP/s; this copy: 0.05 ms).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) transpose_v_kernel(const bf16* __restrict__ qkv, int ld, int v0, int rows,
                                                          const int* __restrict__ row_chunk, bf16* __restrict__ vt, int ldt) {
  __shared__ __align__(4) bf16 tile[64][66];
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int item = threadIdx.x + 256 * it, r = item >> 3, ch = item & 7;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (r0 + r < rows && row_chunk[r0 + r] >= 0) u = *reinterpret_cast<const uint4*>(qkv + (size_t)(r0 + r) * ld + v0 + c0 + 8 * ch);
    uint32_t* d = reinterpret_cast<uint32_t*>(&tile[r][8 * ch]);
    d[0] = u.x; d[1] = u.y; d[2] = u.z; d[3] = u.w;
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int item = threadIdx.x + 256 * it, c = item >> 3, rb = (item & 7) * 8;
    if (r0 + rb >= rows) continue;   // the last 8-row piece may run up to 7 columns past `rows` (zeros; ldt has the room)
    __align__(16) bf16 o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = tile[rb + k][c];
    *reinterpret_cast<uint4*>(vt + (size_t)(c0 + c) * ldt + r0 + rb) = *reinterpret_cast<const uint4*>(o);
  }
}

// ---------------------------------------------------------------------------------------------------
// The same operation in strips (the product path at every batch size).  The one-warp-per-frame kernel above re-reads every
// input row 7 times and the 21 KB of depthwise weights once per frame through L1: ~42 KB of L1 traffic per frame, which
// bounds it at a third of the HBM rate (136 us per launch at 61,440 frames against 43 us of HBM time).  Here a CTA of 6
// warps owns a strip of DWB_R frames; warp w owns channels [128 w, +128) with its 7 x 4 weights in REGISTERS and slides a
// 7-row window down the strip (28 FMAs per row), leaving the convolution in shared memory; after one barrier the warps
// normalise the strip's rows (AdaLayerNorm over all 768 channels) and store.  Chunk edges: at least ROW_PAD = 3 padding
// rows separate chunks, so a window never reaches another chunk's frames -- padding rows count as zeros (the reference's
// per-chunk zero padding), whatever the residual stream holds there.
//
// A first version pulled the rows through registers (two sets of 8 prefetched rows, or every row costs a full HBM latency:
// 193 us per launch without the prefetch, 111 with): 139 registers, 12 warps per SM, 0.44 of the HBM rate under ncu.
// Now one thread issues cp.async.bulk copies of the whole strip (DWB_R + 6 contiguous rows)
// into shared memory and the convolution runs IN PLACE from there (row o is stored to slot o once slot o + 6 has been read;
// a thread only ever touches its own 4 channels), which frees the registers: short strips of 12 frames (54 KB) put four
// CTAs on an SM, so three strips are in flight while one computes.  The 1.5x re-read of the halo rows hits L2.  Measured on
// 61,440 frames (engine profiler, per launch incl. its events): 121 us registers -> 91 us (30 rows, 2 CTAs) -> 84 us
// (12 rows, 4 CTAs); it also beats the one-warp-per-frame kernel on the small streaming batches (64 x 90 frames: 23 -> 15 us).
// Same summation order as the kernel above (tests/test_gpu_parity.py compares the two bit for bit).
// ---------------------------------------------------------------------------------------------------
constexpr int DWB_R = 12;
constexpr int DWB_SMEM = (DWB_R + 6) * 768 * 4 + 64 * 4 + 16;
template <typename TOut, int C>
__global__ void __launch_bounds__(192, 4) dwconv_adaln_bulk_kernel(const float* __restrict__ x, int rows,
                                                                   const int* __restrict__ row_chunk,
                                                                   const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                                                                   const float* __restrict__ scale,
                                                                   const float* __restrict__ shift, float eps,
                                                                   TOut* __restrict__ out) {
  static_assert(C == 768, "6 warps x 128 channels");
  constexpr int TR = DWB_R + 6;
  extern __shared__ __align__(128) float dwb_tile[];   // [TR][C], then TR validity flags, then the mbarrier
  int* valid = reinterpret_cast<int*>(dwb_tile + TR * C);
  const uint32_t bar = smem_u32(dwb_tile + TR * C + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = warp * 128 + lane * 4;
  const int r0 = blockIdx.x * DWB_R;
  const int g_lo = max(r0 - 3, 0), g_hi = min(r0 + DWB_R + 3, rows);   // rows of x that exist
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < TR) {
    const int r = r0 - 3 + (int)threadIdx.x;
    valid[threadIdx.x] = (r >= 0 && r < rows && row_chunk[r] >= 0) ? 1 : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, (uint32_t)(g_hi - g_lo) * C * 4);
    for (int r = g_lo; r < g_hi; r += 9) {   // a few copies rather than one: they spread over the copy engine's queues
      const int nr = min(9, g_hi - r);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(dwb_tile + (size_t)(r - (r0 - 3)) * C)),
                   "l"(x + (size_t)r * C), "r"((uint32_t)nr * C * 4), "r"(bar)
                   : "memory");
    }
  }
  float4 wt[7];
#pragma unroll
  for (int t = 0; t < 7; ++t) wt[t] = load4(dw_w + t * C + c0);
  const float4 b4 = load4(dw_b + c0);
  mbar_wait(bar, 0);
  auto slot = [&](int i) -> float4 {   // padding rows and rows outside the batch count as zeros (their slots hold anything)
    return valid[i] ? *reinterpret_cast<const float4*>(dwb_tile + i * C + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
  };
  float4 win[7];
#pragma unroll
  for (int t = 0; t < 6; ++t) win[t] = slot(t);
#pragma unroll 3
  for (int o = 0; o < DWB_R; ++o) {
    win[6] = slot(o + 6);
    float4 v = b4;
#pragma unroll
    for (int t = 0; t < 7; ++t) {
      v.x = fmaf(win[t].x, wt[t].x, v.x);
      v.y = fmaf(win[t].y, wt[t].y, v.y);
      v.z = fmaf(win[t].z, wt[t].z, v.z);
      v.w = fmaf(win[t].w, wt[t].w, v.w);
    }
    *reinterpret_cast<float4*>(dwb_tile + o * C + c0) = v;
#pragma unroll
    for (int t = 0; t < 6; ++t) win[t] = win[t + 1];
  }
  __syncthreads();
  constexpr int V = C / 128;
#pragma unroll 1
  for (int o = warp; o < DWB_R; o += 6) {
    const int row = r0 + o;
    if (!valid[o + 3]) continue;   // warp-uniform
    const float* src = dwb_tile + o * C;
    float4 v[V];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      v[i] = *reinterpret_cast<const float4*>(src + (lane + 32 * i) * 4);
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      q += a * a + b * b + c * c + d * d;
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / C) + eps);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int c = (lane + 32 * i) * 4;
      const float4 sc = load4(scale + c), sh = load4(shift + c);
      store4(out + (size_t)row * C + c,
             make_float4((v[i].x - mean) * rstd * sc.x + sh.x, (v[i].y - mean) * rstd * sc.y + sh.y,
                         (v[i].z - mean) * rstd * sc.z + sh.z, (v[i].w - mean) * rstd * sc.w + sh.w));
    }
  }
}

// ---------------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void __launch_bounds__(256) attn_softmax_kernel(const float* __restrict__ S, TOut* __restrict__ P,
                                                           const ChunkInfo* __restrict__ chunks,
                                                           const int* __restrict__ row_chunk, int rows, int sld) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int ch = row_chunk[row];
  if (ch < 0) return;
  const ChunkInfo ci = chunks[ch];
  const int L = ci.len, Lp = min(sld, (L + 63) & ~63);
  const size_t off = (size_t)row * sld;
  const float* s = S + off;
  if constexpr (sizeof(TOut) == 2) {
    // bf16 destination, rows of up to 1280 scores: the row lives in registers (10 float4 per lane) -- one read instead of three,
    // one exponential per score (fast ex2: the result is rounded to bf16) instead of two; the three-pass form was issue-bound (72 %)
    if (sld <= 1280) {
      float4 v[10];
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        const int j = 4 * (lane + 32 * i);
        v[i] = j < Lp ? load4(s + j) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        if (j + 0 >= L) v[i].x = -INFINITY;
        if (j + 1 >= L) v[i].y = -INFINITY;
        if (j + 2 >= L) v[i].z = -INFINITY;
        if (j + 3 >= L) v[i].w = -INFINITY;
        mx = fmaxf(mx, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
      }
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        v[i].x = __expf(v[i].x - mx); v[i].y = __expf(v[i].y - mx); v[i].z = __expf(v[i].z - mx); v[i].w = __expf(v[i].w - mx);
        sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      }
      const float inv = 1.0f / warp_sum(sum);
#pragma unroll
      for (int i = 0; i < 10; ++i) {
        const int j = 4 * (lane + 32 * i);
        if (j < Lp) store4(P + off + j, make_float4(v[i].x * inv, v[i].y * inv, v[i].z * inv, v[i].w * inv));
      }
      return;
    }
  }
  float mx = -INFINITY;
  for (int j = lane; j < L; j += 32) mx = fmaxf(mx, s[j]);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < L; j += 32) sum += expf(s[j] - mx);
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  TOut* p = P + off;
  for (int j = lane; j < Lp; j += 32) store1(p + j, j < L ? expf(s[j] - mx) * inv : 0.f);
}

// ---------------------------------------------------------------------------------------------------
// ISTFT head activation (heads.py:54-65): columns [0,641) = log-magnitude, [641,1282) = phase ->
// S = min(exp(m), 100) * (cos p, sin p), written as the iDFT GEMM's A operand
// [re_0..re_640 | im_0..im_640 | zero pad to ldo].
// ---------------------------------------------------------------------------------------------------
template <typename TOut>
__global__ void __launch_bounds__(256) head_activation_kernel(const float* __restrict__ raw, int ld_raw, int rows,
                                                              const int* __restrict__ row_chunk, int bins,
                                                              TOut* __restrict__ out, int ldo) {
  const int row = blockIdx.x;   // rows on grid.x: a launch group may hold more than 65535 rows
  if (row >= rows || row_chunk[row] < 0) return;
  const int j = blockIdx.y * 256 + threadIdx.x;
  if (j >= ldo) return;
  const float* r = raw + (size_t)row * ld_raw;
  float v = 0.f;
  if (j < 2 * bins) {
    const int k = j < bins ? j : j - bins;
    const float mag = fminf(expf(r[k]), 100.0f);
    const float ph = r[bins + k];
    v = mag * (j < bins ? cosf(ph) : sinf(ph));
  }
  store1(out + (size_t)row * ldo + j, v);
}

// ---------------------------------------------------------------------------------------------------
// Overlap-add + "same" trim + window-envelope normalisation (spectral_ops.py:59-73).  frames[r, n] already
// carries the Hann window (folded into the iDFT basis).  Output sample t of a chunk sits at padded position
// p = t + (n_fft - hop) / 2 and sums frames i with i*hop <= p < i*hop + n_fft, 0 <= i < L; the envelope is
// the same sum over window^2.  One thread per output sample.
// ---------------------------------------------------------------------------------------------------
template <int V>   // samples per thread: 4 (16-byte accesses; needs a 16-byte aligned PCM pointer) or 1
__global__ void __launch_bounds__(256) overlap_add_kernel(const float* __restrict__ frames, int ldf,
                                                          const ChunkInfo* __restrict__ chunks,
                                                          const float* __restrict__ window, int n_fft, int hop,
                                                          float* __restrict__ pcm) {
  // V consecutive samples per thread: hop, n_fft and the trim are multiples of 4, so the samples of a thread share their
  // frames; per sample the same sums in the same order whatever V is
  const ChunkInfo ci = chunks[blockIdx.y];
  const int t = (blockIdx.x * 256 + threadIdx.x) * V;
  if (t >= ci.len * hop) return;
  const int p = t + (n_fft - hop) / 2;
  int i_lo = (p - n_fft + hop) / hop;  // ceil((p - n_fft + 1) / hop) for p - n_fft + 1 > 0
  if (p - n_fft + 1 <= 0) i_lo = 0;
  int i_hi = p / hop;
  if (i_hi > ci.len - 1) i_hi = ci.len - 1;
  float y[V], env[V];
#pragma unroll
  for (int j = 0; j < V; ++j) y[j] = env[j] = 0.f;
  for (int i = i_lo; i <= i_hi; ++i) {
    const int n = p - i * hop;
    float w[V], f[V];
    if constexpr (V == 4) {
      *reinterpret_cast<float4*>(w) = load4(window + n);
      *reinterpret_cast<float4*>(f) = load4(frames + (size_t)(ci.row0 + i) * ldf + n);
    } else {
      w[0] = window[n];
      f[0] = frames[(size_t)(ci.row0 + i) * ldf + n];
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      y[j] += f[j];
      env[j] += w[j] * w[j];
    }
  }
  float* o = pcm + (size_t)ci.out0 * hop + t;
  if constexpr (V == 4) store4(o, make_float4(y[0] / env[0], y[1] / env[1], y[2] / env[2], y[3] / env[3]));
  else o[0] = y[0] / env[0];
}

__global__ void build_row_chunk_kernel(const ChunkInfo* __restrict__ chunks, int n_chunks, int* __restrict__ row_chunk,
                                       int* __restrict__ code_rows) {
  const ChunkInfo ci = chunks[blockIdx.x];
  for (int j = threadIdx.x; j < ci.len; j += blockDim.x) {
    row_chunk[ci.row0 + j] = blockIdx.x;
    code_rows[ci.out0 + j] = ci.row0 + j;
  }
}

// copy valid rows of a padded activation to packed order (test hook)
__global__ void unpad_rows_kernel(const float* __restrict__ src, int ld, int width, const int* __restrict__ code_rows,
                                  int n, float* __restrict__ dst) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const float* s = src + (size_t)code_rows[i] * ld;
  for (int c = threadIdx.x; c < width; c += blockDim.x) dst[(size_t)i * width + c] = s[c];
}
template <typename T>
__global__ void unpad_rows_typed_kernel(const T* __restrict__ src, int ld, int width, const int* __restrict__ code_rows,
                                        int n, float* __restrict__ dst) {
  const int i = blockIdx.x;
  if (i >= n) return;
  const T* s = src + (size_t)code_rows[i] * ld;
  for (int c = threadIdx.x; c < width; c += blockDim.x) dst[(size_t)i * width + c] = load1(s + c);
}

}  // namespace lvx
