// Persistent fused decode kernel (bf16 mode, greedy pick): ONE cooperative launch runs `n_iters` whole decode
// iterations (streaming_server.py:323-354 -> src/model.py:201-237) for n <= 128 sessions.
//
// Why: a decode iteration is ~35 dependent, latency-bound kernels (~5 us each even for one session, measured), and
// across streams the GPU's kernel-launch rate becomes the limit.  Here the kernel boundaries become grid barriers
// (~1.5 us), barrier/TMEM/descriptor set-up is paid once per launch instead of once per GEMM, and no launch happens
// per token at all.
//
// Grid = n_clusters x 4 CTAs (clusters of 4 = split-K ranks), 192 threads, one CTA per SM, all co-resident
// (cooperative launch).  Phases of one iteration, each followed by a grid barrier:
//   rows    : assemble input + LN1 of layer 0                        (CTA b = session b)
//   per layer: QKV GEMM | attention (items (b, head) strided over CTAs) | proj GEMM (+x) | LN2 rows |
//              fc GEMM + tanh-GELU | proj2 GEMM (+x) | next LN rows
//   lm_head GEMM | greedy pick rows (argmax, lowest index wins; appends the code, bumps ctx_len)
// GEMM phases are the swap-mode tcgen05 tiles of tc_gemm.cuh: weight rows fill UMMA M = 128, sessions are UMMA N,
// K is split over the 4 CTAs of a cluster and reduced through distributed shared memory in fixed rank order.
// Data produced by one phase and consumed by another CTA is read through L2 only (TMA, ld.global.cg).
#pragma once
#include <cooperative_groups.h>

#include "decode_kernels.cuh"
#include "tc_gemm.cuh"

namespace lvx {

constexpr int FD_CLUSTER = 4;
constexpr int FD_MAX_LAYERS = 8;
constexpr int FD_NG = 24;  // attention token groups per CTA (6 warps x 4 groups of 8 lanes)

struct FusedGemm {
  int map_a, map_w;       // indices into FusedParams::maps
  int N, K;               // weight rows (output features), reduction length
  const float* bias;      // [N] or null
  const float* residual;  // fp32 [n, ldc] or null (aliases C for the residual GEMMs)
  void* C;
  int ldc, c_bf16, act;
};

struct FusedLayer {
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  FusedGemm qkv, proj, fc, proj2;
};

struct FusedParams {
  int n, n_iters, BN, stages, tmem_cols, n_clusters, n_layer, n_head, C, vocab;
  const CUtensorMap* maps;
  const int* slots;
  SessionState st;
  const float *text_table, *codebook, *wpe;
  int text_dim, code_dim, pad_id;
  FusedLayer layer[FD_MAX_LAYERS];
  const float *lnf_w, *lnf_b;
  FusedGemm lm_head;
  float *x, *qkv, *logits;
  bf16 *h, *y;
  bf16* kv;
  int page_tokens;
  long long pool_pages;
  unsigned* bar;  // grid-barrier counter, zeroed before the launch
  long long* trace;  // optional: CTA 0 stamps clock64() at every phase boundary of the last iteration (test hook)
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Sense-free grid barrier on a monotonically increasing counter (epoch * blocks).  Writers' generic-proxy stores are
// ordered before later TMA (async-proxy) reads by fence.proxy.async on both sides; the leader's gpu-scope fences give
// release / acquire and drop stale L1 lines.  The spin is bounded: a bug ends in a trap, not in a hung GPU.
__device__ __forceinline__ void fd_grid_sync(unsigned* counter, unsigned& epoch, unsigned nblocks) {
  asm volatile("fence.proxy.async;" ::: "memory");
  __syncthreads();
  epoch += 1;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned target = epoch * nblocks;
    const long long t0 = clock64();
    while (ld_acquire_u32(counter) < target) {
      if (clock64() - t0 > 4000000000LL) __trap();
    }
    __threadfence();
    asm volatile("fence.proxy.async;" ::: "memory");
  }
  __syncthreads();
}

// DSMEM load without a memory clobber: consecutive loads stay in flight together (ordering against the cluster
// barrier is kept because both are volatile asm statements)
__device__ __forceinline__ float4 ld_dsmem_v4_nc(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra));
  return v;
}
// tanh-GELU (src/model.py:21-26) with the hardware tanh approximation (MUFU.TANH, rel. error ~5e-4: below the bf16
// rounding applied to this value right after)
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float k = 0.7978845608028654f;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(k * (x + 0.044715f * x * x * x)));
  return 0.5f * x * (1.0f + t);
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldcg4(const bf16* p) {
  const uint2 u = __ldcg(reinterpret_cast<const uint2*>(p));
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x);
  const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}

// LayerNorm of one 768-wide row held as one float4 per thread (192 threads), two-pass like ATen; bf16 out.
__device__ __forceinline__ void fd_ln_row(float4 v, const float* __restrict__ w, const float* __restrict__ b, float eps,
                                          bf16* __restrict__ out, float* red, int C) {
  const int c = threadIdx.x * 4;
  const float mean = block_sum(v.x + v.y + v.z + v.w, red) / (float)C;
  const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
  const float rstd = rsqrtf(block_sum(dx * dx + dy * dy + dz * dz + dw * dw, red) / (float)C + eps);
  const float4 ww = load4(w + c);
  float4 r = make_float4(dx * rstd * ww.x, dy * rstd * ww.y, dz * rstd * ww.z, dw * rstd * ww.w);
  if (b) {
    const float4 bb = load4(b + c);
    r.x += bb.x; r.y += bb.y; r.z += bb.z; r.w += bb.w;
  }
  store4(out + c, r);
}

// ---- one swap-mode GEMM phase: tiles of 128 weight rows strided over the clusters
struct FdPipe {
  uint32_t tiles, bar0, tmem_d;
  uint32_t red;        // shared-memory address of the split-K partial tile (outside the operand ring)
  int it_total;    // k-blocks issued so far in this kernel (ring position / phase of the smem pipeline)
  int tile_total;  // tiles finished so far (phase of the tmem-full barrier)
};

__device__ __noinline__ void fd_gemm_phase(const FusedParams& P, const FusedGemm& G, FdPipe& pp, uint8_t* smem_raw, int cid,
                                              int rank, long long* tr = nullptr) {
#define FD_T(i) do { if (tr) tr[i] = clock64(); } while (0)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = P.BN, stages = P.stages, S = FD_CLUSTER;
  // P and G live in kernel-parameter space behind references: copy what the loops need into registers once
  // (otherwise every use is a load that misses L1 after each barrier)
  const int gN = G.N, gK = G.K, g_ldc = G.ldc, g_act = G.act, g_bf16 = G.c_bf16, n_sess = P.n, n_clusters = P.n_clusters;
  const float* g_bias = G.bias;
  const float* g_res = G.residual;
  void* g_C = G.C;
  const uint32_t stage_bytes = TC_X_BYTES + (uint32_t)BN * TC_BK * 2;
  auto full_bar = [&](int s) { return pp.bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return pp.bar0 + 8u * (TC_MAX_STAGES + s); };
  const uint32_t tmem_full_bar = pp.bar0 + 8u * (2 * TC_MAX_STAGES);
  const int num_kb = gK / TC_BK;
  const int kb0 = num_kb * rank / S, kb1 = num_kb * (rank + 1) / S, nk = kb1 - kb0;
  const int n_tiles = ceil_div(gN, TC_BM);
  const int RS = BN + 4;
  const int q = warp & 3, drow = q * 32 + lane;
  const uint32_t trow = pp.tmem_d + ((uint32_t)(q * 32) << 16);
  const CUtensorMap* mapX = P.maps + G.map_w;
  const CUtensorMap* mapY = P.maps + G.map_a;

  for (int tile = cid; tile < n_tiles; tile += n_clusters) {
    const int x0 = tile * TC_BM;
    if (warp == 0) {
      if (lane == 0) {
        for (int it = 0; it < nk; ++it) {
          const int gi = pp.it_total + it, s = gi % stages;
          const uint32_t ph = (uint32_t)(gi / stages) & 1u;
          const uint32_t dst = pp.tiles + (uint32_t)s * stage_bytes;
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), stage_bytes);
          tma_load_2d(mapX, full_bar(s), dst, (kb0 + it) * TC_BK, x0);
          tma_load_2d(mapY, full_bar(s), dst + TC_X_BYTES, (kb0 + it) * TC_BK, 0);
        }
        FD_T(0);
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = umma_idesc_bf16(TC_BM, BN);
        for (int it = 0; it < nk; ++it) {
          const int gi = pp.it_total + it, s = gi % stages;
          const uint32_t ph = (uint32_t)(gi / stages) & 1u;
          mbar_wait(full_bar(s), ph);
          if (it == 0) FD_T(1);
          tc_fence_after();
          const uint32_t xs = pp.tiles + (uint32_t)s * stage_bytes;
          const uint64_t adesc = umma_smem_desc(xs), bdesc = umma_smem_desc(xs + TC_X_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            umma_bf16(pp.tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
          umma_commit(empty_bar(s));
        }
        umma_commit(tmem_full_bar);
        FD_T(2);
      }
    } else {
      mbar_wait(tmem_full_bar, (uint32_t)pp.tile_total & 1u);
      __syncwarp();   // reconverge before the .sync.aligned tcgen05.ld
      if (threadIdx.x == 64) FD_T(3);
      tc_fence_after();
      float* red = reinterpret_cast<float*>(smem_raw + (pp.red - smem_u32(smem_raw)));
      for (int c = 0; c < BN; c += 16) {
        float v[16];
        tmem_ld16(trow + (uint32_t)c, v);
        float* dst = red + (size_t)drow * RS + c;
#pragma unroll
        for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
    if (threadIdx.x == 64) FD_T(4);
    pp.it_total += nk;
    pp.tile_total += 1;
    // split-K reduction through distributed shared memory, fixed rank order; rank r finishes BN / 4 columns
    tc_fence_before();
    __syncwarp();
    cluster_sync_all();
    if (threadIdx.x == 64) FD_T(5);
    if (warp >= 2) {
      // thread = weight row nrow (output feature); this rank finishes session columns [rank * cw, (rank + 1) * cw)
      const int cw = BN / S, nrow = x0 + drow;
      const float bv = (g_bias && nrow < gN) ? g_bias[nrow] : 0.f;
      float* Cf = reinterpret_cast<float*>(g_C);
      bf16* Cb = reinterpret_cast<bf16*>(g_C);
#pragma unroll 1
      for (int c0 = rank * cw; c0 < (rank + 1) * cw; c0 += 16) {
        const int nc = min(16, (rank + 1) * cw - c0);   // multiple of 4
        float4 t[4][FD_CLUSTER];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          if (4 * q4 < nc) {
            const uint32_t off = pp.red + (uint32_t)(((size_t)drow * RS + c0 + 4 * q4) * 4);
#pragma unroll
            for (int r = 0; r < FD_CLUSTER; ++r) t[q4][r] = ld_dsmem_v4_nc(off, (uint32_t)r);   // up to 16 loads in flight
          }
        }
        float rs[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          rs[j] = (g_res && nrow < gN && j < nc && c0 + j < n_sess) ? __ldcg(g_res + (size_t)(c0 + j) * g_ldc + nrow) : 0.f;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          if (4 * q4 < nc && nrow < gN) {
            float a[4] = {t[q4][0].x, t[q4][0].y, t[q4][0].z, t[q4][0].w};
#pragma unroll
            for (int r = 1; r < FD_CLUSTER; ++r) {   // fixed rank order: deterministic
              a[0] += t[q4][r].x; a[1] += t[q4][r].y; a[2] += t[q4][r].z; a[3] += t[q4][r].w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int m = c0 + 4 * q4 + j;
              if (m < n_sess) {
                float v = a[j] + bv;
                if (g_act == ACT_GELU_TANH) v = gelu_tanh_fast(v);
                v += rs[4 * q4 + j];
                if (g_bf16)
                  Cb[(size_t)m * g_ldc + nrow] = __float2bfloat16_rn(v);
                else
                  Cf[(size_t)m * g_ldc + nrow] = v;
              }
            }
          }
        }
      }
    }
    if (threadIdx.x == 64) FD_T(6);
    __syncwarp();
    cluster_sync_all();  // peers may still be reading this CTA's partial; also orders TMEM / smem reuse by the next tile
    tc_fence_after();
    if (threadIdx.x == 64) FD_T(7);
  }
#undef FD_T
}

// ---- attention for one (session, head) by ONE WARP: src/model.py:68-98, one query row against [cache ; new row].
// 4 token groups x 8 lanes (HD / 8 dims each); a group takes tokens g, g+4, ...; two tokens' K and V are loaded per
// step before any math (memory-level parallelism); groups are merged with shuffles.  No shared memory, no CTA
// barrier: the 6 warps of a CTA work on 6 different items.
template <int HD>
__device__ __noinline__ void fd_attention_warp(const FusedParams& P, int layer, int b, int h) {
  constexpr int DPL = HD / 8;          // 12
  constexpr int NV = DPL / 4;          // uint2 (4 x bf16) loads per token per lane
  const int lane = threadIdx.x & 31;
  const int g = lane >> 3, sub = lane & 7;
  const int slot = P.slots[b];
  const int C = P.C, page_tokens = P.page_tokens, n_head = P.n_head;
  const long long pool_pages = P.pool_pages;
  bf16* const kv = P.kv;
  bf16* const y = P.y;
  const int T = (int)__ldcg(reinterpret_cast<const unsigned*>(P.st.ctx_len + slot));
  const float scale = rsqrtf((float)HD);
  const float* qrow = P.qkv + (size_t)b * 3 * C + h * HD + sub * DPL;
  float q[DPL], kn[DPL], vn[DPL];
#pragma unroll
  for (int i = 0; i < DPL; i += 4) {
    const float4 a = ldcg4(qrow + i), k4 = ldcg4(qrow + C + i), v4 = ldcg4(qrow + 2 * C + i);
    q[i] = a.x * scale; q[i + 1] = a.y * scale; q[i + 2] = a.z * scale; q[i + 3] = a.w * scale;
    kn[i] = round_to<bf16>(k4.x); kn[i + 1] = round_to<bf16>(k4.y); kn[i + 2] = round_to<bf16>(k4.z); kn[i + 3] = round_to<bf16>(k4.w);
    vn[i] = round_to<bf16>(v4.x); vn[i + 1] = round_to<bf16>(v4.y); vn[i + 2] = round_to<bf16>(v4.z); vn[i + 3] = round_to<bf16>(v4.w);
  }
  const size_t head_stride = (size_t)page_tokens * HD;
  const size_t page_stride = (size_t)n_head * head_stride;
  bf16* kbase = kv + ((size_t)(layer * 2 + 0) * pool_pages) * page_stride + (size_t)h * head_stride + sub * DPL;
  bf16* vbase = kv + ((size_t)(layer * 2 + 1) * pool_pages) * page_stride + (size_t)h * head_stride + sub * DPL;
  const int* pt = P.st.page_table + (size_t)slot * P.st.max_pages;
  // the session's page table goes into registers 64 entries at a time (lane i holds entries i and i + 32 of the
  // window); token -> page is then a shuffle instead of a dependent global load in front of every K/V load
  const int n_pages = (T + page_tokens) / page_tokens;   // pages that hold tokens 0..T
  int win = 0, pt0 = (lane < n_pages) ? pt[lane] : 0, pt1 = (lane + 32 < n_pages) ? pt[lane + 32] : 0;
  auto page_of = [&](int pidx) {   // warp-uniform call sites; pidx inside the current 64-entry window
    const int r = pidx - win;
    const int a = __shfl_sync(0xffffffffu, pt0, r & 31), c = __shfl_sync(0xffffffffu, pt1, r & 31);
    return r < 32 ? a : c;
  };
  if (T / page_tokens >= 64) {   // rare: context beyond the first window; the append below needs its own lookup
    win = (T / page_tokens) & ~63;
    pt0 = (win + lane < n_pages) ? pt[win + lane] : 0;
    pt1 = (win + lane + 32 < n_pages) ? pt[win + lane + 32] : 0;
  }
  const int new_page = page_of(T / page_tokens);
  if (win != 0) {
    win = 0;
    pt0 = (lane < n_pages) ? pt[lane] : 0;
    pt1 = (lane + 32 < n_pages) ? pt[lane + 32] : 0;
  }
  if (g == 0) {   // append the new token (O(1); the reference torch.cat's the whole cache, model.py:74-77)
    const size_t o = (size_t)new_page * page_stride + (size_t)(T % page_tokens) * HD;
#pragma unroll
    for (int i = 0; i < DPL; i += 4) {
      store4(kbase + o + i, make_float4(kn[i], kn[i + 1], kn[i + 2], kn[i + 3]));
      store4(vbase + o + i, make_float4(vn[i], vn[i + 1], vn[i + 2], vn[i + 3]));
    }
  }
  float m = -INFINITY, l = 0.f, acc[DPL];
#pragma unroll
  for (int i = 0; i < DPL; ++i) acc[i] = 0.f;
  // online-softmax update with one token's score d (already reduced over the group's 8 lanes; -inf = no token).
  // No shuffles in here: groups of one warp may take different branches.
  auto absorb = [&](float d, const float* vv) {
    if (d == -INFINITY && m == -INFINITY) return;
    const float mn = fmaxf(m, d);
    const float corr = __expf(m - mn), pr = __expf(d - mn);
    l = l * corr + pr;
#pragma unroll
    for (int i = 0; i < DPL; ++i) acc[i] = acc[i] * corr + pr * vv[i];
    m = mn;
  };
  auto unpack = [](const uint2 (&raw)[NV], float* out) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw[i].x));
      const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw[i].y));
      out[4 * i] = lo.x; out[4 * i + 1] = lo.y; out[4 * i + 2] = hi.x; out[4 * i + 3] = hi.y;
    }
  };
  // every lane runs the same trip count (shuffles are warp-wide); out-of-range tokens contribute d = -inf
  constexpr int UN = 4;   // tokens per group per batch: 4 x 6 loads of 8 bytes in flight per lane
  for (int base = 0; base < T; base += 4 * UN) {
    if (base / page_tokens >= win + 64) {   // next 64-page window (warp-uniform)
      win += 64;
      pt0 = (win + lane < n_pages) ? pt[win + lane] : 0;
      pt1 = (win + lane + 32 < n_pages) ? pt[win + lane + 32] : 0;
    }
    uint2 kr[UN][NV], vr[UN][NV];
    bool ok[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int tok = base + 4 * u + g;
      ok[u] = tok < T;
      const int tk = ok[u] ? tok : 0;
      // shuffles need the whole warp: every lane asks for its own token's page (indices differ per group)
      const int pidx = tk / page_tokens - win;
      const int pa = __shfl_sync(0xffffffffu, pt0, pidx & 31), pc = __shfl_sync(0xffffffffu, pt1, pidx & 31);
      const size_t o = (size_t)(pidx < 32 ? pa : pc) * page_stride + (size_t)(tk % page_tokens) * HD;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        kr[u][i] = __ldcg(reinterpret_cast<const uint2*>(kbase + o) + i);
        vr[u][i] = __ldcg(reinterpret_cast<const uint2*>(vbase + o) + i);
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      float kk[DPL], vv[DPL];
      unpack(kr[u], kk);
      unpack(vr[u], vv);
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < DPL; ++i) d = fmaf(q[i], kk[i], d);
      d += __shfl_xor_sync(0xffffffffu, d, 4);   // executed by every lane of the warp
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      absorb(ok[u] ? d : -INFINITY, vv);
    }
  }
  {   // the new token: taken by group T % 4 (the group that would own index T)
    float d = 0.f;
#pragma unroll
    for (int i = 0; i < DPL; ++i) d = fmaf(q[i], kn[i], d);
    d += __shfl_xor_sync(0xffffffffu, d, 4);
    d += __shfl_xor_sync(0xffffffffu, d, 2);
    d += __shfl_xor_sync(0xffffffffu, d, 1);
    absorb(g == (T & 3) ? d : -INFINITY, vn);
  }
  // merge the 4 groups (lanes with equal `sub` hold the same dims)
#pragma unroll
  for (int o = 8; o <= 16; o <<= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), l2 = __shfl_xor_sync(0xffffffffu, l, o);
    const float M = fmaxf(m, m2);
    const float w1 = (m == -INFINITY) ? 0.f : __expf(m - M), w2 = (m2 == -INFINITY) ? 0.f : __expf(m2 - M);
    l = l * w1 + l2 * w2;
#pragma unroll
    for (int i = 0; i < DPL; ++i) {
      const float a2 = __shfl_xor_sync(0xffffffffu, acc[i], o);
      acc[i] = acc[i] * w1 + a2 * w2;
    }
    m = M;
  }
  if (g == 0) {
    const float inv = 1.0f / l;
    bf16* yo = y + (size_t)b * C + h * HD + sub * DPL;
#pragma unroll
    for (int i = 0; i < DPL; i += 4) store4(yo + i, make_float4(acc[i] * inv, acc[i + 1] * inv, acc[i + 2] * inv, acc[i + 3] * inv));
  }
}

// rows: input assembly (streaming_server.py:313-334, src/model.py:206-212) + LN1 of layer 0; CTA b = session b.
// `code_known`: the previous code was picked by this very CTA a moment ago (no global round trip).  One non-inlined
// copy, so the first iteration of a launch and the later ones run bit-identical arithmetic.
__device__ __noinline__ void fd_assemble_rows(const FusedParams& P, int bid, int nblocks, float* red, bool code_known, int known_code,
                                              int known_t) {
  const int tid = threadIdx.x, C = P.C;
  for (int b = bid; b < P.n; b += nblocks) {
    const int slot = P.slots[b];
    const int t = code_known ? known_t : (int)__ldcg(reinterpret_cast<const unsigned*>(P.st.ctx_len + slot));
    const int c = tid * 4;
    int text_id = P.pad_id;
    if (t < P.st.text_len[slot]) text_id = P.st.text_ids[(size_t)slot * P.st.max_context + t];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < P.text_dim) {
      v = load4(P.text_table + (size_t)text_id * P.text_dim + c);
    } else if (t > 0) {
      const int prev = code_known ? known_code
                                  : (int)__ldcg(reinterpret_cast<const unsigned*>(P.st.codes + (size_t)slot * P.st.max_context + t - 1));
      v = load4(P.codebook + (size_t)prev * P.code_dim + (c - P.text_dim));
    }
    const float ss = block_sum(v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w, red);
    const float denom = fmaxf(sqrtf(ss), 1e-8f);
    const float4 pe = load4(P.wpe + (size_t)t * C + c);
    const float4 xv = make_float4(v.x / denom + pe.x, v.y / denom + pe.y, v.z / denom + pe.z, v.w / denom + pe.w);
    store4(P.x + (size_t)b * C + c, xv);
    fd_ln_row(xv, P.layer[0].ln1_w, P.layer[0].ln1_b, 1e-5f, P.h + (size_t)b * C, red, C);
  }
}

__global__ void __launch_bounds__(TC_THREADS, 1) fused_decode_kernel(const __grid_constant__ FusedParams P) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * TC_MAX_STAGES + 1];
  __shared__ uint32_t tmem_base_sh;
  __shared__ float red[32];
  __shared__ int redi[32];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  const int rank = (int)cluster_ctarank();
  const int cid = (int)blockIdx.x / FD_CLUSTER;
  const int bid = (int)blockIdx.x;
  const unsigned nblocks = gridDim.x;
  const int C = P.C;

  FdPipe pp;
  pp.tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
  pp.red = pp.tiles + (uint32_t)P.stages * (TC_X_BYTES + (uint32_t)P.BN * TC_BK * 2);
  pp.bar0 = smem_u32(bars);
  pp.it_total = 0;
  pp.tile_total = 0;
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < P.stages; ++s) {
      mbar_init(pp.bar0 + 8u * s, 1);
      mbar_init(pp.bar0 + 8u * (TC_MAX_STAGES + s), 1);
    }
    mbar_init(pp.bar0 + 8u * (2 * TC_MAX_STAGES), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc(smem_u32(&tmem_base_sh), (uint32_t)P.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pp.tmem_d = tmem_base_sh;
  cluster_sync_all();   // peers' barriers exist before any DSMEM traffic

  unsigned epoch = 0;
  int tr = 0;
#define FD_STAMP()                                                                         \
  do {                                                                                     \
    if (P.trace && bid == 0 && tid == 0 && iter == P.n_iters - 1 && tr < 250) P.trace[tr++] = clock64(); \
  } while (0)
#define FD_SYNC()                           \
  do {                                      \
    FD_STAMP();                             \
    fd_grid_sync(P.bar, epoch, nblocks);    \
    FD_STAMP();                             \
  } while (0)
  int iter = 0;
  fd_assemble_rows(P, bid, (int)nblocks, red, false, 0, 0);
  FD_SYNC();
  for (iter = 0; iter < P.n_iters; ++iter) {
    FD_STAMP();
    for (int l = 0; l < P.n_layer; ++l) {
      const FusedLayer& L = P.layer[l];
      fd_gemm_phase(P, L.qkv, pp, smem_raw, cid, rank,
                    (P.trace && bid == 0 && iter == P.n_iters - 1 && l == P.n_layer - 1) ? P.trace + 200 : nullptr);
      FD_SYNC();
      for (int item = bid * (TC_THREADS / 32) + warp; item < P.n * P.n_head; item += (int)nblocks * (TC_THREADS / 32))
        fd_attention_warp<96>(P, l, item / P.n_head, item % P.n_head);
      FD_SYNC();
      fd_gemm_phase(P, L.proj, pp, smem_raw, cid, rank);
      FD_SYNC();
      for (int b = bid; b < P.n; b += (int)nblocks)
        fd_ln_row(ldcg4(P.x + (size_t)b * C + tid * 4), L.ln2_w, L.ln2_b, 1e-5f, P.h + (size_t)b * C, red, C);
      FD_SYNC();
      fd_gemm_phase(P, L.fc, pp, smem_raw, cid, rank);
      FD_SYNC();
      fd_gemm_phase(P, L.proj2, pp, smem_raw, cid, rank);
      FD_SYNC();
      const float* nw = (l + 1 < P.n_layer) ? P.layer[l + 1].ln1_w : P.lnf_w;
      const float* nb = (l + 1 < P.n_layer) ? P.layer[l + 1].ln1_b : P.lnf_b;
      for (int b = bid; b < P.n; b += (int)nblocks)
        fd_ln_row(ldcg4(P.x + (size_t)b * C + tid * 4), nw, nb, 1e-5f, P.h + (size_t)b * C, red, C);
      FD_SYNC();
    }
    fd_gemm_phase(P, P.lm_head, pp, smem_raw, cid, rank);
    FD_SYNC();

    // ---------------- rows: greedy pick = argmax, lowest index wins ties (streaming_server.py:342-346), then the next
    // iteration's input assembly by the same CTA (one phase, one barrier)
    for (int b = bid; b < P.n; b += (int)nblocks) {
      const float* lg = P.logits + (size_t)b * P.vocab;
      float best = -INFINITY;
      int bi = 0x7fffffff;
      for (int i = tid * 4; i < P.vocab; i += TC_THREADS * 4) {
        const float4 v = ldcg4(lg + i);
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (vv[j] > best) { best = vv[j]; bi = i + j; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      __syncthreads();
      if (lane == 0) { red[warp] = best; redi[warp] = bi; }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < TC_THREADS / 32; ++w)
          if (red[w] > best || (red[w] == best && redi[w] < bi)) { best = red[w]; bi = redi[w]; }
        const int slot = P.slots[b];
        const int t = (int)__ldcg(reinterpret_cast<const unsigned*>(P.st.ctx_len + slot));
        const int code = (bi == 0x7fffffff) ? 0 : bi;
        P.st.codes[(size_t)slot * P.st.max_context + t] = code;
        P.st.ctx_len[slot] = t + 1;
        redi[30] = code;
        redi[31] = t + 1;
      }
      __syncthreads();
    }
    if (iter + 1 < P.n_iters) {
      const int code = redi[30], t1 = redi[31];   // valid in the CTAs that own a row (the only ones that use them)
      __syncthreads();
      fd_assemble_rows(P, bid, (int)nblocks, red, true, code, t1);
      FD_SYNC();
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(pp.tmem_d, (uint32_t)P.tmem_cols);
  }
}

// host side: launch plan + launch ------------------------------------------------------------------------------------
struct FusedPlan {
  int n_clusters = 0, stages = 0, BN = 0, tmem_cols = 0;
  size_t smem = 0;
  bool cooperative = true;
};

inline int fused_plan(int n, FusedPlan* pl) {
  pl->BN = std::max(16, ceil_div(n, 16) * 16);
  pl->tmem_cols = 32;
  while (pl->tmem_cols < pl->BN) pl->tmem_cols *= 2;
  const int stage_bytes = TC_X_BYTES + pl->BN * TC_BK * 2;
  const size_t red_bytes = (size_t)TC_BM * (pl->BN + 4) * 4;
  int stages = std::max(2, std::min(TC_MAX_STAGES, (int)((196 * 1024 - red_bytes) / stage_bytes)));
  if (const char* env = getenv("LLMVOX_FD_STAGES")) stages = std::max(2, std::min(stages, atoi(env)));
  pl->stages = stages;
  pl->smem = (size_t)stages * stage_bytes + red_bytes + 1024;
  cudaError_t err = cudaFuncSetAttribute(fused_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 16 * 1024);
  if (err != cudaSuccess) {
    set_error(std::string("cudaFuncSetAttribute(fused_decode): ") + cudaGetErrorString(err));
    return LVX_ERR_CUDA;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(FD_CLUSTER * 37);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = pl->smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = FD_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = 0;
  err = cudaOccupancyMaxActiveClusters(&max_clusters, fused_decode_kernel, &cfg);
  if (err != cudaSuccess || max_clusters < 4) {
    set_error(std::string("cudaOccupancyMaxActiveClusters(fused_decode): ") + cudaGetErrorString(err) + " clusters " +
              std::to_string(max_clusters));
    return LVX_ERR_CUDA;
  }
  pl->n_clusters = std::min(max_clusters, 36);
  return LVX_OK;
}

inline int fused_launch(const FusedParams& P, FusedPlan& pl, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(FD_CLUSTER * pl.n_clusters);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = FD_CLUSTER;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pl.cooperative ? 2 : 1;
  cudaError_t err = cudaLaunchKernelEx(&cfg, fused_decode_kernel, P);
  if (err != cudaSuccess && pl.cooperative) {
    // cooperative + cluster launches may be refused by the driver; the grid is sized to be co-resident anyway
    cudaGetLastError();
    pl.cooperative = false;
    cfg.numAttrs = 1;
    err = cudaLaunchKernelEx(&cfg, fused_decode_kernel, P);
  }
  if (err != cudaSuccess) {
    set_error(std::string("fused_decode launch: ") + cudaGetErrorString(err));
    return LVX_ERR_CUDA;
  }
  return LVX_OK;
}

}  // namespace lvx
