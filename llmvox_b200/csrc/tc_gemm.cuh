// tcgen05 / TMEM / TMA GEMM for sm_100a (bf16 operands, fp32 accumulate in tensor memory).
//
//   D[i, j] = sum_k X[i, k] * Y[j, k]        X tile: 128 rows (UMMA M), Y tile: BN rows (UMMA N, 16..256)
//
// Both operands are K-major bf16 matrices read by TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) into a
// multi-stage shared-memory ring; one elected thread issues tcgen05.mma (kind::f16, cta_group::1) into a TMEM
// accumulator; four epilogue warps read it back with tcgen05.ld and apply the GemmParams epilogue.
//
//   normal mode : X = activations (rows m), Y = weight (rows n).  TMEM lane = m, column = n.  Multi-tap
//                 (conv-as-GEMM) shifts the X row coordinate per tap; TMA zero-fills rows outside the tensor.
//   swap mode   : X = weight (rows n), Y = activations (rows m, BN = M rounded up to 16).  TMEM lane = n,
//                 column = m.  For the decode step, where M = sessions in flight is small: the 128-row UMMA M
//                 dimension is filled by weight rows instead of padding, and C stores are coalesced along n.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue
// (TMEM lane quarter = warp_idx % 4).  One output tile per CTA; several CTAs are resident per SM so one CTA's
// epilogue overlaps another's main loop.
#pragma once
#include <cuda.h>

#include <map>
#include <tuple>

#include "common.cuh"
#include "gemm.cuh"

#ifndef LVX_OK
#define LVX_OK 0
#define LVX_ERR_INVALID 1
#define LVX_ERR_CUDA 2
#endif

namespace lvx {

constexpr int TC_BM = 128;  // UMMA M
constexpr int TC_BK = 64;   // bf16 elements per k-block = one 128-byte swizzle row
constexpr int TC_X_BYTES = TC_BM * TC_BK * 2;
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_THREADS = 192;

struct TmaDesc {
  CUtensorMap map;  // box {64, 128}
  bool valid = false;
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct TcWorkspace {
  int num_sms = 0;
  PFN_encodeTiled encode = nullptr;
  std::map<std::tuple<const void*, int, int, int, int>, CUtensorMap> act_maps;
  int* d_err = nullptr;
};

static PFN_encodeTiled g_encode = nullptr;

inline int tc_encode(CUtensorMap* m, const void* ptr, int rows, int cols, int ld, int box_rows) {
  if (!g_encode) {
    set_error("tensor-map encoder not initialised");
    return LVX_ERR_CUDA;
  }
  if ((ld % 8) != 0 || (reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || box_rows < 1 || box_rows > 256) {
    set_error("tensor map: leading dimension must be a multiple of 8 bf16 and the base 16-byte aligned");
    return LVX_ERR_INVALID;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return LVX_ERR_CUDA;
  }
  return LVX_OK;
}

inline int tc_configure();
inline int tc_init(TcWorkspace* ws, int num_sms) {
  ws->num_sms = num_sms;
  {
    int s = tc_configure();
    if (s != LVX_OK) return s;
  }
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (err != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed");
      return LVX_ERR_CUDA;
    }
    g_encode = reinterpret_cast<PFN_encodeTiled>(fn);
  }
  ws->encode = g_encode;
  return LVX_OK;
}

inline int tc_make_desc(TmaDesc* d, const bf16* ptr, int rows, int cols, int ld) {
  int s = tc_encode(&d->map, ptr, rows, cols, ld, TC_BM);
  d->valid = (s == LVX_OK);
  return s;
}

// ------------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must end in a trap (reported as a launch failure), never in a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread t of the warp gets lane (quarter * 32 + t)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// The same load split into issue and wait: the coalesced epilogue starts the TMEM load, then fetches residual / bias /
// scale, and only then waits.  The wait names the registers as in / out operands: nothing that reads them may be
// scheduled above it.
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16], float* v) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                 "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// TMA load multicast to the CTAs of `mask`: the tile lands at the same shared-memory offset in each of them and each CTA's
// mbarrier (same offset) receives the byte count
__device__ __forceinline__ void tma_load_2d_mcast(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// tcgen05.commit that arrives on the mbarrier at this offset in every CTA of `mask`
__device__ __forceinline__ void umma_commit_mcast(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}

// K-major, 128-byte-swizzled operand tile (rows x 64 bf16, 8-row groups 1024 bytes apart)
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);  // start address
  d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D = fp32, A = B = bf16, both K-major, M x N
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

inline bool tc_no_persistent() {
  static const bool off = [] { const char* e = getenv("LLMVOX_B200_NO_PERSISTENT"); return e && e[0] == '1'; }();
  return off;
}

// One problem of a GROUPED launch (normal mode, grid.z = problem): both operands are windows of two big matrices that
// every problem shares -- X rows [x_row0, +M) x K columns [x_k0, +64 num_kb), Y rows [y_row0, +N) x K columns [y_k0, ..)
// -- so ONE pair of tensor maps serves all problems (the ragged pos_net attention of the vocoder: per-chunk Q K^T and P V
// as one launch each instead of one launch per chunk).  Rows past M / N read whatever follows in the shared matrix
// (finite by construction) and are masked at the store.
struct TcGroup {
  int x_row0, y_row0, x_k0, y_k0;
  int M, N, num_kb, ldc;
  long long c_off;   // element offset of the problem's C
};

struct TcParams {
  GemmParams g;
  const TcGroup* groups = nullptr;   // grouped launch: device array, one entry per blockIdx.z
  int BN;          // UMMA N = Y rows per tile
  int stages;
  int num_kb;      // K / 64 (rounded up; TMA zero-fills the tail)
  int tmem_cols;   // power of two >= max(32, BN)
  int splits;      // split-K factor = cluster size along z (1, 2, 4 or 8)
};

// ---- thread-block cluster helpers (split-K partial sums are exchanged through distributed shared memory)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ float4 ld_dsmem_v4(uint32_t local_addr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra) : "memory");
  return v;
}

// exact-erf GELU (torch.nn.GELU(), WavTokenizer/decoder/modules.py:35) with erf by Abramowitz-Stegun 7.1.26
// (|error| <= 1.5e-7, i.e. fp32-exact for this purpose) on MUFU.RCP / MUFU.EX2: ~16 instructions instead of erff's ~40.
// The epilogue of the ConvNeXt pw1 GEMM (K = 768) is instruction-bound, so this is what sets its tensor utilisation.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.7071067811865476f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  const float poly = t * (0.254829592f + t * (-0.284496736f + t * (1.421413741f + t * (-1.453152027f + t * 1.061405429f))));
  const float erfv = copysignf(fmaf(-poly, e, 1.0f), x);
  const float hx = 0.5f * x;
  return fmaf(hx, erfv, hx);
}
// erf-GELU for a bf16 destination: Phi(x) = 0.5 (1 + tanh(a x + b x^3 + c x^5)) with (a, b, c) fitted to the erf form over
// [-8, 8] (max |error| 1.9e-4, below half a bf16 ulp of every |GELU| >= 0.05 and -74 dB against unit-scale activations)
// on MUFU.TANH: 7 instructions and one MUFU instead of 16 and two -- the pw1 epilogue then hides under the tile's MMAs.
// Only the tensor-core path's bf16 outputs use it (LLMVOX_B200_GELU_AS=1 selects the Abramowitz-Stegun form instead).
__device__ __forceinline__ float gelu_erf_bf16(float x) {
  const float x2 = x * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * fmaf(x2, fmaf(x2, -3.68454816e-04f, 3.69380522e-02f), 7.98212028e-01f)));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}
template <int ACT>
__device__ __forceinline__ float tc_act(float v) {
  if (ACT == ACT_GELU_TANH) return gelu_tanh(v);
  if (ACT == ACT_GELU_ERF) return gelu_erf_fast(v);
  if (ACT == ACT_GELU_ERF_BF16) return gelu_erf_bf16(v);
  return v;
}

// Epilogue of `ncols` (multiple of 4, <= 16) accumulator columns j0.. of D row i.  Order as GemmParams states:
// alpha, + bias, activation, * col_scale, + residual.  All loads of a chunk are issued before its stores
// (C may alias residual: every thread reads and writes only its own elements).  ACT is a compile-time switch so the
// per-element path carries no activation branch.
template <bool kSwap, typename TC, int ACT>
__device__ __forceinline__ void tc_epilogue_chunk(const GemmParams& p, int i, int j0, const float* v, int ncols) {
  TC* C = reinterpret_cast<TC*>(p.C);
  if (!kSwap) {
    const int m = i;
    if (m >= p.M || (p.row_chunk && p.row_chunk[m] < 0)) return;
    const bool vec = (p.ldc & (sizeof(TC) == 2 ? 7 : 3)) == 0 && (!p.residual || (p.ldr & 3) == 0) && j0 + ncols <= p.N && (ncols & 7) == 0;
    if (vec) {
      TC* dst = C + (size_t)m * p.ldc + j0;
      const float* res = p.residual ? p.residual + (size_t)m * p.ldr + j0 : nullptr;
#pragma unroll
      for (int j = 0; j < 16; j += 8) {
        if (j < ncols) {
          float o[8];
#pragma unroll
          for (int h = 0; h < 8; h += 4) {
            float4 b = make_float4(0.f, 0.f, 0.f, 0.f), cs = make_float4(1.f, 1.f, 1.f, 1.f), r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) b = load4(p.bias + j0 + j + h);
            if (p.col_scale) cs = load4(p.col_scale + j0 + j + h);
            if (res) r = __ldcg(reinterpret_cast<const float4*>(res + j + h));
            o[h] = tc_act<ACT>(v[j + h] * p.alpha + b.x) * cs.x + r.x;
            o[h + 1] = tc_act<ACT>(v[j + h + 1] * p.alpha + b.y) * cs.y + r.y;
            o[h + 2] = tc_act<ACT>(v[j + h + 2] * p.alpha + b.z) * cs.z + r.z;
            o[h + 3] = tc_act<ACT>(v[j + h + 3] * p.alpha + b.w) * cs.w + r.w;
          }
          if (sizeof(TC) == 2) {   // 8 bf16 = one 16-byte store
            uint4 u;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(o[4], o[5]), h3 = __floats2bfloat162_rn(o[6], o[7]);
            u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
            u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(dst) + j) = u;
          } else {
            store4(dst + j, make_float4(o[0], o[1], o[2], o[3]));
            store4(dst + j + 4, make_float4(o[4], o[5], o[6], o[7]));
          }
        }
      }
    } else {
#pragma unroll 4
      for (int j = 0; j < ncols; ++j) {
        const int nn = j0 + j;
        if (nn < p.N) {
          float xv = v[j] * p.alpha;
          if (p.bias) xv += p.bias[nn];
          xv = tc_act<ACT>(xv);
          if (p.col_scale) xv *= p.col_scale[nn];
          if (p.residual) xv += __ldcg(p.residual + (size_t)m * p.ldr + nn);
          store1(C + (size_t)m * p.ldc + nn, xv);
        }
      }
    }
  } else {
    const int n = i;  // weight row = output column; the 32 lanes of a warp cover 32 consecutive n
    if (n >= p.N) return;
    const float b = p.bias ? p.bias[n] : 0.f;
    const float cs = p.col_scale ? p.col_scale[n] : 1.f;
    if (p.c_transposed) {
      // C[n, m]: the thread's 16 columns are contiguous (the vocoder attention's V^T operand)
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int m = j0 + j;
        if (j < ncols && m < p.M) store1(C + (size_t)n * p.ldc + m, tc_act<ACT>(v[j] * p.alpha + b) * cs);
      }
      return;
    }
#pragma unroll
    for (int j4 = 0; j4 < 16; j4 += 4) {
      if (j4 >= ncols) break;
      float r[4];
      bool ok[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int m = j0 + j4 + t;
        ok[t] = m < p.M && (!p.row_chunk || p.row_chunk[m] >= 0);
        r[t] = (ok[t] && p.residual) ? __ldcg(p.residual + (size_t)m * p.ldr + n) : 0.f;
      }
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (ok[t]) store1(C + (size_t)(j0 + j4 + t) * p.ldc + n, tc_act<ACT>(v[j4 + t] * p.alpha + b) * cs + r[t]);
    }
  }
}

// Normal-mode epilogue of a full 16-column chunk with COALESCED global accesses (persistent kernel).  In TMEM a lane owns a
// D row, so the plain epilogue above makes every warp-wide load / store touch 32 different rows (16 bytes each: 32
// sectors per instruction, every sector fetched twice by successive instructions).  Measured on the ConvNeXt GEMMs at
// 61,440 rows: the residual read alone cost pw2 27 % (697 vs 947 TFLOP/s without it), the row-strided bf16 stores pw1 about
// as much.  Here a warp's 32 x 16 block goes through a per-warp shared-memory scratch (row pitch 80 B: conflict-free
// 16-byte writes per quarter warp) so that global instructions move whole 64-byte row segments, 8 rows per instruction,
// every sector touched once (bf16 destinations: two chunks side by side).
//   row0  = first D row of the warp (lane = row - row0), j0 = first column, v = this lane's 16 accumulator values
//   Requirements (checked by the caller): j0 + 16 <= N, ldc / ldr multiples of 4 elements (8 for bf16).
constexpr int TC_EPI_SCRATCH = 32 * 80;   // bytes per epilogue warp
//   vmask = ballot of the rows that exist and are not padding (computed once per tile by the caller: the row_chunk load
//   cost 17 % of the epilogue's stall samples when it sat in front of every chunk)
__device__ __forceinline__ unsigned tc_row_mask(const GemmParams& p, int m) {
  const bool valid = m < p.M && (!p.row_chunk || p.row_chunk[m] >= 0);
  return __ballot_sync(0xffffffffu, valid);
}
template <typename TC, int ACT>
__device__ __forceinline__ void tc_epilogue_chunk_coalesced(const GemmParams& p, int row0, int lane, int j0, uint32_t taddr, float* scratch,
                                                            int pair, unsigned vmask) {
  if (vmask == 0u) return;   // warp-uniform
  uint32_t acc_raw[16];
  tmem_ld16_issue(taddr, acc_raw);   // on its way while the residual / bias / scale loads below are issued
  float r[16];
  if (p.residual) {
    const int seg = 4 * (lane & 3);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int R = 8 * k + (lane >> 2);
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if ((vmask >> R) & 1u) t = __ldcg(reinterpret_cast<const float4*>(p.residual + (size_t)(row0 + R) * p.ldr + j0 + seg));
      *reinterpret_cast<float4*>(scratch + R * 20 + seg) = t;
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 t = *reinterpret_cast<const float4*>(scratch + lane * 20 + j);
      r[j] = t.x; r[j + 1] = t.y; r[j + 2] = t.z; r[j + 3] = t.w;
    }
    __syncwarp();
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) r[j] = 0.f;
  }
  float4 bb[4], cc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bb[j] = p.bias ? load4(p.bias + j0 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    cc[j] = p.col_scale ? load4(p.col_scale + j0 + 4 * j) : make_float4(1.f, 1.f, 1.f, 1.f);
  }
  float v[16];
  tmem_ld16_wait(acc_raw, v);
  float o[16];
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    const float4 b = bb[j >> 2], cs = cc[j >> 2];
    o[j] = tc_act<ACT>(v[j] * p.alpha + b.x) * cs.x + r[j];
    o[j + 1] = tc_act<ACT>(v[j + 1] * p.alpha + b.y) * cs.y + r[j + 1];
    o[j + 2] = tc_act<ACT>(v[j + 2] * p.alpha + b.z) * cs.z + r[j + 2];
    o[j + 3] = tc_act<ACT>(v[j + 3] * p.alpha + b.w) * cs.w + r[j + 3];
  }
  TC* C = reinterpret_cast<TC*>(p.C);
  if (sizeof(TC) == 2) {
    // bf16: a chunk is only 32 bytes per row, so two consecutive chunks (j0 a multiple of 32, then j0 + 16) are packed side
    // by side into the scratch rows (pitch 80 B) and leave as 64-byte row segments, four lanes per row, after the second
    // one (`pair` = 0: first half, stored later; 1: second half, stores both; 2: a lone chunk, 32-byte segments).
    uint4* sc = reinterpret_cast<uint4*>(scratch);
    const int half = pair == 1 ? 2 : 0;
#pragma unroll
    for (int j = 0; j < 16; j += 8) {
      uint4 u;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(o[j], o[j + 1]), h1 = __floats2bfloat162_rn(o[j + 2], o[j + 3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(o[j + 4], o[j + 5]), h3 = __floats2bfloat162_rn(o[j + 6], o[j + 7]);
      u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
      u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
      sc[lane * 5 + half + (j >> 3)] = u;
    }
    if (pair == 0) return;   // (the next call, for the same rows, completes the pair)
    __syncwarp();
    if (pair == 1) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int R = 8 * k + (lane >> 2), piece = lane & 3;
        if ((vmask >> R) & 1u)
          *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(C) + (size_t)(row0 + R) * p.ldc + (j0 - 16) + 8 * piece) = sc[R * 5 + piece];
      }
    } else {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int R = 16 * k + (lane >> 1), piece = lane & 1;
        if ((vmask >> R) & 1u)
          *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(C) + (size_t)(row0 + R) * p.ldc + j0 + 8 * piece) = sc[R * 5 + piece];
      }
    }
    __syncwarp();
  } else {
#pragma unroll
    for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(scratch + lane * 20 + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
    __syncwarp();
    const int seg = 4 * (lane & 3);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int R = 8 * k + (lane >> 2);
      if ((vmask >> R) & 1u)
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(C) + (size_t)(row0 + R) * p.ldc + j0 + seg) = *reinterpret_cast<const float4*>(scratch + R * 20 + seg);
    }
    __syncwarp();
  }
}
// bf16 destination without a residual (the ConvNeXt pw1 GEMM: K = 768, so the epilogue of a 128 x 256 tile has to keep
// up with 12 k-blocks of MMAs): 32 columns per step -- both TMEM loads in flight together, one wait, one __syncwarp, and
// the 64-byte row segments leave four lanes per row.
template <int ACT>
__device__ __forceinline__ void tc_epilogue_pair_bf16(const GemmParams& p, int row0, int lane, int j0, uint32_t taddr, float* scratch,
                                                      unsigned vmask) {
  if (vmask == 0u) return;   // warp-uniform
  uint32_t ra[16], rb[16];
  tmem_ld16_issue(taddr, ra);
  tmem_ld16_issue(taddr + 16u, rb);
  float4 bb[8], cc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    bb[j] = p.bias ? load4(p.bias + j0 + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    cc[j] = p.col_scale ? load4(p.col_scale + j0 + 4 * j) : make_float4(1.f, 1.f, 1.f, 1.f);
  }
  float v[32];
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(ra[0]), "+r"(ra[1]), "+r"(ra[2]), "+r"(ra[3]), "+r"(ra[4]), "+r"(ra[5]), "+r"(ra[6]), "+r"(ra[7]), "+r"(ra[8]), "+r"(ra[9]),
                 "+r"(ra[10]), "+r"(ra[11]), "+r"(ra[12]), "+r"(ra[13]), "+r"(ra[14]), "+r"(ra[15]), "+r"(rb[0]), "+r"(rb[1]), "+r"(rb[2]),
                 "+r"(rb[3]), "+r"(rb[4]), "+r"(rb[5]), "+r"(rb[6]), "+r"(rb[7]), "+r"(rb[8]), "+r"(rb[9]), "+r"(rb[10]), "+r"(rb[11]),
                 "+r"(rb[12]), "+r"(rb[13]), "+r"(rb[14]), "+r"(rb[15])
               :
               : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    v[i] = __uint_as_float(ra[i]);
    v[16 + i] = __uint_as_float(rb[i]);
  }
  uint4* sc = reinterpret_cast<uint4*>(scratch);
#pragma unroll
  for (int j = 0; j < 32; j += 8) {
    const float4 b0 = bb[j >> 2], b1 = bb[(j >> 2) + 1], c0 = cc[j >> 2], c1 = cc[(j >> 2) + 1];
    const float o0 = tc_act<ACT>(v[j] * p.alpha + b0.x) * c0.x, o1 = tc_act<ACT>(v[j + 1] * p.alpha + b0.y) * c0.y;
    const float o2 = tc_act<ACT>(v[j + 2] * p.alpha + b0.z) * c0.z, o3 = tc_act<ACT>(v[j + 3] * p.alpha + b0.w) * c0.w;
    const float o4 = tc_act<ACT>(v[j + 4] * p.alpha + b1.x) * c1.x, o5 = tc_act<ACT>(v[j + 5] * p.alpha + b1.y) * c1.y;
    const float o6 = tc_act<ACT>(v[j + 6] * p.alpha + b1.z) * c1.z, o7 = tc_act<ACT>(v[j + 7] * p.alpha + b1.w) * c1.w;
    uint4 u;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(o0, o1), h1 = __floats2bfloat162_rn(o2, o3);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(o4, o5), h3 = __floats2bfloat162_rn(o6, o7);
    u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
    sc[lane * 5 + (j >> 3)] = u;
  }
  __syncwarp();
  bf16* C = reinterpret_cast<bf16*>(p.C);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int R = 8 * k + (lane >> 2), piece = lane & 3;
    if ((vmask >> R) & 1u) *reinterpret_cast<uint4*>(C + (size_t)(row0 + R) * p.ldc + j0 + 8 * piece) = sc[R * 5 + piece];
  }
  __syncwarp();
}
__device__ __forceinline__ void tc_epilogue_pair_bf16_dispatch(const GemmParams& p, int row0, int lane, int j0, uint32_t taddr, float* scratch,
                                                               unsigned vmask) {
  if (p.act == ACT_GELU_ERF_BF16)
    tc_epilogue_pair_bf16<ACT_GELU_ERF_BF16>(p, row0, lane, j0, taddr, scratch, vmask);
  else if (p.act == ACT_GELU_ERF)
    tc_epilogue_pair_bf16<ACT_GELU_ERF>(p, row0, lane, j0, taddr, scratch, vmask);
  else if (p.act == ACT_GELU_TANH)
    tc_epilogue_pair_bf16<ACT_GELU_TANH>(p, row0, lane, j0, taddr, scratch, vmask);
  else
    tc_epilogue_pair_bf16<ACT_NONE>(p, row0, lane, j0, taddr, scratch, vmask);
}

template <typename TC>
__device__ __forceinline__ void tc_epilogue_dispatch_coalesced(const GemmParams& p, int row0, int lane, int j0, uint32_t taddr, float* scratch,
                                                               int pair, unsigned vmask) {
  if (p.act == ACT_GELU_ERF_BF16)
    tc_epilogue_chunk_coalesced<TC, ACT_GELU_ERF_BF16>(p, row0, lane, j0, taddr, scratch, pair, vmask);
  else if (p.act == ACT_GELU_ERF)
    tc_epilogue_chunk_coalesced<TC, ACT_GELU_ERF>(p, row0, lane, j0, taddr, scratch, pair, vmask);
  else if (p.act == ACT_GELU_TANH)
    tc_epilogue_chunk_coalesced<TC, ACT_GELU_TANH>(p, row0, lane, j0, taddr, scratch, pair, vmask);
  else
    tc_epilogue_chunk_coalesced<TC, ACT_NONE>(p, row0, lane, j0, taddr, scratch, pair, vmask);
}

template <bool kSwap, typename TC>
__device__ __forceinline__ void tc_epilogue_dispatch(const GemmParams& p, int i, int j0, const float* v, int ncols) {
  if (p.act == ACT_GELU_ERF_BF16)
    tc_epilogue_chunk<kSwap, TC, ACT_GELU_ERF_BF16>(p, i, j0, v, ncols);
  else if (p.act == ACT_GELU_ERF)
    tc_epilogue_chunk<kSwap, TC, ACT_GELU_ERF>(p, i, j0, v, ncols);
  else if (p.act == ACT_GELU_TANH)
    tc_epilogue_chunk<kSwap, TC, ACT_GELU_TANH>(p, i, j0, v, ncols);
  else
    tc_epilogue_chunk<kSwap, TC, ACT_NONE>(p, i, j0, v, ncols);
}

// warps: 0 = TMA producer, 1 = MMA issuer, 2.. = epilogue.  Swap mode (small N tile) keeps 4 epilogue warps; normal
// mode uses 8 (two per TMEM lane quarter, each taking half of the tile's columns): its epilogues (GELU, gamma,
// residual) are instruction-bound at K = 768.
template <bool kSwap>
struct TcShape {
  static constexpr int kEpiWarps = kSwap ? 4 : 8;
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
};

template <bool kSwap, typename TC, bool kGrouped = false>
__global__ void __launch_bounds__(TcShape<kSwap>::kThreads, kSwap ? 1 : 2) tc_gemm_kernel(const __grid_constant__ CUtensorMap mapX,
                                                                          const __grid_constant__ CUtensorMap mapY,
                                                                          const TcParams tp) {
  static_assert(!(kGrouped && kSwap), "grouped launches are normal-mode only");
  constexpr int kHalves = TcShape<kSwap>::kEpiWarps / 4;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * TC_MAX_STAGES + 1];
  __shared__ uint32_t tmem_base_sh;

  // grouped launch: this CTA's problem; tiles outside the problem's extent (the grid is sized for the largest) leave
  GemmParams pg;
  int g_x0 = 0, g_y0 = 0, g_xk = 0, g_yk = 0, g_nkb = 0;
  if (kGrouped) {
    const TcGroup gr = tp.groups[blockIdx.z];
    if ((int)blockIdx.y * TC_BM >= gr.M || (int)blockIdx.x * tp.BN >= gr.N) return;
    pg = tp.g;
    pg.M = gr.M; pg.N = gr.N; pg.ldc = gr.ldc;
    pg.C = reinterpret_cast<TC*>(pg.C) + gr.c_off;
    g_x0 = gr.x_row0; g_y0 = gr.y_row0; g_xk = gr.x_k0; g_yk = gr.y_k0; g_nkb = gr.num_kb;
  }
  pdl_launch_dependents();   // the successor may be scheduled as soon as every CTA of this grid is running
  const GemmParams& p = kGrouped ? pg : tp.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = tp.BN, stages = tp.stages, S = kGrouped ? 1 : tp.splits;
  const uint32_t y_bytes = (uint32_t)BN * TC_BK * 2;
  const uint32_t stage_bytes = TC_X_BYTES + y_bytes;
  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (TC_MAX_STAGES + s); };
  const uint32_t tmem_full_bar = bar0 + 8u * (2 * TC_MAX_STAGES);

  // tile origin: x0 = first X row, y0 = first Y row; this CTA's share of the K loop
  const int x0 = (kSwap ? blockIdx.x : blockIdx.y) * TC_BM;
  const int y0 = (kSwap ? blockIdx.y : blockIdx.x) * BN;
  const int rank = S > 1 ? (int)cluster_ctarank() : 0;
  const int num_kb = kGrouped ? g_nkb : tp.num_kb;
  const int kb0 = (int)((long long)num_kb * rank / S), kb1 = (int)((long long)num_kb * (rank + 1) / S);
  const int nk = kb1 - kb0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapY);
    for (int s = 0; s < stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc(smem_u32(&tmem_base_sh), (uint32_t)tp.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_sh;

  // split-K partial tile in shared memory (reuses the operand ring once the MMAs are done): [128][BN + 4] fp32
  const int RS = BN + 4;
  const int q = warp & 3;              // TMEM lane quarter an epilogue warp may read
  const int half = (warp - 2) >> 2;    // which share of the columns this epilogue warp takes (0 .. kHalves-1)
  const int drow = q * 32 + lane;      // D row (TMEM lane) of an epilogue thread
  const uint32_t trow = tmem_d + ((uint32_t)(q * 32) << 16);

  if (warp == 0) {
    if (lane == 0) {
      // swap mode: X tiles are WEIGHTS, which do not depend on the predecessor kernel -- their loads for the first ring
      // pass are issued before waiting for it (PDL); the activation (Y) tiles follow after the wait
      const int pre = kSwap ? min(nk, stages) : 0;
      for (int it = 0; it < pre; ++it) {
        mbar_expect_tx(full_bar(it), stage_bytes);
        tma_load_2d(&mapX, full_bar(it), tiles + (uint32_t)it * stage_bytes, (kb0 + it) * TC_BK, x0);
      }
      pdl_wait();
      for (int it = 0; it < pre; ++it)
        tma_load_2d(&mapY, full_bar(it), tiles + (uint32_t)it * stage_bytes + TC_X_BYTES, (kb0 + it) * TC_BK, y0);
      for (int it = pre; it < nk; ++it) {
        const int s = it % stages;
        const uint32_t ph = (uint32_t)(it / stages) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        mbar_expect_tx(full_bar(s), stage_bytes);
        const int k0 = (kb0 + it) * TC_BK;
        int xc = k0, xr = x0;
        if (!kSwap && p.taps > 1) {
          const int tap = k0 / p.tap_K;
          xc = k0 - tap * p.tap_K;
          xr = x0 + tap - p.tap_pad;
        }
        const uint32_t dst = tiles + (uint32_t)s * stage_bytes;
        tma_load_2d(&mapX, full_bar(s), dst, xc + g_xk, xr + g_x0);
        tma_load_2d(&mapY, full_bar(s), dst + TC_X_BYTES, k0 + g_yk, y0 + g_y0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(TC_BM, BN);
      for (int it = 0; it < nk; ++it) {
        const int s = it % stages;
        const uint32_t ph = (uint32_t)(it / stages) & 1u;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t xs = tiles + (uint32_t)s * stage_bytes;
        const uint64_t adesc = umma_smem_desc(xs), bdesc = umma_smem_desc(xs + TC_X_BYTES);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          // +32 bytes (16 bf16) along K inside the swizzle atom = +2 in the encoded start address
          umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(empty_bar(s));  // frees the smem stage once these MMAs have read it
      }
      umma_commit(tmem_full_bar);   // accumulator complete
    }
  } else {
    // ---------------- epilogue warps: TMEM -> registers -> (global | partial tile in smem)
    mbar_wait(tmem_full_bar, 0);
    __syncwarp();   // the polling loop may release lanes at different iterations; tcgen05.ld is .sync.aligned
    tc_fence_after();
    pdl_wait();   // residual reads and output writes must follow the predecessor's completion
    const int c_lo = half * (BN / kHalves), c_hi = (half + 1) * (BN / kHalves);   // multiples of 8 (BN % 16 == 0)
    if (S == 1) {
      // normal mode: the coalesced epilogue of the persistent kernel, with the (now idle) operand ring as its scratch
      float* const scratch = reinterpret_cast<float*>(smem_raw + (tiles - smem_u32(smem_raw)) + (warp - 2) * TC_EPI_SCRATCH);
      const bool coalesced = !kSwap && (size_t)stages * stage_bytes >= (size_t)TcShape<kSwap>::kEpiWarps * TC_EPI_SCRATCH &&
                             (p.ldc & (sizeof(TC) == 2 ? 7 : 3)) == 0 && (!p.residual || (p.ldr & 3) == 0) &&
                             (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0;
      const unsigned vmask = coalesced ? tc_row_mask(p, x0 + drow) : 0u;
      for (int c = c_lo; c < c_hi; c += 16) {
        if (!kSwap && coalesced && c + 16 <= c_hi && y0 + c + 16 <= p.N) {   // warp-uniform
          if (sizeof(TC) == 2 && !p.residual && c + 32 <= c_hi && y0 + c + 32 <= p.N) {
            tc_epilogue_pair_bf16_dispatch(p, x0 + q * 32, lane, y0 + c, trow + (uint32_t)c, scratch, vmask);
            c += 16;
          } else {
            tc_epilogue_dispatch_coalesced<TC>(p, x0 + q * 32, lane, y0 + c, trow + (uint32_t)c, scratch, 2, vmask);
          }
          continue;
        }
        float v[16];
        tmem_ld16(trow + (uint32_t)c, v);
        tc_epilogue_dispatch<kSwap, TC>(p, x0 + drow, y0 + c, v, min(16, c_hi - c));
      }
    } else {
      float* red = reinterpret_cast<float*>(smem_raw + (tiles - smem_u32(smem_raw)));
      for (int c = c_lo; c < c_hi; c += 16) {
        float v[16];
        tmem_ld16(trow + (uint32_t)c, v);
        float* dst = red + (size_t)drow * RS + c;
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          if (c + j < c_hi) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    }
  }
  if (S > 1) {
    // every CTA of the cluster now holds its K-slice partial; rank r reduces and finishes columns
    // [r * BN / S, (r + 1) * BN / S) in fixed rank order (deterministic), reading peers through DSMEM
    tc_fence_before();
    __syncwarp();
    cluster_sync_all();
    if (warp >= 2) {
      const int cw = BN / S;                       // multiple of 4 (launch plan)
      const int cwh = (cw / kHalves + 3) & ~3;     // this warp's share of the rank's columns
      const int c_lo = rank * cw + half * cwh, c_hi = min((rank + 1) * cw, c_lo + cwh);
      const uint32_t red0 = tiles;
      for (int c = c_lo; c < c_hi; c += 16) {
        const int ncols = min(16, c_hi - c);
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j < ncols) {
            const uint32_t off = red0 + (uint32_t)(((size_t)drow * RS + c + j) * 4);
            for (int r = 0; r < S; ++r) {
              const float4 t = ld_dsmem_v4(off, (uint32_t)r);
              a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
            }
          }
          v[j] = a.x; v[j + 1] = a.y; v[j + 2] = a.z; v[j + 3] = a.w;
        }
        tc_epilogue_dispatch<kSwap, TC>(p, x0 + drow, y0 + c, v, ncols);
      }
    }
    __syncwarp();
    cluster_sync_all();   // peers may still be reading this CTA's partial
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_d, (uint32_t)tp.tmem_cols);
  }
}

inline int tc_act_map(TcWorkspace* ws, const void* ptr, int rows, int cols, int ld, int box_rows, CUtensorMap* out) {
  auto key = std::make_tuple(ptr, rows, cols, ld, box_rows);
  auto it = ws->act_maps.find(key);
  if (it == ws->act_maps.end()) {
    CUtensorMap m;
    int s = tc_encode(&m, ptr, rows, cols, ld, box_rows);
    if (s != LVX_OK) return s;
    if (ws->act_maps.size() > 4096) ws->act_maps.clear();
    it = ws->act_maps.emplace(key, m).first;
  }
  *out = it->second;
  return LVX_OK;
}

template <typename TC>
__global__ void tc_gemm_persistent_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapY,
                                          const TcParams tp);

// opt every instantiation into the large dynamic shared-memory carve-out (done once per device at engine creation,
// never inside a stream capture)
inline int tc_configure() {
  cudaError_t err = cudaFuncSetAttribute(tc_gemm_kernel<true, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
  if (err == cudaSuccess)
    err = cudaFuncSetAttribute(tc_gemm_kernel<true, bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
  if (err == cudaSuccess)
    err = cudaFuncSetAttribute(tc_gemm_kernel<false, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
  if (err == cudaSuccess)
    err = cudaFuncSetAttribute(tc_gemm_kernel<false, bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
  if (err == cudaSuccess)
    err = cudaFuncSetAttribute(tc_gemm_kernel<false, float, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
  if (err == cudaSuccess)
    err = cudaFuncSetAttribute(tc_gemm_kernel<false, bf16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
  if (err == cudaSuccess)
    err = cudaFuncSetAttribute(tc_gemm_persistent_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
  if (err == cudaSuccess)
    err = cudaFuncSetAttribute(tc_gemm_persistent_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 2048);
  if (err != cudaSuccess) {
    set_error(std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(err));
    return LVX_ERR_CUDA;
  }
  return LVX_OK;
}

template <bool kSwap, typename TC>
inline int tc_launch(const CUtensorMap& mx, const CUtensorMap& my, const TcParams& tp, dim3 grid, size_t smem, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(TcShape<kSwap>::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = (unsigned)tp.splits;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tp.g.pdl ? 2 : 1;
  cudaError_t err = cudaLaunchKernelEx(&cfg, tc_gemm_kernel<kSwap, TC>, mx, my, tp);
  if (err != cudaSuccess) {
    set_error(std::string("tc_gemm launch: ") + cudaGetErrorString(err));
    return LVX_ERR_CUDA;
  }
  return LVX_OK;
}

// Grouped launch (see TcGroup): n_groups problems of at most max_M x max_N, 128 x 128 tiles, no split-K.  mx / my: tensor
// maps (box 128 rows x 64 columns) of the two shared matrices; p carries alpha and the output base pointer.
inline int tc_gemm_grouped(const CUtensorMap& mx, const CUtensorMap& my, const GemmParams& p, const TcGroup* d_groups, int n_groups, int max_M,
                           int max_N, int max_kb, bool c_bf16, cudaStream_t st) {
  if (n_groups <= 0) return LVX_OK;
  TcParams tp;
  tp.g = p;
  tp.groups = d_groups;
  tp.num_kb = max_kb;
  tp.BN = 128;
  tp.tmem_cols = 128;
  tp.splits = 1;
  const int stage_bytes = TC_X_BYTES + tp.BN * TC_BK * 2;
  tp.stages = std::max(2, std::min(std::min(TC_MAX_STAGES, max_kb), (110 * 1024 - 1024) / stage_bytes));
  const size_t smem = (size_t)tp.stages * stage_bytes + 1024;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ceil_div(max_N, tp.BN), ceil_div(max_M, TC_BM), n_groups);
  cfg.blockDim = dim3(TcShape<false>::kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaError_t err = c_bf16 ? cudaLaunchKernelEx(&cfg, tc_gemm_kernel<false, bf16, true>, mx, my, tp)
                           : cudaLaunchKernelEx(&cfg, tc_gemm_kernel<false, float, true>, mx, my, tp);
  if (err != cudaSuccess) {
    set_error(std::string("tc_gemm grouped launch: ") + cudaGetErrorString(err));
    return LVX_ERR_CUDA;
  }
  return LVX_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent normal-mode GEMM for large M (the vocoder's bulk shapes): 128 x 256 tiles, one CTA per SM looping over
// tiles, TWO TMEM accumulators (2 x 256 columns = all 512) so that the epilogue of tile i overlaps the MMAs of tile
// i + 1 inside the same CTA, 4-stage operand ring (48 KB per stage).  CTAs work in CLUSTERS OF TWO on vertically adjacent
// tiles (same weight columns): each CTA loads its own activation tile and HALF of the weight tile, the latter with TMA
// multicast into both CTAs, so the weight operand crosses L2 -> SM once per pair (operand bytes per flop drop by a third;
// these GEMMs run against the L2 bandwidth ceiling).  A ring stage is reused only after BOTH CTAs' MMAs have read it: the
// tcgen05.commit that frees a stage arrives on the empty barrier of both CTAs.
// Warps: 0 TMA producer, 1 MMA issuer, 2..9 epilogue (two per TMEM lane quarter, 128 columns each).
constexpr int TCP_BN = 256;
constexpr int TCP_STAGES = 4;
constexpr int TCP_THREADS = 320;
constexpr int TCP_STAGE_BYTES = TC_X_BYTES + TCP_BN * TC_BK * 2;   // 48 KB

template <typename TC>
__global__ void __launch_bounds__(TCP_THREADS, 1) tc_gemm_persistent_kernel(const __grid_constant__ CUtensorMap mapX,
                                                                            const __grid_constant__ CUtensorMap mapY,
                                                                            const TcParams tp) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bars[2 * TCP_STAGES + 4];
  __shared__ uint32_t tmem_base_sh;

  const GemmParams& p = tp.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (TCP_STAGES + s); };
  auto tfull_bar = [&](int a) { return bar0 + 8u * (2 * TCP_STAGES + a); };
  auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * TCP_STAGES + 2 + a); };
  const int num_kb = tp.num_kb;
  const int n_tiles_n = ceil_div(p.N, TCP_BN), n_pairs_m = ceil_div(ceil_div(p.M, TC_BM), 2);
  const int n_items = n_tiles_n * n_pairs_m;           // one item = two vertically adjacent tiles
  const int rank = (int)cluster_ctarank();             // 0 / 1 inside the pair
  const int cluster_id = (int)blockIdx.x >> 1, n_clusters = (int)gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapX);
    tma_prefetch_desc(&mapY);
    for (int s = 0; s < TCP_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 2);    // both CTAs of the pair have consumed the stage
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8);   // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    __syncwarp();
    tmem_alloc(smem_u32(&tmem_base_sh), 512u);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_sh;
  cluster_sync_all();   // the peer's barriers exist before any multicast reaches them

  if (warp == 0) {
    if (lane == 0) {
      int gi = 0;
      for (int w = cluster_id; w < n_items; w += n_clusters) {
        // n fastest: neighbouring clusters share the activation tiles in L2
        const int x0 = ((w / n_tiles_n) * 2 + rank) * TC_BM, y0 = (w % n_tiles_n) * TCP_BN;
        for (int kb = 0; kb < num_kb; ++kb, ++gi) {
          const int s = gi % TCP_STAGES;
          const uint32_t ph = (uint32_t)(gi / TCP_STAGES) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          mbar_expect_tx(full_bar(s), TCP_STAGE_BYTES);
          const int k0 = kb * TC_BK;
          int xc = k0, xr = x0;
          if (p.taps > 1) {
            const int tap = k0 / p.tap_K;
            xc = k0 - tap * p.tap_K;
            xr = x0 + tap - p.tap_pad;
          }
          const uint32_t dst = tiles + (uint32_t)s * TCP_STAGE_BYTES;
          tma_load_2d(&mapX, full_bar(s), dst, xc, xr);
          // this CTA's half of the weight tile (rows y0 + 128 * rank ..), delivered to both CTAs
          tma_load_2d_mcast(&mapY, full_bar(s), dst + TC_X_BYTES + (uint32_t)rank * TC_X_BYTES, k0, y0 + rank * TC_BM, (uint16_t)3);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(TC_BM, TCP_BN);
      int gi = 0, i = 0;
      for (int w = cluster_id; w < n_items; w += n_clusters, ++i) {
        const int acc = i & 1;
        mbar_wait(tempty_bar(acc), ((uint32_t)(i >> 1) & 1u) ^ 1u);   // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t dcol = tmem_d + (uint32_t)(acc * TCP_BN);
        for (int kb = 0; kb < num_kb; ++kb, ++gi) {
          const int s = gi % TCP_STAGES;
          const uint32_t ph = (uint32_t)(gi / TCP_STAGES) & 1u;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t xs = tiles + (uint32_t)s * TCP_STAGE_BYTES;
          const uint64_t adesc = umma_smem_desc(xs), bdesc = umma_smem_desc(xs + TC_X_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)
            umma_bf16(dcol, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_mcast(empty_bar(s), (uint16_t)3);   // frees the stage in BOTH CTAs once these MMAs have read it
        }
        umma_commit(tfull_bar(acc));
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int drow = q * 32 + lane;
    // per-warp scratch of the coalesced epilogue, behind the operand ring
    float* const scratch = reinterpret_cast<float*>(smem_raw + (tiles - smem_u32(smem_raw)) + TCP_STAGES * TCP_STAGE_BYTES + (warp - 2) * TC_EPI_SCRATCH);
    const bool coalesced = (p.ldc & (sizeof(TC) == 2 ? 7 : 3)) == 0 && (!p.residual || (p.ldr & 3) == 0) &&
                           (reinterpret_cast<uintptr_t>(p.C) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.residual) & 15) == 0;
    // validity of this lane's row, loaded one tile ahead (behind the wait its global load cost 3.7 % of the kernel's stall samples)
    auto row_valid = [&](int w) -> bool {
      const int m = ((w / n_tiles_n) * 2 + rank) * TC_BM + drow;
      return w < n_items && m < p.M && (!p.row_chunk || p.row_chunk[m] >= 0);
    };
    bool rv_next = row_valid(cluster_id);
    int i = 0;
    for (int w = cluster_id; w < n_items; w += n_clusters, ++i) {
      const int acc = i & 1;
      const int x0 = ((w / n_tiles_n) * 2 + rank) * TC_BM, y0 = (w % n_tiles_n) * TCP_BN;
      const bool rv = rv_next;
      rv_next = row_valid(w + n_clusters);
      const unsigned vmask = __ballot_sync(0xffffffffu, rv);
      mbar_wait(tfull_bar(acc), (uint32_t)(i >> 1) & 1u);
      __syncwarp();   // reconverge before the .sync.aligned tcgen05.ld
      tc_fence_after();
      const uint32_t trow = tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TCP_BN);
      const int c_lo = half * (TCP_BN / 2), c_hi = c_lo + TCP_BN / 2;
      for (int c = c_lo; c < c_hi; c += 16) {
        if (y0 + c >= p.N) break;   // warp-uniform: columns past N (last N tile)
        // (issuing the NEXT chunk's TMEM load before processing this one: 160 registers, 8 % slower -- measured;
        //  12 epilogue warps, three per lane quarter with 96 / 96 / 64 columns: pw1 2.88 ms against 2.85 with 8 -- no gain)
        if (sizeof(TC) == 2 && coalesced && !p.residual && y0 + c + 32 <= p.N) {   // warp-uniform: 32 columns per step
          tc_epilogue_pair_bf16_dispatch(p, x0 + q * 32, lane, y0 + c, trow + (uint32_t)c, scratch, vmask);
          c += 16;
          continue;
        }
        if (coalesced && y0 + c + 16 <= p.N) {   // warp-uniform
          // bf16 destination: chunks leave in pairs (c is a multiple of 16; the pair starts at a multiple of 32)
          // (not with a residual: its staging uses the same scratch rows between the two halves)
          const int pair = (sizeof(TC) == 2 && !p.residual) ? ((c & 16) ? 1 : (y0 + c + 32 <= p.N ? 0 : 2)) : 2;
          tc_epilogue_dispatch_coalesced<TC>(p, x0 + q * 32, lane, y0 + c, trow + (uint32_t)c, scratch, pair, vmask);
        } else {
          float v[16];
          tmem_ld16(trow + (uint32_t)c, v);
          tc_epilogue_dispatch<false, TC>(p, x0 + drow, y0 + c, v, 16);   // rows >= M (odd tile count: the pair's second tile) store nothing
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty_bar(acc)) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer may still multicast into this CTA's smem / arrive on its barriers
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_d, 512u);
  }
}

template <typename TC>
inline int tc_launch_persistent(const CUtensorMap& mx, const CUtensorMap& my, const TcParams& tp, int grid, cudaStream_t st) {
  const size_t smem = (size_t)TCP_STAGES * TCP_STAGE_BYTES + 8 * TC_EPI_SCRATCH + 1024;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(grid & ~1));
  cfg.blockDim = dim3(TCP_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t err = cudaLaunchKernelEx(&cfg, tc_gemm_persistent_kernel<TC>, mx, my, tp);
  if (err != cudaSuccess) {
    set_error(std::string("tc_gemm_persistent launch: ") + cudaGetErrorString(err));
    return LVX_ERR_CUDA;
  }
  return LVX_OK;
}

// Launch plan: tile shape, split-K factor and pipeline depth for one problem.
struct TcPlan {
  bool swap;
  int BN, splits, stages, tmem_cols;
  dim3 grid;
  size_t smem;
};

inline TcPlan tc_plan(const TcWorkspace* ws, const GemmParams& p) {
  TcPlan pl;
  const int num_kb = ceil_div(p.K, TC_BK);
  pl.swap = p.c_transposed || (p.taps == 1 && p.M <= 256);
  pl.BN = pl.swap ? std::min(256, std::max(16, ceil_div(p.M, 16) * 16)) : 128;
  const int gx = pl.swap ? ceil_div(p.N, TC_BM) : ceil_div(p.N, pl.BN);
  const int gy = pl.swap ? ceil_div(p.M, pl.BN) : ceil_div(p.M, TC_BM);
  const int tiles = gx * gy;
  const int all_sms = ws->num_sms > 0 ? ws->num_sms : 148;
  const int sms = p.cta_budget > 0 ? std::min(p.cta_budget, all_sms) : all_sms;
  // few tiles: split K across a cluster so that about 1.5 CTAs per SM pull operands concurrently.  Swap mode only (M <= 256,
  // where the weights are the whole traffic): in normal mode the DSMEM reduction and the cluster barriers cost more than
  // the extra CTAs bring -- measured on the kernel-per-op decode chain at T = 300: 288 sessions 675 -> 500 us per
  // iteration, 512: 820 -> 680, 2048: 2011 -> 1926 (proj 31 -> 17 us, proj2 38 -> 30 us per launch); vocoder 64 x 10 frames 1.18 -> 1.07 ms
  int S = 1;
  while (pl.swap && S < 8 && tiles * (S * 2) <= sms + sms / 2 && num_kb >= S * 2 && (pl.BN / (S * 2)) % 4 == 0 && pl.BN / (S * 2) >= 4) S *= 2;
  pl.splits = S;
  pl.tmem_cols = 32;
  while (pl.tmem_cols < pl.BN) pl.tmem_cols *= 2;
  const int stage_bytes = TC_X_BYTES + pl.BN * TC_BK * 2;
  const int nk = ceil_div(num_kb, S);
  const bool one_per_sm = tiles * S <= all_sms || pl.BN > 128;
  const int budget = one_per_sm ? 200 * 1024 : 110 * 1024;
  int stages = std::max(2, std::min(std::min(TC_MAX_STAGES, nk), (budget - 1024) / stage_bytes));
  const size_t red_bytes = S > 1 ? (size_t)TC_BM * (pl.BN + 4) * 4 : 0;
  while ((size_t)stages * stage_bytes < red_bytes) ++stages;
  pl.stages = stages;
  pl.smem = (size_t)stages * stage_bytes + 1024;
  pl.grid = dim3(gx, gy, S);
  return pl;
}

// p.A: bf16 activations (M x K view, lda), p.a_cap rows addressable.  w: tensor map of the bf16 (N, K) weight.
inline int tc_gemm(TcWorkspace* ws, const GemmParams& p, const TmaDesc& w, bool c_bf16, cudaStream_t st,
                   const CUtensorMap* a_map = nullptr) {
  if (!w.valid) {
    set_error("tc_gemm: weight has no tensor map");
    return LVX_ERR_INVALID;
  }
  if (p.M <= 0) return LVX_OK;
  if (p.batch || p.w_kn || (p.taps > 1 && (p.tap_K % TC_BK) != 0)) {
    set_error("tc_gemm: unsupported problem shape");
    return LVX_ERR_INVALID;
  }
  const int a_cols = p.taps > 1 ? p.tap_K : p.K;
  const int a_cap = p.a_cap ? p.a_cap : (p.a_rows ? p.a_rows : p.M);
  const TcPlan pl = tc_plan(ws, p);
  TcParams tp;
  tp.g = p;
  tp.num_kb = ceil_div(p.K, TC_BK);
  tp.BN = pl.BN;
  tp.tmem_cols = pl.tmem_cols;
  tp.stages = pl.stages;
  tp.splits = pl.splits;
  // large-M normal-mode problems: persistent 128 x 256 tiles (at least ~1.3 tiles per SM, N wide enough to fill them)
  if (!pl.swap && !tc_no_persistent() && p.N >= 192) {
    const int n_tiles = ceil_div(p.M, TC_BM) * ceil_div(p.N, TCP_BN);
    const int sms = ws->num_sms > 0 ? ws->num_sms : 148;
    if (n_tiles >= sms + sms / 3) {
      CUtensorMap pam;
      int ps = LVX_OK;
      if (a_map)
        pam = *a_map;
      else
        ps = tc_act_map(ws, p.A, a_cap, a_cols, p.lda, TC_BM, &pam);
      if (ps != LVX_OK) return ps;
      TcParams pp = tp;
      pp.BN = TCP_BN;
      pp.splits = 1;
      const int grid = std::max(2, std::min(2 * ceil_div(ceil_div(p.M, TC_BM), 2) * ceil_div(p.N, TCP_BN), sms));
      return c_bf16 ? tc_launch_persistent<bf16>(pam, w.map, pp, grid, st) : tc_launch_persistent<float>(pam, w.map, pp, grid, st);
    }
  }
  CUtensorMap am;
  int s = LVX_OK;
  if (a_map)
    am = *a_map;   // caller-built (per-chunk attention operands): box rows must be 128
  else
    s = tc_act_map(ws, p.A, a_cap, a_cols, p.lda, pl.swap ? pl.BN : TC_BM, &am);
  if (s != LVX_OK) return s;
  if (pl.swap)
    return c_bf16 ? tc_launch<true, bf16>(w.map, am, tp, pl.grid, pl.smem, st) : tc_launch<true, float>(w.map, am, tp, pl.grid, pl.smem, st);
  return c_bf16 ? tc_launch<false, bf16>(am, w.map, tp, pl.grid, pl.smem, st) : tc_launch<false, float>(am, w.map, tp, pl.grid, pl.smem, st);
}

}  // namespace lvx
