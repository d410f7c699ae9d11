// GEMM interface shared by the FMA-pipe (fp32 parity mode) and tcgen05 (bf16 mode) kernels.
//
//   C[m, n] = epilogue( alpha * sum_k A'[m, k] * W[n, k] )
//
// A' is the activation matrix, row-major, optionally a multi-tap view for conv-as-GEMM: with taps > 1,
// k = tap * tap_K + c reads A[m + tap - tap_pad, c] (rows outside [0, a_rows) read as zero; rows between
// chunks are physical zeros in the padded vocoder layout, see engine.cu).  W is the weight matrix in the
// reference's nn.Linear layout (N, K) row-major, or (K, N) row-major when w_kn is set (P.V of the vocoder
// attention).  Epilogue, in order: + bias[n]; activation; * col_scale[n]; + residual[m, n]; rows with
// row_chunk[m] < 0 are not stored.  `batch` turns grid.z into independent problems (ragged attention).
#pragma once
#include "common.cuh"

namespace lvx {

struct GemmProblem {
  long long a_off, w_off, c_off;  // element offsets into A, W, C
  int M, N, K;
  int lda, ldc;                   // per-problem leading dimensions (0 = GemmParams')
};

struct GemmParams {
  const void* A = nullptr;
  const void* W = nullptr;
  void* C = nullptr;
  const float* bias = nullptr;
  const float* col_scale = nullptr;
  const float* residual = nullptr;
  const int* row_chunk = nullptr;
  const GemmProblem* batch = nullptr;
  int n_batch = 0;
  int M = 0, N = 0, K = 0;
  int lda = 0, ldw = 0, ldc = 0, ldr = 0;
  int a_rows = 0;  // rows addressable in A (for tap shifts); 0 = M
  int a_cap = 0;   // rows physically allocated behind A (tensor-map extent of the tcgen05 path); 0 = a_rows
  int taps = 1, tap_K = 0, tap_pad = 0;
  int act = ACT_NONE;
  float alpha = 1.0f;
  int w_kn = 0;
  int max_M = 0, max_N = 0;  // batch mode: largest problem (grid sizing)
  int pdl = 0;                // launch with programmatic dependent launch (decode chain)
  int c_transposed = 0;       // tcgen05 swap mode only: store C[n, m] instead of C[m, n]
  const char* tag = nullptr;  // profiler label (host only)
  int kdup = 1;               // host only: K holds this many side-by-side copies of the algorithmic K (FLOP accounting)
  int cta_budget = 0;        // tcgen05 path: SMs this launch should aim to fill (0 = all); see decode lanes
};

template <typename TC>
__device__ __forceinline__ void gemm_epilogue_store4(const GemmParams& p, TC* C, int ldc, int N, int m, int n,
                                                     float4 acc) {
  float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int nn = n + i;
    if (nn < N) {
      float x = v[i] * p.alpha;
      if (p.bias) x += p.bias[nn];
      x = apply_act(x, p.act);
      if (p.col_scale) x *= p.col_scale[nn];
      if (p.residual) x += p.residual[(size_t)m * p.ldr + nn];
      v[i] = x;
    }
  }
  TC* dst = C + (size_t)m * ldc + n;
  if (n + 3 < N && (ldc & 3) == 0) {
    store4(dst, make_float4(v[0], v[1], v[2], v[3]));
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (n + i < N) store1(dst + i, v[i]);
  }
}

// ---------------------------------------------------------------------------------------------------
// FMA-pipe GEMM.  BM x BN x 16 tiles, 256 threads, (BM/16) x (BN/16) register tile per thread, register
// prefetch + double-buffered shared memory.  fp32 accumulate; A / W may be fp32 or bf16 storage.
// ---------------------------------------------------------------------------------------------------
template <int BM, int BN, typename TA, typename TW, typename TC>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmParams p) {
  constexpr int BK = 16;
  constexpr int TM = BM / 16, TN = BN / 16;
  constexpr int NV = TN / 4;       // float4 column groups per thread
  constexpr int A_V4 = BM * 4 / 256;  // float4 loads per thread per tile
  constexpr int W_V4 = BN * 4 / 256;
  static_assert(TN % 4 == 0 && A_V4 >= 1 && W_V4 >= 1, "tile");
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Ws[2][BK][BN + 4];

  const TA* A = reinterpret_cast<const TA*>(p.A);
  const TW* W = reinterpret_cast<const TW*>(p.W);
  TC* C = reinterpret_cast<TC*>(p.C);
  int M = p.M, N = p.N, K = p.K;
  int lda = p.lda, ldc = p.ldc;
  if (p.batch) {
    const GemmProblem pr = p.batch[blockIdx.z];
    A += pr.a_off;
    W += pr.w_off;
    C += pr.c_off;
    M = pr.M;
    N = pr.N;
    K = pr.K;
    if (pr.lda) lda = pr.lda;
    if (pr.ldc) ldc = pr.ldc;
  }
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  if (m0 >= M || n0 >= N) return;
  const int a_rows = p.a_rows ? p.a_rows : M;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[A_V4], rw[W_V4];
  const int nk = (K + BK - 1) / BK;

  auto load_tiles = [&](int kb) {
    const int k0 = kb * BK;
    int shift = 0, c0 = k0;
    if (p.taps > 1) {
      const int tap = k0 / p.tap_K;
      shift = tap - p.tap_pad;
      c0 = k0 - tap * p.tap_K;
    }
#pragma unroll
    for (int i = 0; i < A_V4; ++i) {
      const int idx = tid + i * 256, r = idx >> 2, j = idx & 3;
      const int row = m0 + r + shift, kk = k0 + 4 * j;
      if (m0 + r < M && row >= 0 && row < a_rows && kk < K)
        ra[i] = load4(A + (size_t)row * lda + c0 + 4 * j);
      else
        ra[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (!p.w_kn) {
#pragma unroll
      for (int i = 0; i < W_V4; ++i) {
        const int idx = tid + i * 256, r = idx >> 2, j = idx & 3;
        const int kk = k0 + 4 * j;
        if (n0 + r < N && kk < K)
          rw[i] = load4(W + (size_t)(n0 + r) * p.ldw + kk);
        else
          rw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      // W is (K, N) row-major: float4 along n
#pragma unroll
      for (int i = 0; i < W_V4; ++i) {
        const int idx = tid + i * 256, kr = idx / (BN / 4), nc = (idx % (BN / 4)) * 4;
        const int kk = k0 + kr, nn = n0 + nc;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kk < K) {
          const TW* src = W + (size_t)kk * p.ldw + nn;
          if (nn + 3 < N && (p.ldw & 3) == 0) {
            v = load4(src);
          } else {
            if (nn + 0 < N) v.x = load1(src + 0);
            if (nn + 1 < N) v.y = load1(src + 1);
            if (nn + 2 < N) v.z = load1(src + 2);
            if (nn + 3 < N) v.w = load1(src + 3);
          }
        }
        rw[i] = v;
      }
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_V4; ++i) {
      const int idx = tid + i * 256, r = idx >> 2, j = idx & 3;
      As[buf][4 * j + 0][r] = ra[i].x;
      As[buf][4 * j + 1][r] = ra[i].y;
      As[buf][4 * j + 2][r] = ra[i].z;
      As[buf][4 * j + 3][r] = ra[i].w;
    }
    if (!p.w_kn) {
#pragma unroll
      for (int i = 0; i < W_V4; ++i) {
        const int idx = tid + i * 256, r = idx >> 2, j = idx & 3;
        Ws[buf][4 * j + 0][r] = rw[i].x;
        Ws[buf][4 * j + 1][r] = rw[i].y;
        Ws[buf][4 * j + 2][r] = rw[i].z;
        Ws[buf][4 * j + 3][r] = rw[i].w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < W_V4; ++i) {
        const int idx = tid + i * 256, kr = idx / (BN / 4), nc = (idx % (BN / 4)) * 4;
        *reinterpret_cast<float4*>(&Ws[buf][kr][nc]) = rw[i];
      }
    }
  };

  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) load_tiles(kb + 1);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][ty * TM + i]);
        a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
      }
#pragma unroll
      for (int v4 = 0; v4 < NV; ++v4) {
        const float4 v = *reinterpret_cast<const float4*>(&Ws[buf][k][v4 * (BN / NV) + tx * 4]);
        b[4 * v4] = v.x; b[4 * v4 + 1] = v.y; b[4 * v4 + 2] = v.z; b[4 * v4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kb + 1 < nk) {
      store_tiles(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    if (m >= M) continue;
    if (p.row_chunk && p.row_chunk[m] < 0) continue;
#pragma unroll
    for (int v4 = 0; v4 < NV; ++v4) {
      const int n = n0 + v4 * (BN / NV) + tx * 4;
      if (n >= N) continue;
      gemm_epilogue_store4<TC>(p, C, ldc, N, m, n,
                               make_float4(acc[i][4 * v4], acc[i][4 * v4 + 1], acc[i][4 * v4 + 2], acc[i][4 * v4 + 3]));
    }
  }
}

template <typename TA, typename TW, typename TC>
inline cudaError_t launch_gemm_simt(const GemmParams& p, cudaStream_t st) {
  const int M = p.batch ? p.max_M : p.M, N = p.batch ? p.max_N : p.N;
  const int z = p.batch ? p.n_batch : 1;
  if (M <= 0 || N <= 0 || z <= 0) return cudaSuccess;
  // small problems: 64x64 tiles so that more SMs take part; otherwise 128x128
  const long long tiles128 = (long long)ceil_div(M, 128) * ceil_div(N, 128) * z;
  if (M <= 64 || tiles128 < 148) {
    dim3 grid(ceil_div(N, 64), ceil_div(M, 64), z);
    gemm_simt_kernel<64, 64, TA, TW, TC><<<grid, 256, 0, st>>>(p);
  } else {
    dim3 grid(ceil_div(N, 128), ceil_div(M, 128), z);
    gemm_simt_kernel<128, 128, TA, TW, TC><<<grid, 256, 0, st>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace lvx
