// llmvox_b200 engine: the C ABI of include/llmvox_b200.h on top of the kernels in this directory.
//
// One engine per GPU.  Host code here only sequences kernels on the caller's stream; the decode /
// vocode entry points never synchronise the host.  The reference interfaces each entry point replaces are
// cited in the header.
#include "../../include/llmvox_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "decode_kernels.cuh"
#include "gemm.cuh"
#include "tc_gemm.cuh"
#include "cluster_decode.cuh"
#include "text_kernels.cuh"

#include <deque>
#include "vocoder_kernels.cuh"

namespace lvx {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

enum DT { F32 = 0, B16 = 1 };
static inline size_t dt_size(DT t) { return t == F32 ? 4 : 2; }

struct Tensor {
  std::vector<int64_t> shape;
  size_t n = 0;
  float* d = nullptr;
  bool loaded = false;
};

// A GEMM weight in the (N, K) row-major layout the kernels read, in both storage types.
struct GemmW {
  float* f32 = nullptr;
  bf16* b16 = nullptr;
  int N = 0, K = 0, ld = 0;
  TmaDesc tma;  // bf16 mode: tensor map of the (N, K) matrix
};

__global__ void cast_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(in[i]);
}
// conv weight (O, C, T) -> tap-major GEMM weight (O, T*C): out[o][t*C + c] = in[o][c][t]
__global__ void conv_to_tapmajor_kernel(const float* __restrict__ in, float* __restrict__ out, int O, int C, int T) {
  const size_t n = (size_t)O * C * T;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int t = (int)((i / C) % T);
    const int o = (int)(i / ((size_t)C * T));
    out[i] = in[((size_t)o * C + c) * T + t];
  }
}
// depthwise weight (C, 1, 7) -> (7, C)
__global__ void dw_to_tapmajor_kernel(const float* __restrict__ in, float* __restrict__ out, int C, int T) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C * T) out[(i % T) * C + i / T] = in[i];
}
__global__ void set_ctx_kernel(const int* __restrict__ slots, const int* __restrict__ pos, int n, SessionState st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) st.ctx_len[slots[i]] = pos[i] + 1;
}
// page-table patch list: (slot, index, page) triples
__global__ void patch_pages_kernel(const int* __restrict__ upd, int n, SessionState st) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) st.page_table[(size_t)upd[3 * i] * st.max_pages + upd[3 * i + 1]] = upd[3 * i + 2];
}
// padded codebook gather: row r of the padded layout <- codebook[codes[out0 + r - row0]] or zeros
template <typename TOut>
__global__ void gather_codebook_padded_kernel(const int* __restrict__ codes, const float* __restrict__ codebook,
                                              int width, int n_codes, const int* __restrict__ row_chunk,
                                              const ChunkInfo* __restrict__ chunks, TOut* __restrict__ out) {
  const int r = blockIdx.x, ch = row_chunk[r];
  TOut* o = out + (size_t)r * width;
  if (ch < 0) {
    for (int c = threadIdx.x * 4; c < width; c += blockDim.x * 4) store4(o + c, make_float4(0.f, 0.f, 0.f, 0.f));
    return;
  }
  const ChunkInfo ci = chunks[ch];
  int code = codes[ci.out0 + (r - ci.row0)];
  code = min(max(code, 0), n_codes - 1);
  const float* src = codebook + (size_t)code * width;
  for (int c = threadIdx.x * 4; c < width; c += blockDim.x * 4) store4(o + c, load4(src + c));
}
// padded feature copy: row r <- feats[out0 + r - row0] (packed, channels-last fp32) or zeros
template <typename TOut>
__global__ void copy_feats_padded_kernel(const float* __restrict__ feats, int width, const int* __restrict__ row_chunk,
                                         const ChunkInfo* __restrict__ chunks, TOut* __restrict__ out) {
  const int r = blockIdx.x, ch = row_chunk[r];
  TOut* o = out + (size_t)r * width;
  if (ch < 0) {
    for (int c = threadIdx.x * 4; c < width; c += blockDim.x * 4) store4(o + c, make_float4(0.f, 0.f, 0.f, 0.f));
    return;
  }
  const ChunkInfo ci = chunks[ch];
  const float* src = feats + (size_t)(ci.out0 + (r - ci.row0)) * width;
  for (int c = threadIdx.x * 4; c < width; c += blockDim.x * 4) store4(o + c, load4(src + c));
}
__global__ void gather_code_ranges_kernel(const int* __restrict__ slots, const int* __restrict__ starts,
                                          const int* __restrict__ offs, SessionState st, int* __restrict__ out) {
  const int b = blockIdx.x, slot = slots[b];
  const int cnt = offs[b + 1] - offs[b];
  for (int j = threadIdx.x; j < cnt; j += blockDim.x)
    out[offs[b] + j] = st.codes[(size_t)slot * st.max_context + starts[b] + j];
}
}  // namespace lvx

using namespace lvx;

struct lvx_engine {
  lvx_config cfg;
  int device = 0;
  bool finalized = false;
  int64_t launches = 0;
  int64_t bytes = 0;
  std::vector<void*> allocs;
  std::map<std::string, Tensor> w;

  // ---- GPT derived
  struct Layer {
    float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
    GemmW attn, proj, fc, proj2;
    GemmW attn_x2, proj_x2, fc_x2, proj2_x2;   // exact mode: LN-folded bf16 weights laid out twice along K
    float *attn_b, *proj_b, *fc_b, *proj2_b;
  };
  std::vector<Layer> layers;
  float *lnf_w = nullptr, *lnf_b = nullptr;
  GemmW lm_head, lm_head_x2;
  // ---- vocoder derived
  struct Res {
    float *n1w, *n1b, *n2w, *n2b, *c1b, *c2b;
    GemmW c1, c2;
  };
  struct CNX {
    float *dw_w, *dw_b, *scale, *shift, *b1, *b2, *gamma;
    GemmW pw1, pw2;
  };
  GemmW embed;
  float* embed_b = nullptr;
  Res res[4];
  float *at_nw = nullptr, *at_nb = nullptr, *at_qkv_b = nullptr, *at_proj_b = nullptr;
  GemmW at_qkv, at_proj;
  float *pn5_w = nullptr, *pn5_b = nullptr, *norm_scale = nullptr, *norm_shift = nullptr;
  std::vector<CNX> cnx;
  float *fln_w = nullptr, *fln_b = nullptr, *head_b = nullptr, *window = nullptr;
  GemmW head, idft;
  int spec_ld = 0, raw_ld = 0;

  // ---- sessions
  SessionState st{};
  void* kv = nullptr;
  long long pool_pages = 0;
  int max_pages = 0;
  std::vector<int> h_len, h_text_len, h_open, h_npages, free_pages, stamp;
  uint8_t *tx_bytes = nullptr, *tx_scratch = nullptr;   // device text front-end (lvx_feed_utf8): raw bytes, rewrite scratch
  int* tx_status = nullptr;
  std::vector<std::vector<int>> h_pages;
  int stamp_ctr = 0;
  int *d_slots = nullptr, *d_aux = nullptr, *d_ids = nullptr;   // staging of the control calls (open / feed / gather)
  int Bp = 0;   // decode workspace rows (max_batch padded to the tensor-core tile)
  // ---- vocoder workspace
  int R_max = 0;
  long long score_cap = 0;
  ChunkInfo* d_chunks = nullptr;
  GemmProblem *d_prob_s = nullptr, *d_prob_pv = nullptr;
  TcGroup *d_grp_s = nullptr, *d_grp_pv = nullptr;   // grouped tcgen05 launches of the long chunks' attention
  int *row_chunk = nullptr, *code_rows = nullptr;
  int max_chunks = 0;
  bf16* v_vt = nullptr;   // V^T of the pos_net attention: [voc_dim][R_max] bf16
  void *v_feats = nullptr, *v_h = nullptr, *v_big = nullptr, *v_spec = nullptr, *v_P = nullptr, *v_h3 = nullptr;
  float *v_x = nullptr, *v_t = nullptr, *v_raw = nullptr, *v_frames = nullptr, *v_S = nullptr;
  float2* v_stats = nullptr;
  TcWorkspace tcw;

  // ---- decode lanes: independent workspaces so that groups of sessions can run their (latency-bound) dependent
  // chains concurrently on different streams; each lane caches CUDA graphs of its decode iteration
  struct StepGraph {
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;
  };
  struct Lane {
    unsigned* d_bar = nullptr;
    long long* d_trace = nullptr;
    int *d_slots = nullptr, *d_upd = nullptr, *d_pos = nullptr;
    float *x = nullptr, *qkv = nullptr, *logits = nullptr;
    void *h = nullptr, *y = nullptr, *g = nullptr;
    float *yf = nullptr, *gf = nullptr;   // exact mode: fp32 attention output / fc output before the hi | lo split
    std::map<std::string, StepGraph> graphs;
    int cta_budget = 0;   // CTAs the kernel-per-op chain of the current call sizes its grids for; 0 = lane_cta_budget
  };
  std::vector<Lane> lanes;
  int lane_cta_budget = 0;
  cudaStream_t gstream = nullptr;   // capture only: graphs are launched on the caller's stream
  // helper streams: the per-chunk attention GEMMs of long vocoder chunks are independent and run side by side
  static constexpr int kAux = 4;
  cudaStream_t aux[kAux] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join[kAux] = {nullptr, nullptr, nullptr, nullptr};
  bool use_graphs = true;
  bool use_pdl = true;      // programmatic dependent launch along the decode chain (LLMVOX_B200_NO_PDL=1 disables)
  // cluster-resident decode kernel (cluster_decode.cuh): per-rank weight streams + table norms, built on first use
  bool use_cluster = false;
  bool cd_ready = false;
  uint8_t* cd_stream = nullptr;    // 16-CTA clusters (latency variant)
  long long cd_stream_bytes = 0;
  uint8_t* cd_stream8 = nullptr;   // 8-CTA clusters (throughput variant)
  long long cd_stream8_bytes = 0;
  float *cd_text_ss = nullptr, *cd_code_ss = nullptr;
  int cd_max_clusters = 0, cd_max_clusters8 = 0;   // co-resident clusters of each variant (cudaOccupancyMaxActiveClusters)
  // spread a call's sessions over all co-resident clusters instead of filling clusters to 16 (LLMVOX_B200_CD_SPREAD=1).
  // Measured (profiles/r02_cluster_decode.md): 64 sessions as 7 clusters of 9-10 run at 148.7 us / iteration at T = 20..120
  // (4 clusters of 16: 149.8) but 202.8 vs 184.1 at T = 110..210 -- seven weight streams contend in L2 with the K/V
  // reads -- so filling stays the default.
  bool cd_spread = false;
  // launches that may still be running: never more clusters in flight than are co-resident (see cluster_launch)
  struct CdInflight {
    cudaEvent_t ev;
    cudaStream_t st;
    int clusters;
  };
  std::deque<CdInflight> cd_inflight;

  // ---- pinned staging ring for the small host -> device uploads of the vocoder (chunk layout, problem tables): a copy
  // from pageable memory would block the calling thread until the DMA has read it; a slot is reused only after the copy
  // that read it has completed (event per slot)
  struct StageSlot {
    void* h = nullptr;
    size_t cap = 0;
    cudaEvent_t ev = nullptr;
    bool busy = false;
  };
  StageSlot stage[8];
  int stage_next = 0;
  // copies `bytes` from src to d_dst on `st` through the ring
  int upload(void* d_dst, const void* src, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return LVX_OK;
    StageSlot& sl = stage[stage_next];
    stage_next = (stage_next + 1) % 8;
    if (sl.busy) {
      LVX_CUDA(cudaEventSynchronize(sl.ev));
      sl.busy = false;
    }
    if (sl.cap < bytes) {
      if (sl.h) cudaFreeHost(sl.h);
      sl.cap = std::max<size_t>(bytes, 64 * 1024);
      LVX_CUDA(cudaHostAlloc(&sl.h, sl.cap, cudaHostAllocDefault));
    }
    if (!sl.ev) LVX_CUDA(cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming));
    memcpy(sl.h, src, bytes);
    LVX_CUDA(cudaMemcpyAsync(d_dst, sl.h, bytes, cudaMemcpyHostToDevice, st));
    LVX_CUDA(cudaEventRecord(sl.ev, st));
    sl.busy = true;
    return LVX_OK;
  }

  // ---- optional per-launch profiler (lvx_profile_enable)
  struct ProfRec {
    const char* name;
    cudaEvent_t a, b;
    double flops, bytes;
  };
  bool prof_on = false;
  bool prof_detail = false;   // lvx_profile_enable(2): label vocoder GEMMs by role
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  cudaEvent_t get_event() {
    if (!ev_pool.empty()) {
      cudaEvent_t ev = ev_pool.back();
      ev_pool.pop_back();
      return ev;
    }
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    return ev;
  }

  // storage type of GEMM operands (EXACT: bf16 hi | lo pairs for the decoder, plain bf16 for the vocoder)
  DT adt() const { return cfg.precision == LVX_PRECISION_FP32 ? F32 : B16; }
  bool exact() const { return cfg.precision == LVX_PRECISION_EXACT; }
  DT kvdt() const { return cfg.precision == LVX_PRECISION_BF16 ? B16 : F32; }
};

struct ProfScope {
  lvx_engine* e;
  cudaStream_t st;
  size_t idx = (size_t)-1;
  ProfScope(lvx_engine* e_, const char* name, cudaStream_t st_, double flops = 0, double bytes = 0) : e(e_), st(st_) {
    if (!e->prof_on) return;
    lvx_engine::ProfRec r{name, e->get_event(), e->get_event(), flops, bytes};
    cudaEventRecord(r.a, st);
    idx = e->prof.size();
    e->prof.push_back(r);
  }
  ~ProfScope() {
    if (idx != (size_t)-1) cudaEventRecord(e->prof[idx].b, st);
  }
};
#define PROF(e, name, st) ProfScope _prof_scope_##__LINE__((e), (name), (st))

#define LAUNCHED(e)                                                                                  \
  do {                                                                                               \
    (e)->launches++;                                                                                 \
    cudaError_t _le = cudaGetLastError();                                                            \
    if (_le != cudaSuccess) {                                                                        \
      lvx::set_error(std::string("kernel launch: ") + cudaGetErrorString(_le) + " (" + __FILE__ + ":" + \
                     std::to_string(__LINE__) + ")");                                                \
      return LVX_ERR_CUDA;                                                                           \
    }                                                                                                \
  } while (0)

template <typename T>
static int dev_alloc(lvx_engine* e, T** p, size_t count) {
  void* q = nullptr;
  size_t bytes = std::max<size_t>(count * sizeof(T), 256);
  LVX_CUDA(cudaMalloc(&q, bytes));
  LVX_CUDA(cudaMemset(q, 0, bytes));
  e->allocs.push_back(q);
  e->bytes += (int64_t)bytes;
  *p = reinterpret_cast<T*>(q);
  return LVX_OK;
}
static int dev_alloc_bytes(lvx_engine* e, void** p, size_t bytes) {
  unsigned char* q = nullptr;
  LVX_TRY(dev_alloc<unsigned char>(e, &q, bytes));
  *p = q;
  return LVX_OK;
}

static void expect(lvx_engine* e, const std::string& name, std::vector<int64_t> shape) {
  Tensor t;
  t.shape = shape;
  t.n = 1;
  for (auto s : shape) t.n *= (size_t)s;
  e->w[name] = t;
}

static void declare_tensors(lvx_engine* e) {
  const lvx_config& c = e->cfg;
  const int64_t C = c.n_embd, D = c.voc_dim, I = c.voc_inter;
  expect(e, "text_table", {c.text_vocab, c.text_dim});
  expect(e, "transformer.wpe.weight", {c.block_size, C});
  for (int i = 0; i < c.n_layer; ++i) {
    const std::string p = "transformer.h." + std::to_string(i) + ".";
    expect(e, p + "ln_1.weight", {C});
    expect(e, p + "attn.c_attn.weight", {3 * C, C});
    expect(e, p + "attn.c_proj.weight", {C, C});
    expect(e, p + "ln_2.weight", {C});
    expect(e, p + "mlp.c_fc.weight", {4 * C, C});
    expect(e, p + "mlp.c_proj.weight", {C, 4 * C});
    if (c.bias) {
      expect(e, p + "ln_1.bias", {C});
      expect(e, p + "ln_2.bias", {C});
      expect(e, p + "attn.c_attn.bias", {3 * C});
      expect(e, p + "attn.c_proj.bias", {C});
      expect(e, p + "mlp.c_fc.bias", {4 * C});
      expect(e, p + "mlp.c_proj.bias", {C});
    }
  }
  expect(e, "transformer.ln_f.weight", {C});
  if (c.bias) expect(e, "transformer.ln_f.bias", {C});
  expect(e, "lm_head.weight", {c.vocab_size, C});
  expect(e, "feature_extractor.encodec.quantizer.vq.layers.0._codebook.embed", {c.n_codes, c.code_dim});
  expect(e, "backbone.embed.weight", {D, c.code_dim, 7});
  expect(e, "backbone.embed.bias", {D});
  for (int i : {0, 1, 3, 4}) {
    const std::string p = "backbone.pos_net." + std::to_string(i) + ".";
    for (const char* n : {"1", "2"}) {
      expect(e, p + "norm" + n + ".weight", {D});
      expect(e, p + "norm" + n + ".bias", {D});
      expect(e, p + "conv" + n + ".weight", {D, D, 3});
      expect(e, p + "conv" + n + ".bias", {D});
    }
  }
  expect(e, "backbone.pos_net.2.norm.weight", {D});
  expect(e, "backbone.pos_net.2.norm.bias", {D});
  for (const char* n : {"q", "k", "v", "proj_out"}) {
    expect(e, std::string("backbone.pos_net.2.") + n + ".weight", {D, D, 1});
    expect(e, std::string("backbone.pos_net.2.") + n + ".bias", {D});
  }
  expect(e, "backbone.pos_net.5.weight", {D});
  expect(e, "backbone.pos_net.5.bias", {D});
  expect(e, "backbone.norm.scale.weight", {c.voc_ada_rows, D});
  expect(e, "backbone.norm.shift.weight", {c.voc_ada_rows, D});
  for (int i = 0; i < c.voc_layers; ++i) {
    const std::string p = "backbone.convnext." + std::to_string(i) + ".";
    expect(e, p + "dwconv.weight", {D, 1, 7});
    expect(e, p + "dwconv.bias", {D});
    expect(e, p + "norm.scale.weight", {c.voc_ada_rows, D});
    expect(e, p + "norm.shift.weight", {c.voc_ada_rows, D});
    expect(e, p + "pwconv1.weight", {I, D});
    expect(e, p + "pwconv1.bias", {I});
    expect(e, p + "pwconv2.weight", {D, I});
    expect(e, p + "pwconv2.bias", {D});
    expect(e, p + "gamma", {D});
  }
  expect(e, "backbone.final_layer_norm.weight", {D});
  expect(e, "backbone.final_layer_norm.bias", {D});
  expect(e, "head.out.weight", {c.n_fft + 2, D});
  expect(e, "head.out.bias", {c.n_fft + 2});
  expect(e, "head.istft.window", {c.n_fft});
}

extern "C" const char* lvx_last_error(void) { return g_last_error.c_str(); }
extern "C" int lvx_version(void) { return 100; }

extern "C" int lvx_config_default(lvx_config* cfg) {
  LVX_CHECK(cfg != nullptr, LVX_ERR_INVALID, "cfg is NULL");
  memset(cfg, 0, sizeof(*cfg));
  cfg->n_layer = 4; cfg->n_head = 8; cfg->n_embd = 768; cfg->block_size = 8192; cfg->vocab_size = 4096; cfg->bias = 0;
  cfg->text_vocab = 386; cfg->text_dim = 256; cfg->code_dim = 512; cfg->n_codes = 4096;
  cfg->voc_dim = 768; cfg->voc_inter = 2304; cfg->voc_layers = 12; cfg->voc_ada_rows = 4; cfg->n_fft = 1280; cfg->hop = 320;
  cfg->max_sessions = 256;
  cfg->max_context = 1024;
  cfg->kv_page_tokens = 16;
  cfg->kv_pages = 0;
  cfg->max_batch = 256;
  cfg->max_vocode_frames = 32768;
  cfg->precision = LVX_PRECISION_FP32;
  cfg->pad_token_id = 384;
  cfg->eoa_token_id = 453;
  cfg->decode_lanes = 1;
  return LVX_OK;
}

static int engine_alloc(lvx_engine* e) {
  const lvx_config& c = e->cfg;
  const int S = c.max_sessions, B = c.max_batch, C = c.n_embd;
  const DT a = e->adt();
  // sessions
  e->max_pages = ceil_div(c.max_context, c.kv_page_tokens);
  e->pool_pages = c.kv_pages > 0 ? c.kv_pages : (long long)S * e->max_pages;
  e->st.max_context = c.max_context;
  e->st.max_pages = e->max_pages;
  LVX_TRY(dev_alloc(e, &e->st.ctx_len, S));
  LVX_TRY(dev_alloc(e, &e->st.text_len, S));
  LVX_TRY(dev_alloc(e, &e->st.text_ids, (size_t)S * c.max_context));
  LVX_TRY(dev_alloc(e, &e->st.codes, (size_t)S * c.max_context));
  LVX_TRY(dev_alloc(e, &e->st.page_table, (size_t)S * e->max_pages));
  LVX_TRY(dev_alloc(e, &e->st.eoa_pos, S));
  LVX_CUDA(cudaMemset(e->st.eoa_pos, 0xFF, (size_t)S * sizeof(int)));
  e->st.eoa_id = c.eoa_token_id;
  const size_t kv_elems = (size_t)c.n_layer * 2 * e->pool_pages * c.kv_page_tokens * C;
  LVX_TRY(dev_alloc_bytes(e, &e->kv, kv_elems * dt_size(e->kvdt())));
  e->h_len.assign(S, 0);
  e->h_text_len.assign(S, 0);
  e->h_open.assign(S, 0);
  e->stamp.assign(S, 0);
  e->h_pages.assign(S, {});
  e->free_pages.resize(e->pool_pages);
  for (long long i = 0; i < e->pool_pages; ++i) e->free_pages[i] = (int)(e->pool_pages - 1 - i);
  LVX_TRY(dev_alloc(e, &e->d_slots, B));
  LVX_TRY(dev_alloc(e, &e->d_aux, 2 * B + 2));
  LVX_TRY(dev_alloc(e, &e->d_ids, (size_t)B * c.max_context));
  // decode workspaces, one per lane (rows padded to the tensor-core tile so TMA boxes stay in bounds)
  const int Bp = ceil_div(B, 128) * 128;
  e->Bp = Bp;
  e->lanes.resize(std::max(1, c.decode_lanes));
  for (auto& ln : e->lanes) {
    LVX_TRY(dev_alloc(e, &ln.d_slots, B));
    LVX_TRY(dev_alloc(e, &ln.d_bar, 64));
    LVX_TRY(dev_alloc(e, &ln.d_trace, 256));
    LVX_TRY(dev_alloc(e, &ln.d_pos, B));
    LVX_TRY(dev_alloc(e, &ln.d_upd, (size_t)3 * B * e->max_pages));
    LVX_TRY(dev_alloc(e, &ln.x, (size_t)Bp * C));
    LVX_TRY(dev_alloc(e, &ln.qkv, (size_t)Bp * 3 * C));
    LVX_TRY(dev_alloc(e, &ln.logits, (size_t)Bp * c.vocab_size));
    const size_t act_bytes = e->exact() ? 4 : dt_size(a);   // exact: 2 bf16 (hi | lo) per element
    LVX_TRY(dev_alloc_bytes(e, &ln.h, (size_t)Bp * C * act_bytes));
    LVX_TRY(dev_alloc_bytes(e, &ln.y, (size_t)Bp * C * act_bytes));
    LVX_TRY(dev_alloc_bytes(e, &ln.g, (size_t)Bp * 4 * C * act_bytes));
    if (e->exact()) {
      LVX_TRY(dev_alloc(e, &ln.yf, (size_t)Bp * C));
      LVX_TRY(dev_alloc(e, &ln.gf, (size_t)Bp * 4 * C));
    }
  }
  // vocoder workspace
  const int D = c.voc_dim, I = c.voc_inter;
  // padded rows of one launch group: frames + (alignment to 8 + ROW_PAD) per chunk; plan_groups also closes a group
  // when its padded rows would exceed this capacity (batches of very many tiny chunks)
  e->R_max = ceil_div(c.max_vocode_frames + c.max_vocode_frames / 4 + 4096, 128) * 128 + 128;
  e->max_chunks = c.max_vocode_frames;
  e->raw_ld = (c.n_fft + 2 + 3) & ~3;
  e->spec_ld = ceil_div(c.n_fft + 2, 64) * 64;
  // attention scores / probabilities of a launch group: [padded rows][sld], sld = the group's longest chunk rounded up to 64
  e->score_cap = std::max((long long)c.max_vocode_frames * 1280, (long long)e->R_max * 1280);
  const size_t R = e->R_max;
  LVX_TRY(dev_alloc(e, &e->d_chunks, e->max_chunks));
  LVX_TRY(dev_alloc(e, &e->d_prob_s, e->max_chunks));
  LVX_TRY(dev_alloc(e, &e->d_prob_pv, e->max_chunks));
  LVX_TRY(dev_alloc(e, &e->d_grp_s, e->max_chunks));
  LVX_TRY(dev_alloc(e, &e->d_grp_pv, e->max_chunks));
  LVX_TRY(dev_alloc(e, &e->row_chunk, R));
  LVX_TRY(dev_alloc(e, &e->code_rows, R));
  LVX_TRY(dev_alloc_bytes(e, &e->v_feats, R * c.code_dim * dt_size(a)));
  LVX_TRY(dev_alloc(e, &e->v_x, R * D));
  LVX_TRY(dev_alloc(e, &e->v_t, R * D));
  LVX_TRY(dev_alloc_bytes(e, &e->v_h, R * D * dt_size(a)));
  LVX_TRY(dev_alloc_bytes(e, &e->v_big, R * std::max<size_t>((size_t)std::max(I, 3 * D) * dt_size(a), (size_t)e->spec_ld * 4)));
  LVX_TRY(dev_alloc(e, &e->v_raw, R * e->raw_ld));
  LVX_TRY(dev_alloc(e, &e->v_frames, R * c.n_fft));
  LVX_TRY(dev_alloc(e, &e->v_S, (size_t)e->score_cap));
  LVX_TRY(dev_alloc_bytes(e, &e->v_P, (size_t)e->score_cap * dt_size(a)));
  LVX_TRY(dev_alloc(e, &e->v_stats, (size_t)e->max_chunks * 32));
  if (a == B16) {
    LVX_TRY(dev_alloc(e, &e->v_vt, (size_t)D * R));
    LVX_TRY(dev_alloc_bytes(e, &e->v_spec, R * 3 * e->spec_ld * sizeof(bf16)));
    LVX_TRY(dev_alloc_bytes(e, &e->v_h3, R * 3 * D * sizeof(bf16)));
  } else {
    LVX_TRY(dev_alloc_bytes(e, &e->v_spec, R * e->spec_ld * sizeof(float)));
  }
  return LVX_OK;
}

extern "C" int lvx_engine_create(const lvx_config* cfg, int device, lvx_engine** out) {
  LVX_CHECK(cfg && out, LVX_ERR_INVALID, "cfg / out is NULL");
  const lvx_config& c = *cfg;
  LVX_CHECK(c.n_embd == 768 && c.voc_dim == 768, LVX_ERR_INVALID,
            "kernels are specialised for n_embd == voc_dim == 768 (english-tiny / frame75)");
  LVX_CHECK(c.n_head > 0 && c.n_embd % c.n_head == 0 && c.n_embd / c.n_head == 96, LVX_ERR_INVALID,
            "kernels are specialised for head_dim 96");
  LVX_CHECK(c.text_dim + c.code_dim == c.n_embd && c.text_dim % 4 == 0, LVX_ERR_INVALID, "text_dim + code_dim != n_embd");
  LVX_CHECK(c.vocab_size > 0 && c.vocab_size <= 4096 && c.vocab_size % 4 == 0, LVX_ERR_INVALID, "vocab_size must be <= 4096");
  LVX_CHECK(c.n_fft == 4 * c.hop && c.n_fft % 64 == 0, LVX_ERR_INVALID, "iSTFT kernels need n_fft == 4 * hop");
  LVX_CHECK(c.max_sessions > 0 && c.max_batch > 0 && c.max_batch <= c.max_sessions, LVX_ERR_INVALID, "bad session capacity");
  LVX_CHECK(c.max_context > 0 && c.max_context <= c.block_size, LVX_ERR_INVALID, "max_context must be <= block_size");
  LVX_CHECK(c.kv_page_tokens > 0 && c.max_vocode_frames > 0, LVX_ERR_INVALID, "bad capacity");
  LVX_CHECK(c.precision == LVX_PRECISION_BF16 || c.kv_page_tokens % 8 == 0, LVX_ERR_INVALID,
            "the fp32 KV pool is laid out in 8-token half pages: kv_page_tokens must be a multiple of 8");
  LVX_CHECK(c.voc_inter % 64 == 0 && c.code_dim % 64 == 0, LVX_ERR_INVALID, "voc_inter / code_dim must be multiples of 64");
  LVX_CHECK(c.precision == LVX_PRECISION_FP32 || c.precision == LVX_PRECISION_BF16 || c.precision == LVX_PRECISION_EXACT,
            LVX_ERR_INVALID, "bad precision");
  LVX_CHECK(c.precision != LVX_PRECISION_EXACT || !c.bias, LVX_ERR_INVALID,
            "exact mode folds the LayerNorm weights into the GEMMs and needs bias == 0 (english-tiny)");
  LVX_CHECK(c.decode_lanes >= 0 && c.decode_lanes <= 16, LVX_ERR_INVALID, "decode_lanes must be in [0, 16]");
  int ndev = 0;
  LVX_CUDA(cudaGetDeviceCount(&ndev));
  LVX_CHECK(device >= 0 && device < ndev, LVX_ERR_INVALID, "no such CUDA device");
  LVX_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  LVX_CUDA(cudaGetDeviceProperties(&prop, device));
  LVX_CHECK(prop.major == 10, LVX_ERR_INVALID,
            std::string("llmvox_b200 is built for sm_100a only; device is sm_") + std::to_string(prop.major) +
                std::to_string(prop.minor));
  lvx_engine* e = new lvx_engine();
  e->cfg = c;
  e->device = device;
  declare_tensors(e);
  int s = engine_alloc(e);
  if (s == LVX_OK && c.precision != LVX_PRECISION_FP32) s = tc_init(&e->tcw, prop.multiProcessorCount);
  if (s == LVX_OK) {   // the depthwise-conv strip kernel keeps DWB_R + 6 fp32 rows (54 KB) in shared memory
    cudaError_t ce = cudaFuncSetAttribute(dwconv_adaln_bulk_kernel<float, 768>, cudaFuncAttributeMaxDynamicSharedMemorySize, DWB_SMEM);
    if (ce == cudaSuccess) ce = cudaFuncSetAttribute(dwconv_adaln_bulk_kernel<bf16, 768>, cudaFuncAttributeMaxDynamicSharedMemorySize, DWB_SMEM);
    if (ce != cudaSuccess) {
      set_error(std::string("cudaFuncSetAttribute(dwconv_adaln_bulk): ") + cudaGetErrorString(ce));
      s = LVX_ERR_CUDA;
    }
  }
  if (s == LVX_OK) {
    const char* env = getenv("LLMVOX_B200_NO_GRAPH");
    e->use_graphs = !(env && env[0] == '1');
    const char* env3 = getenv("LLMVOX_B200_NO_PDL");
    e->use_pdl = !(env3 && env3[0] == '1');
    // cluster-resident decode kernel (cluster_decode.cuh): the default greedy bf16 path; LLMVOX_B200_CLUSTER=0 selects the
    // kernel-per-op path (CUDA graphs + programmatic dependent launch) instead
    const char* env5 = getenv("LLMVOX_B200_CD_SPREAD");
    e->cd_spread = env5 && env5[0] == '1';
    const char* env4 = getenv("LLMVOX_B200_CLUSTER");
    e->use_cluster = !(env4 && env4[0] == '0');
    if (cudaStreamCreateWithFlags(&e->gstream, cudaStreamNonBlocking) != cudaSuccess) {
      set_error("could not create the engine's capture stream");
      s = LVX_ERR_CUDA;
    }
    e->lane_cta_budget = std::max(24, prop.multiProcessorCount / (int)e->lanes.size());
    bool ok = cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    for (int a = 0; a < lvx_engine::kAux && ok; ++a)
      ok = cudaStreamCreateWithFlags(&e->aux[a], cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&e->ev_join[a], cudaEventDisableTiming) == cudaSuccess;
    if (!ok && s == LVX_OK) {
      set_error("could not create the engine's helper streams");
      s = LVX_ERR_CUDA;
    }
  }
  if (s != LVX_OK) {
    lvx_engine_destroy(e);
    return s;
  }
  *out = e;
  return LVX_OK;
}

extern "C" int lvx_engine_destroy(lvx_engine* e) {
  if (!e) return LVX_OK;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  for (auto& ln : e->lanes)
    for (auto& g : ln.graphs)
      if (g.second.exec) cudaGraphExecDestroy(g.second.exec);
  if (e->gstream) cudaStreamDestroy(e->gstream);
  for (int a = 0; a < lvx_engine::kAux; ++a) {
    if (e->aux[a]) cudaStreamDestroy(e->aux[a]);
    if (e->ev_join[a]) cudaEventDestroy(e->ev_join[a]);
  }
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  for (auto& r : e->prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  for (auto& f : e->cd_inflight) e->ev_pool.push_back(f.ev);
  e->cd_inflight.clear();
  for (cudaEvent_t ev : e->ev_pool) cudaEventDestroy(ev);
  for (auto& sl : e->stage) {
    if (sl.ev) cudaEventDestroy(sl.ev);
    if (sl.h) cudaFreeHost(sl.h);
  }
  for (void* p : e->allocs) cudaFree(p);
  for (auto& kv : e->w)
    if (kv.second.d) cudaFree(kv.second.d);
  delete e;
  return LVX_OK;
}

extern "C" int lvx_load_tensor(lvx_engine* e, const char* name, const float* h_data, const int64_t* shape, int ndim) {
  LVX_CHECK(e && name && h_data && shape, LVX_ERR_INVALID, "NULL argument");
  LVX_CHECK(!e->finalized, LVX_ERR_STATE, "weights already finalised");
  auto it = e->w.find(name);
  LVX_CHECK(it != e->w.end(), LVX_ERR_INVALID, std::string("unknown tensor name: ") + name);
  Tensor& t = it->second;
  bool ok = (int)t.shape.size() == ndim;
  for (int i = 0; ok && i < ndim; ++i) ok = t.shape[i] == shape[i];
  // the position table may be shorter than block_size (only max_context rows are ever read)
  if (!ok && it->first == "transformer.wpe.weight" && ndim == 2 && shape[1] == t.shape[1] && shape[0] >= e->cfg.max_context &&
      shape[0] <= t.shape[0]) {
    t.shape[0] = shape[0];
    t.n = (size_t)shape[0] * shape[1];
    ok = true;
  }
  LVX_CHECK(ok, LVX_ERR_INVALID, std::string("shape mismatch for ") + name);
  LVX_CUDA(cudaSetDevice(e->device));
  if (!t.d) {
    LVX_CUDA(cudaMalloc(&t.d, std::max<size_t>(t.n * sizeof(float), 256)));
    e->bytes += (int64_t)(t.n * sizeof(float));
  }
  LVX_CUDA(cudaMemcpy(t.d, h_data, t.n * sizeof(float), cudaMemcpyHostToDevice));
  t.loaded = true;
  return LVX_OK;
}

static float* W(lvx_engine* e, const std::string& name) {
  auto it = e->w.find(name);
  return it == e->w.end() ? nullptr : it->second.d;
}

// Registers an (N, K) fp32 matrix as a GEMM weight; bf16 mode adds the bf16 copy and its tensor map.
static int make_gemm_w(lvx_engine* e, GemmW* g, float* f32, int N, int K, int ld, bool with_bf16 = true) {
  g->f32 = f32;
  g->N = N;
  g->K = K;
  g->ld = ld;
  if (e->adt() == B16 && with_bf16) {
    const size_t n = (size_t)N * ld;
    LVX_TRY(dev_alloc(e, &g->b16, n));
    cast_bf16_kernel<<<std::min<size_t>(4096, (n + 255) / 256), 256>>>(f32, g->b16, n);
    LAUNCHED(e);
    LVX_TRY(tc_make_desc(&g->tma, g->b16, N, K, ld));
  }
  return LVX_OK;
}

// Exact mode: bf16(W * diag(scale)) laid out twice along K (decode_kernels.cuh: fold_dup_kernel) + its tensor map.
static int make_gemm_w_x2(lvx_engine* e, GemmW* g, const float* f32, const float* scale, int N, int K, int ld) {
  g->f32 = nullptr;
  g->N = N;
  g->K = 2 * K;
  g->ld = 2 * K;
  const size_t n = (size_t)N * 2 * K;
  LVX_TRY(dev_alloc(e, &g->b16, n));
  fold_dup_kernel<<<std::min<size_t>(4096, ((size_t)N * K + 255) / 256), 256>>>(f32, scale, N, K, ld, g->b16);
  LAUNCHED(e);
  return tc_make_desc(&g->tma, g->b16, N, 2 * K, 2 * K);
}

static int make_conv_w(lvx_engine* e, GemmW* g, const std::string& name, int O, int Cin, int T) {
  float* dst = nullptr;
  LVX_TRY(dev_alloc(e, &dst, (size_t)O * Cin * T));
  const size_t n = (size_t)O * Cin * T;
  conv_to_tapmajor_kernel<<<std::min<size_t>(4096, (n + 255) / 256), 256>>>(W(e, name), dst, O, Cin, T);
  LAUNCHED(e);
  return make_gemm_w(e, g, dst, O, Cin * T, Cin * T);
}

extern "C" int lvx_finalize_weights(lvx_engine* e) {
  LVX_CHECK(e, LVX_ERR_INVALID, "engine is NULL");
  LVX_CHECK(!e->finalized, LVX_ERR_STATE, "weights already finalised");
  for (auto& kv : e->w) LVX_CHECK(kv.second.loaded, LVX_ERR_STATE, std::string("tensor not loaded: ") + kv.first);
  LVX_CUDA(cudaSetDevice(e->device));
  const lvx_config& c = e->cfg;
  const int C = c.n_embd, D = c.voc_dim, I = c.voc_inter;
  e->layers.resize(c.n_layer);
  for (int i = 0; i < c.n_layer; ++i) {
    const std::string p = "transformer.h." + std::to_string(i) + ".";
    auto& L = e->layers[i];
    L.ln1_w = W(e, p + "ln_1.weight");
    L.ln2_w = W(e, p + "ln_2.weight");
    L.ln1_b = W(e, p + "ln_1.bias");
    L.ln2_b = W(e, p + "ln_2.bias");
    L.attn_b = W(e, p + "attn.c_attn.bias");
    L.proj_b = W(e, p + "attn.c_proj.bias");
    L.fc_b = W(e, p + "mlp.c_fc.bias");
    L.proj2_b = W(e, p + "mlp.c_proj.bias");
    const bool x = e->exact();
    LVX_TRY(make_gemm_w(e, &L.attn, W(e, p + "attn.c_attn.weight"), 3 * C, C, C, !x));
    LVX_TRY(make_gemm_w(e, &L.proj, W(e, p + "attn.c_proj.weight"), C, C, C, !x));
    LVX_TRY(make_gemm_w(e, &L.fc, W(e, p + "mlp.c_fc.weight"), 4 * C, C, C, !x));
    LVX_TRY(make_gemm_w(e, &L.proj2, W(e, p + "mlp.c_proj.weight"), C, 4 * C, 4 * C, !x));
    if (x) {
      LVX_TRY(make_gemm_w_x2(e, &L.attn_x2, L.attn.f32, L.ln1_w, 3 * C, C, C));
      LVX_TRY(make_gemm_w_x2(e, &L.proj_x2, L.proj.f32, nullptr, C, C, C));
      LVX_TRY(make_gemm_w_x2(e, &L.fc_x2, L.fc.f32, L.ln2_w, 4 * C, C, C));
      LVX_TRY(make_gemm_w_x2(e, &L.proj2_x2, L.proj2.f32, nullptr, C, 4 * C, 4 * C));
    }
  }
  e->lnf_w = W(e, "transformer.ln_f.weight");
  e->lnf_b = W(e, "transformer.ln_f.bias");
  LVX_TRY(make_gemm_w(e, &e->lm_head, W(e, "lm_head.weight"), c.vocab_size, C, C, !e->exact()));
  if (e->exact()) LVX_TRY(make_gemm_w_x2(e, &e->lm_head_x2, e->lm_head.f32, e->lnf_w, c.vocab_size, C, C));

  // vocoder
  LVX_TRY(make_conv_w(e, &e->embed, "backbone.embed.weight", D, c.code_dim, 7));
  e->embed_b = W(e, "backbone.embed.bias");
  const int ridx[4] = {0, 1, 3, 4};
  for (int r = 0; r < 4; ++r) {
    const std::string p = "backbone.pos_net." + std::to_string(ridx[r]) + ".";
    auto& R = e->res[r];
    R.n1w = W(e, p + "norm1.weight"); R.n1b = W(e, p + "norm1.bias");
    R.n2w = W(e, p + "norm2.weight"); R.n2b = W(e, p + "norm2.bias");
    R.c1b = W(e, p + "conv1.bias");   R.c2b = W(e, p + "conv2.bias");
    LVX_TRY(make_conv_w(e, &R.c1, p + "conv1.weight", D, D, 3));
    LVX_TRY(make_conv_w(e, &R.c2, p + "conv2.weight", D, D, 3));
  }
  {
    const std::string p = "backbone.pos_net.2.";
    e->at_nw = W(e, p + "norm.weight");
    e->at_nb = W(e, p + "norm.bias");
    float *wq = nullptr, *bq = nullptr;
    LVX_TRY(dev_alloc(e, &wq, (size_t)3 * D * D));
    LVX_TRY(dev_alloc(e, &bq, (size_t)3 * D));
    const char* names[3] = {"q", "k", "v"};
    for (int j = 0; j < 3; ++j) {
      LVX_CUDA(cudaMemcpy(wq + (size_t)j * D * D, W(e, p + names[j] + ".weight"), (size_t)D * D * sizeof(float),
                          cudaMemcpyDeviceToDevice));
      LVX_CUDA(cudaMemcpy(bq + (size_t)j * D, W(e, p + names[j] + ".bias"), (size_t)D * sizeof(float),
                          cudaMemcpyDeviceToDevice));
    }
    e->at_qkv_b = bq;
    LVX_TRY(make_gemm_w(e, &e->at_qkv, wq, 3 * D, D, D));
    LVX_TRY(make_gemm_w(e, &e->at_proj, W(e, p + "proj_out.weight"), D, D, D));
    e->at_proj_b = W(e, p + "proj_out.bias");
  }
  e->pn5_w = W(e, "backbone.pos_net.5.weight");
  e->pn5_b = W(e, "backbone.pos_net.5.bias");
  e->norm_scale = W(e, "backbone.norm.scale.weight");
  e->norm_shift = W(e, "backbone.norm.shift.weight");
  e->cnx.resize(c.voc_layers);
  for (int i = 0; i < c.voc_layers; ++i) {
    const std::string p = "backbone.convnext." + std::to_string(i) + ".";
    auto& X = e->cnx[i];
    LVX_TRY(dev_alloc(e, &X.dw_w, (size_t)7 * D));
    dw_to_tapmajor_kernel<<<ceil_div(7 * D, 256), 256>>>(W(e, p + "dwconv.weight"), X.dw_w, D, 7);
    LAUNCHED(e);
    X.dw_b = W(e, p + "dwconv.bias");
    X.scale = W(e, p + "norm.scale.weight");
    X.shift = W(e, p + "norm.shift.weight");
    X.b1 = W(e, p + "pwconv1.bias");
    X.b2 = W(e, p + "pwconv2.bias");
    X.gamma = W(e, p + "gamma");
    LVX_TRY(make_gemm_w(e, &X.pw1, W(e, p + "pwconv1.weight"), I, D, D));
    LVX_TRY(make_gemm_w(e, &X.pw2, W(e, p + "pwconv2.weight"), D, I, I));
  }
  e->fln_w = W(e, "backbone.final_layer_norm.weight");
  e->fln_b = W(e, "backbone.final_layer_norm.bias");
  e->head_b = W(e, "head.out.bias");
  e->window = W(e, "head.istft.window");
  const int NF = c.n_fft, bins = NF / 2 + 1;
  // windowed inverse real DFT basis (spectral_ops.py:56-57): frame[n] = w[n]/N * sum_k c_k (Re_k cos - Im_k sin),
  // c_0 = c_{N/2} = 1 else 2; irfft ignores Im of the DC and Nyquist bins.
  std::vector<float> hw(NF);
  LVX_CUDA(cudaMemcpy(hw.data(), e->window, NF * sizeof(float), cudaMemcpyDeviceToHost));
  if (e->adt() == F32) {
    LVX_TRY(make_gemm_w(e, &e->head, W(e, "head.out.weight"), NF + 2, D, D));
    std::vector<float> basis((size_t)NF * e->spec_ld, 0.f);
    for (int n = 0; n < NF; ++n)
      for (int k = 0; k < bins; ++k) {
        const double ang = 2.0 * M_PI * (double)(((long long)k * n) % NF) / NF;
        const double ck = (k == 0 || k == NF / 2) ? 1.0 : 2.0;
        basis[(size_t)n * e->spec_ld + k] = (float)(hw[n] * ck * cos(ang) / NF);
        basis[(size_t)n * e->spec_ld + bins + k] = (k == 0 || k == NF / 2) ? 0.f : (float)(-hw[n] * 2.0 * sin(ang) / NF);
      }
    float* db = nullptr;
    LVX_TRY(dev_alloc(e, &db, basis.size()));
    LVX_CUDA(cudaMemcpy(db, basis.data(), basis.size() * sizeof(float), cudaMemcpyHostToDevice));
    LVX_TRY(make_gemm_w(e, &e->idft, db, NF, 2 * bins, e->spec_ld));
  } else {
    // bf16 mode keeps the head Linear and the iDFT at ~fp32 accuracy with three bf16 products laid side by
    // side along K:  a.w ~= a_hi.w_hi + a_lo.w_hi + a_hi.w_lo  ->  A' = [hi | lo | hi], W' = [hi | hi | lo].
    auto split3 = [&](const std::vector<float>& src, int N, int K, int seg, GemmW* g) -> int {
      std::vector<float> w3((size_t)N * 3 * seg, 0.f);
      for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) {
          const float v = src[(size_t)n * K + k];
          const float hi = __bfloat162float(__float2bfloat16_rn(v));
          const float lo = v - hi;
          w3[(size_t)n * 3 * seg + k] = hi;
          w3[(size_t)n * 3 * seg + seg + k] = hi;
          w3[(size_t)n * 3 * seg + 2 * seg + k] = lo;
        }
      float* d = nullptr;
      LVX_TRY(dev_alloc(e, &d, w3.size()));
      LVX_CUDA(cudaMemcpy(d, w3.data(), w3.size() * sizeof(float), cudaMemcpyHostToDevice));
      return make_gemm_w(e, g, d, N, 3 * seg, 3 * seg);
    };
    std::vector<float> hwt((size_t)(NF + 2) * D);
    LVX_CUDA(cudaMemcpy(hwt.data(), W(e, "head.out.weight"), hwt.size() * sizeof(float), cudaMemcpyDeviceToHost));
    LVX_TRY(split3(hwt, NF + 2, D, D, &e->head));
    std::vector<float> basis((size_t)NF * 2 * bins, 0.f);
    for (int n = 0; n < NF; ++n)
      for (int k = 0; k < bins; ++k) {
        const double ang = 2.0 * M_PI * (double)(((long long)k * n) % NF) / NF;
        const double ck = (k == 0 || k == NF / 2) ? 1.0 : 2.0;
        basis[(size_t)n * 2 * bins + k] = (float)(hw[n] * ck * cos(ang) / NF);
        basis[(size_t)n * 2 * bins + bins + k] = (k == 0 || k == NF / 2) ? 0.f : (float)(-hw[n] * 2.0 * sin(ang) / NF);
      }
    LVX_TRY(split3(basis, NF, 2 * bins, e->spec_ld, &e->idft));
  }
  LVX_CUDA(cudaDeviceSynchronize());
  e->finalized = true;
  return LVX_OK;
}

// ------------------------------------------------------------------------------------------------ GEMM dispatch
// C = epilogue(A . W^T).  fp32 mode: FMA-pipe kernel.  bf16 mode: tcgen05 kernel (tc_gemm.cuh) when the
// problem is a plain (optionally multi-tap) GEMM against a registered weight.
static int run_gemm(lvx_engine* e, GemmParams p, const GemmW& w, DT ta, DT tc, cudaStream_t st) {
  if (!p.a_cap) p.a_cap = p.row_chunk || p.taps > 1 ? e->R_max : e->Bp;
  p.N = w.N;
  p.K = w.K;
  p.ldw = w.ld;
  const double el = (double)dt_size(e->adt());
  const int a_cols = p.taps > 1 ? p.tap_K : p.K;
  const bool swap = e->adt() == B16 && p.taps == 1 && p.M <= 256;
  // algorithmic FLOPs: operands laid side by side along K (bf16x3 of the head / iDFT, hi | lo of exact mode) count once
  ProfScope prof(e, p.tag ? p.tag : (e->adt() == F32 ? "gemm_simt" : (swap ? "tc_gemm_swap" : "tc_gemm")), st,
                 2.0 * p.M * (double)w.N * w.K / std::max(1, p.kdup),
                 (double)p.M * a_cols * el + (double)w.N * w.K * el + (double)p.M * w.N * (double)dt_size(tc) +
                     (p.residual ? (double)p.M * w.N * 4.0 : 0.0));
  if (e->adt() == F32) {
    p.W = w.f32;
    cudaError_t err = launch_gemm_simt<float, float, float>(p, st);
    e->launches++;
    LVX_CHECK(err == cudaSuccess, LVX_ERR_CUDA, std::string("gemm launch: ") + cudaGetErrorString(err));
    return LVX_OK;
  }
  LVX_CHECK(ta == B16, LVX_ERR_INVALID, "bf16 mode GEMM needs a bf16 A operand");
  p.W = w.b16;
  int s = tc_gemm(&e->tcw, p, w.tma, tc == B16, st);
  e->launches++;
  return s;
}

static int run_gemm_batched(lvx_engine* e, GemmParams p, DT ta, DT tw, DT tc, cudaStream_t st) {
  ProfScope prof(e, "gemm_simt_batched", st);
  cudaError_t err;
  if (ta == F32)
    err = launch_gemm_simt<float, float, float>(p, st);
  else if (tc == B16)
    err = launch_gemm_simt<bf16, bf16, bf16>(p, st);
  else
    err = launch_gemm_simt<bf16, bf16, float>(p, st);
  (void)tw;
  e->launches++;
  LVX_CHECK(err == cudaSuccess, LVX_ERR_CUDA, std::string("gemm launch: ") + cudaGetErrorString(err));
  return LVX_OK;
}

template <int C>
static int run_layernorm(lvx_engine* e, const float* x, int rows, const float* w, const float* b, float eps,
                         const int* row_chunk, void* out, cudaStream_t st, bool pdl = false) {
  if (rows <= 0) return LVX_OK;
  PROF(e, "layernorm", st);
  cudaError_t err;
  if (e->adt() == F32)
    err = launch_pdl(layernorm_kernel<float, C>, dim3(ceil_div(rows, 8)), dim3(256), 0, st, pdl, x, rows, w, b, eps, row_chunk, (float*)out);
  else
    err = launch_pdl(layernorm_kernel<bf16, C>, dim3(ceil_div(rows, 8)), dim3(256), 0, st, pdl, x, rows, w, b, eps, row_chunk, (bf16*)out);
  LVX_CHECK(err == cudaSuccess, LVX_ERR_CUDA, std::string("layernorm launch: ") + cudaGetErrorString(err));
  LAUNCHED(e);
  return LVX_OK;
}

// ------------------------------------------------------------------------------------------------ sessions
static int check_engine(lvx_engine* e) {
  LVX_CHECK(e, LVX_ERR_INVALID, "engine is NULL");
  LVX_CHECK(e->finalized, LVX_ERR_STATE, "lvx_finalize_weights has not been called");
  LVX_CUDA(cudaSetDevice(e->device));
  return LVX_OK;
}

static int check_slots(lvx_engine* e, const int32_t* h_slots, int n, bool must_be_open) {
  LVX_CHECK(h_slots && n > 0, LVX_ERR_INVALID, "no slots given");
  LVX_CHECK(n <= e->cfg.max_batch, LVX_ERR_CAPACITY, "more sessions than max_batch in one call");
  ++e->stamp_ctr;
  for (int i = 0; i < n; ++i) {
    const int s = h_slots[i];
    LVX_CHECK(s >= 0 && s < e->cfg.max_sessions, LVX_ERR_INVALID, "slot out of range");
    LVX_CHECK(e->stamp[s] != e->stamp_ctr, LVX_ERR_INVALID, "duplicate slot in one call");
    e->stamp[s] = e->stamp_ctr;
    if (must_be_open) LVX_CHECK(e->h_open[s], LVX_ERR_STATE, "slot is not open");
  }
  return LVX_OK;
}

static int upload_slots(lvx_engine* e, const int32_t* h_slots, int n, cudaStream_t st) {
  LVX_CUDA(cudaMemcpyAsync(e->d_slots, h_slots, n * sizeof(int), cudaMemcpyHostToDevice, st));
  return LVX_OK;
}

static void release_pages(lvx_engine* e, int slot) {
  for (int p : e->h_pages[slot]) e->free_pages.push_back(p);
  e->h_pages[slot].clear();
}

// makes sure every slot owns pages for `tokens[i]` KV tokens; patches the device page table
static int ensure_pages(lvx_engine* e, const int32_t* h_slots, int n, const std::vector<int>& tokens, int* d_upd, cudaStream_t st) {
  std::vector<int> upd;
  for (int i = 0; i < n; ++i) {
    const int s = h_slots[i];
    const int need = ceil_div(tokens[i], e->cfg.kv_page_tokens);
    while ((int)e->h_pages[s].size() < need) {
      LVX_CHECK(!e->free_pages.empty(), LVX_ERR_CAPACITY, "KV page pool exhausted");
      const int pg = e->free_pages.back();
      e->free_pages.pop_back();
      upd.push_back(s);
      upd.push_back((int)e->h_pages[s].size());
      upd.push_back(pg);
      e->h_pages[s].push_back(pg);
    }
  }
  if (!upd.empty()) {
    const int m = (int)upd.size() / 3;
    LVX_CUDA(cudaMemcpyAsync(d_upd, upd.data(), upd.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    patch_pages_kernel<<<ceil_div(m, 256), 256, 0, st>>>(d_upd, m, e->st);
    LAUNCHED(e);
  }
  return LVX_OK;
}

extern "C" int lvx_session_open(lvx_engine* e, const int32_t* h_slots, int n, void* stream) {
  LVX_TRY(check_engine(e));
  LVX_TRY(check_slots(e, h_slots, n, false));
  cudaStream_t st = (cudaStream_t)stream;
  for (int i = 0; i < n; ++i) {
    const int s = h_slots[i];
    release_pages(e, s);
    e->h_len[s] = 0;
    e->h_text_len[s] = 0;
    e->h_open[s] = 1;
  }
  LVX_TRY(upload_slots(e, h_slots, n, st));
  reset_sessions_kernel<<<ceil_div(n, 256), 256, 0, st>>>(e->d_slots, n, e->st);
  LAUNCHED(e);
  return LVX_OK;
}

extern "C" int lvx_session_close(lvx_engine* e, const int32_t* h_slots, int n, void* stream) {
  LVX_TRY(check_engine(e));
  LVX_TRY(check_slots(e, h_slots, n, false));
  (void)stream;
  for (int i = 0; i < n; ++i) {
    release_pages(e, h_slots[i]);
    e->h_open[h_slots[i]] = 0;
  }
  return LVX_OK;
}

extern "C" int lvx_feed_text(lvx_engine* e, const int32_t* h_slots, const int32_t* h_offsets, const int32_t* h_ids, int n,
                             void* stream) {
  LVX_TRY(check_engine(e));
  LVX_TRY(check_slots(e, h_slots, n, true));
  LVX_CHECK(h_offsets && h_ids, LVX_ERR_INVALID, "NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  LVX_CHECK(h_offsets[0] == 0, LVX_ERR_INVALID, "offsets must start at 0");
  for (int i = 0; i < n; ++i) {
    const int cnt = h_offsets[i + 1] - h_offsets[i];
    LVX_CHECK(cnt >= 0, LVX_ERR_INVALID, "offsets must be non-decreasing");
    LVX_CHECK(e->h_text_len[h_slots[i]] + cnt <= e->cfg.max_context, LVX_ERR_CAPACITY, "text longer than max_context");
  }
  const int total = h_offsets[n];
  for (int i = 0; i < total; ++i)
    LVX_CHECK(h_ids[i] >= 0 && h_ids[i] < e->cfg.text_vocab, LVX_ERR_INVALID, "text id out of range");
  if (total == 0) return LVX_OK;
  LVX_TRY(upload_slots(e, h_slots, n, st));
  LVX_CUDA(cudaMemcpyAsync(e->d_aux, h_offsets, (n + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  LVX_CUDA(cudaMemcpyAsync(e->d_ids, h_ids, total * sizeof(int), cudaMemcpyHostToDevice, st));
  scatter_text_kernel<<<n, 128, 0, st>>>(e->d_slots, e->d_aux, e->d_ids, e->st);
  LAUNCHED(e);
  for (int i = 0; i < n; ++i) e->h_text_len[h_slots[i]] += h_offsets[i + 1] - h_offsets[i];
  return LVX_OK;
}

// Text front-end on the device (text_kernels.cuh): raw UTF-8 sentences -> [clean_text ->] ByT5 ids appended to the sessions' text.
// Synchronous on `stream` (the id counts come back to the host, which tracks every session's text length).
extern "C" int lvx_feed_utf8(lvx_engine* e, const int32_t* h_slots, const int32_t* h_offsets, const uint8_t* h_bytes, int n, int clean,
                             int32_t* h_counts, void* stream) {
  LVX_TRY(check_engine(e));
  LVX_TRY(check_slots(e, h_slots, n, true));
  LVX_CHECK(h_offsets && h_bytes && h_counts, LVX_ERR_INVALID, "NULL argument");
  LVX_CHECK(h_offsets[0] == 0, LVX_ERR_INVALID, "offsets must start at 0");
  const int S = e->cfg.max_batch, ctx = e->cfg.max_context;   // a call holds at most max_batch sentences (check_slots)
  for (int i = 0; i < n; ++i) {
    const int cnt = h_offsets[i + 1] - h_offsets[i];
    LVX_CHECK(cnt >= 0, LVX_ERR_INVALID, "offsets must be non-decreasing");
    LVX_CHECK(cnt <= ctx, LVX_ERR_CAPACITY, "sentence longer than max_context bytes");
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int cap = 12 * ctx + 64;   // the longest rewrite (one backslash -> " backslash ") grows a sentence 11x
  if (!e->tx_bytes) {
    LVX_TRY(dev_alloc_bytes(e, (void**)&e->tx_bytes, (size_t)S * ctx));
    LVX_TRY(dev_alloc_bytes(e, (void**)&e->tx_scratch, (size_t)2 * S * cap));
    LVX_TRY(dev_alloc_bytes(e, (void**)&e->tx_status, (size_t)S * sizeof(int)));
  }
  const int total = h_offsets[n];
  LVX_TRY(upload_slots(e, h_slots, n, st));
  LVX_CUDA(cudaMemcpyAsync(e->d_aux, h_offsets, (n + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  if (total) LVX_CUDA(cudaMemcpyAsync(e->tx_bytes, h_bytes, total, cudaMemcpyHostToDevice, st));
  text_frontend_kernel<<<ceil_div(n, TX_WARPS), 32 * TX_WARPS, 0, st>>>(e->tx_bytes, e->d_aux, e->d_slots, n, clean, e->tx_scratch, cap, e->st, e->tx_status);
  LAUNCHED(e);
  LVX_CUDA(cudaMemcpyAsync(h_counts, e->tx_status, n * sizeof(int), cudaMemcpyDeviceToHost, st));
  LVX_CUDA(cudaStreamSynchronize(st));
  int worst = 0;   // a sentence that does not fit leaves its session untouched; the others of the batch are appended
  for (int i = 0; i < n; ++i) {
    if (h_counts[i] >= 0) e->h_text_len[h_slots[i]] += h_counts[i];
    else worst = std::min(worst, h_counts[i]);
  }
  LVX_CHECK(worst != -1, LVX_ERR_CAPACITY, "clean_text scratch overflow");
  LVX_CHECK(worst == 0, LVX_ERR_CAPACITY, "text longer than max_context");
  return LVX_OK;
}

// ------------------------------------------------------------------------------------------------ decode
// One GPT.forward over n session rows whose residual stream x is already assembled (src/model.py:220-234).
static int gpt_body_exact(lvx_engine* e, lvx_engine::Lane& ln, int n, const int* pos_override, cudaStream_t st);
static int gpt_body(lvx_engine* e, lvx_engine::Lane& ln, int n, const int* pos_override, cudaStream_t st) {
  if (e->exact()) return gpt_body_exact(e, ln, n, pos_override, st);
  const lvx_config& c = e->cfg;
  const int C = c.n_embd;
  const DT a = e->adt();
  const int budget = ln.cta_budget ? ln.cta_budget : e->lane_cta_budget;
  const bool pdl = e->use_pdl && !e->prof_on;
  for (int l = 0; l < c.n_layer; ++l) {
    auto& L = e->layers[l];
    LVX_TRY(run_layernorm<768>(e, ln.x, n, L.ln1_w, L.ln1_b, 1e-5f, nullptr, ln.h, st, pdl));
    GemmParams p;
    p.A = ln.h; p.C = ln.qkv; p.M = n; p.lda = C; p.ldc = 3 * C; p.bias = L.attn_b; p.cta_budget = budget; p.pdl = pdl;
    if (e->prof_detail) p.tag = "tc_gemm:dec_qkv";
    LVX_TRY(run_gemm(e, p, L.attn, a, F32, st));
    dim3 grid(n, c.n_head);
    {
      PROF(e, "decode_attention", st);
      cudaError_t err;
      if (a == F32)
        err = launch_pdl(decode_attention_kernel<float, float, 96>, grid, dim3(128), 0, st, pdl, (const float*)ln.qkv, (float*)e->kv,
                         (const int*)ln.d_slots, e->st, pos_override, l, c.n_head, c.kv_page_tokens, e->pool_pages, 0, (float*)ln.y);
      else
        err = launch_pdl(decode_attention_kernel<bf16, bf16, 96>, grid, dim3(128), 0, st, pdl, (const float*)ln.qkv, (bf16*)e->kv,
                         (const int*)ln.d_slots, e->st, pos_override, l, c.n_head, c.kv_page_tokens, e->pool_pages, 0, (bf16*)ln.y);
      LVX_CHECK(err == cudaSuccess, LVX_ERR_CUDA, std::string("attention launch: ") + cudaGetErrorString(err));
      LAUNCHED(e);
    }
    GemmParams q;
    q.A = ln.y; q.C = ln.x; q.M = n; q.lda = C; q.ldc = C; q.bias = L.proj_b; q.residual = ln.x; q.ldr = C; q.cta_budget = budget; q.pdl = pdl;
    if (e->prof_detail) q.tag = "tc_gemm:dec_proj";
    LVX_TRY(run_gemm(e, q, L.proj, a, F32, st));
    LVX_TRY(run_layernorm<768>(e, ln.x, n, L.ln2_w, L.ln2_b, 1e-5f, nullptr, ln.h, st, pdl));
    GemmParams f;
    f.A = ln.h; f.C = ln.g; f.M = n; f.lda = C; f.ldc = 4 * C; f.bias = L.fc_b; f.act = ACT_GELU_TANH; f.cta_budget = budget; f.pdl = pdl;
    if (e->prof_detail) f.tag = "tc_gemm:dec_fc";
    LVX_TRY(run_gemm(e, f, L.fc, a, a, st));
    GemmParams r;
    r.A = ln.g; r.C = ln.x; r.M = n; r.lda = 4 * C; r.ldc = C; r.bias = L.proj2_b; r.residual = ln.x; r.ldr = C; r.cta_budget = budget; r.pdl = pdl;
    if (e->prof_detail) r.tag = "tc_gemm:dec_proj2";
    LVX_TRY(run_gemm(e, r, L.proj2, a, F32, st));
  }
  LVX_TRY(run_layernorm<768>(e, ln.x, n, e->lnf_w, e->lnf_b, 1e-5f, nullptr, ln.h, st, pdl));
  return LVX_OK;
}

// Exact mode, kernel-per-op form (the cluster-resident kernel is the fast form): every GEMM operand is a bf16 hi | lo
// pair row against the LN-folded bf16 weight laid out twice along K; fp32 KV cache, fp32 attention, exact tanh GELU.
static int ln_split(lvx_engine* e, const float* x, int n, bf16* out, cudaStream_t st, bool pdl) {
  PROF(e, "ln_split", st);
  cudaError_t err = launch_pdl(ln_split_kernel<768>, dim3(ceil_div(n, 8)), dim3(256), 0, st, pdl, x, n, 1e-5f, out);
  LVX_CHECK(err == cudaSuccess, LVX_ERR_CUDA, std::string("ln_split launch: ") + cudaGetErrorString(err));
  LAUNCHED(e);
  return LVX_OK;
}
static int split2(lvx_engine* e, const float* in, int n, int width, int act, bf16* out, cudaStream_t st, bool pdl) {
  PROF(e, "split2", st);
  const int blocks = std::max(1, std::min(1024, ceil_div(n * width / 4, 256)));
  cudaError_t err = launch_pdl(split2_kernel, dim3(blocks), dim3(256), 0, st, pdl, in, n, width, act, out);
  LVX_CHECK(err == cudaSuccess, LVX_ERR_CUDA, std::string("split2 launch: ") + cudaGetErrorString(err));
  LAUNCHED(e);
  return LVX_OK;
}
static int gpt_body_exact(lvx_engine* e, lvx_engine::Lane& ln, int n, const int* pos_override, cudaStream_t st) {
  const lvx_config& c = e->cfg;
  const int C = c.n_embd;
  const int budget = ln.cta_budget ? ln.cta_budget : e->lane_cta_budget;
  const bool pdl = e->use_pdl && !e->prof_on;
  for (int l = 0; l < c.n_layer; ++l) {
    auto& L = e->layers[l];
    LVX_TRY(ln_split(e, ln.x, n, (bf16*)ln.h, st, pdl));
    GemmParams p;
    p.A = ln.h; p.C = ln.qkv; p.M = n; p.lda = 2 * C; p.ldc = 3 * C; p.cta_budget = budget; p.pdl = pdl; p.kdup = 2;
    LVX_TRY(run_gemm(e, p, L.attn_x2, B16, F32, st));
    {
      PROF(e, "decode_attention", st);
      cudaError_t err = launch_pdl(decode_attention_kernel<float, float, 96>, dim3(n, c.n_head), dim3(128), 0, st, pdl, (const float*)ln.qkv,
                                   (float*)e->kv, (const int*)ln.d_slots, e->st, pos_override, l, c.n_head, c.kv_page_tokens, e->pool_pages,
                                   0, ln.yf);
      LVX_CHECK(err == cudaSuccess, LVX_ERR_CUDA, std::string("attention launch: ") + cudaGetErrorString(err));
      LAUNCHED(e);
    }
    LVX_TRY(split2(e, ln.yf, n, C, ACT_NONE, (bf16*)ln.y, st, pdl));
    GemmParams q;
    q.A = ln.y; q.C = ln.x; q.M = n; q.lda = 2 * C; q.ldc = C; q.residual = ln.x; q.ldr = C; q.cta_budget = budget; q.pdl = pdl; q.kdup = 2;
    LVX_TRY(run_gemm(e, q, L.proj_x2, B16, F32, st));
    LVX_TRY(ln_split(e, ln.x, n, (bf16*)ln.h, st, pdl));
    GemmParams f;
    f.A = ln.h; f.C = ln.gf; f.M = n; f.lda = 2 * C; f.ldc = 4 * C; f.cta_budget = budget; f.pdl = pdl; f.kdup = 2;
    LVX_TRY(run_gemm(e, f, L.fc_x2, B16, F32, st));
    LVX_TRY(split2(e, ln.gf, n, 4 * C, ACT_GELU_TANH, (bf16*)ln.g, st, pdl));
    GemmParams r;
    r.A = ln.g; r.C = ln.x; r.M = n; r.lda = 8 * C; r.ldc = C; r.residual = ln.x; r.ldr = C; r.cta_budget = budget; r.pdl = pdl; r.kdup = 2;
    LVX_TRY(run_gemm(e, r, L.proj2_x2, B16, F32, st));
  }
  LVX_TRY(ln_split(e, ln.x, n, (bf16*)ln.h, st, pdl));
  return LVX_OK;
}

static int lm_head_logits(lvx_engine* e, lvx_engine::Lane& ln, int n, float* d_logits, cudaStream_t st) {
  GemmParams p;
  const int kd = e->exact() ? 2 : 1;
  p.A = ln.h; p.C = d_logits; p.M = n; p.lda = kd * e->cfg.n_embd; p.ldc = e->cfg.vocab_size; p.cta_budget = ln.cta_budget ? ln.cta_budget : e->lane_cta_budget;
  p.pdl = e->use_pdl && !e->prof_on;
  p.kdup = kd;
  return run_gemm(e, p, e->exact() ? e->lm_head_x2 : e->lm_head, e->adt(), F32, st);
}

static SamplerArgs sampler_args(const lvx_sampling* s) {
  SamplerArgs a{};
  a.greedy = (!s || s->greedy || s->top_k == 1) ? 1 : 0;
  a.top_k = s ? s->top_k : 0;
  a.temperature = s ? s->temperature : 1.0f;
  a.seed = s ? s->seed : 0;
  a.uniform = s ? s->d_uniform : nullptr;
  a.record = 1;
  return a;
}

static int decode_one_step(lvx_engine* e, lvx_engine::Lane& ln, int n, const SamplerArgs& sa, float* d_logits, cudaStream_t st) {
  const lvx_config& c = e->cfg;
  {
    PROF(e, "assemble_input", st);
    cudaError_t err = launch_pdl(assemble_input_kernel, dim3(n), dim3(c.n_embd / 4), 0, st, e->use_pdl && !e->prof_on,
                                 (const int*)ln.d_slots, e->st, (const float*)W(e, "text_table"),
                                 (const float*)W(e, "feature_extractor.encodec.quantizer.vq.layers.0._codebook.embed"),
                                 (const float*)W(e, "transformer.wpe.weight"), c.text_dim, c.code_dim, c.pad_token_id, 0, ln.x);
    LVX_CHECK(err == cudaSuccess, LVX_ERR_CUDA, std::string("assemble launch: ") + cudaGetErrorString(err));
    LAUNCHED(e);
  }
  LVX_TRY(gpt_body(e, ln, n, nullptr, st));
  LVX_TRY(lm_head_logits(e, ln, n, d_logits, st));
  PROF(e, "sampler", st);
  {
    cudaError_t err = launch_pdl(sampler_kernel<4096>, dim3(n), dim3(256), 0, st, e->use_pdl && !e->prof_on, (const float*)d_logits,
                                 c.vocab_size, (const int*)ln.d_slots, e->st, sa);
    LVX_CHECK(err == cudaSuccess, LVX_ERR_CUDA, std::string("sampler launch: ") + cudaGetErrorString(err));
  }
  LAUNCHED(e);
  return LVX_OK;
}

// slot checks, page allocation and the lane's slot list / page-table patches on `st`
static int prepare_decode(lvx_engine* e, lvx_engine::Lane& ln, const int32_t* h_slots, int n, const std::vector<int>& need,
                          cudaStream_t st) {
  for (int i = 0; i < n; ++i) LVX_CHECK(need[i] <= e->cfg.max_context, LVX_ERR_CAPACITY, "session would exceed max_context");
  LVX_CUDA(cudaMemcpyAsync(ln.d_slots, h_slots, n * sizeof(int), cudaMemcpyHostToDevice, st));
  LVX_TRY(ensure_pages(e, h_slots, n, need, ln.d_upd, st));
  return LVX_OK;
}

// Graph of `iters` decode iterations for (n sessions, sampler) on this lane: every iteration launches the same
// kernels with the same arguments (all per-session state lives on the device), so it is captured once.
static int lane_graph(lvx_engine* e, lvx_engine::Lane& ln, int n, int iters, const SamplerArgs& sa, lvx_engine::StepGraph** out) {
  char key[128];
  snprintf(key, sizeof(key), "%d|%d|%d|%d|%.9g|%llu|%d", n, iters, sa.greedy, sa.top_k, (double)sa.temperature,
           (unsigned long long)sa.seed, ln.cta_budget);
  auto it = ln.graphs.find(key);
  if (it == ln.graphs.end()) {
    cudaGraph_t graph = nullptr;
    const int64_t l0 = e->launches;
    LVX_CUDA(cudaStreamBeginCapture(e->gstream, cudaStreamCaptureModeThreadLocal));
    int s2 = LVX_OK;
    for (int t = 0; t < iters && s2 == LVX_OK; ++t) s2 = decode_one_step(e, ln, n, sa, ln.logits, e->gstream);
    cudaError_t ce = cudaStreamEndCapture(e->gstream, &graph);
    const int64_t per = e->launches - l0;
    e->launches = l0;
    if (s2 != LVX_OK) {
      if (graph) cudaGraphDestroy(graph);
      return s2;
    }
    LVX_CHECK(ce == cudaSuccess && graph, LVX_ERR_CUDA, std::string("stream capture failed: ") + cudaGetErrorString(ce));
    lvx_engine::StepGraph sg;
    ce = cudaGraphInstantiate(&sg.exec, graph, 0);
    cudaGraphDestroy(graph);
    LVX_CHECK(ce == cudaSuccess, LVX_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce));
    sg.launches = per;
    if (ln.graphs.size() > 64) {
      for (auto& g : ln.graphs) cudaGraphExecDestroy(g.second.exec);
      ln.graphs.clear();
    }
    it = ln.graphs.emplace(key, sg).first;
  }
  *out = &it->second;
  return LVX_OK;
}

static unsigned long long* g_cd_diag_host = nullptr;
extern "C" const unsigned long long* lvx_cluster_diag(void) { return g_cd_diag_host; }

// Cluster-resident decode kernel (cluster_decode.cuh): the per-rank weight streams and the table row norms, built once.
static int cluster_init(lvx_engine* e) {
  if (e->cd_ready) return LVX_OK;
  const lvx_config& c = e->cfg;
  LVX_TRY(cluster_decode_configure(e->exact(), &e->cd_max_clusters, &e->cd_max_clusters8));
  LVX_CHECK(e->cd_max_clusters >= 1, LVX_ERR_CUDA, "cluster decode: no 16-CTA cluster fits on this device");
  if (getenv("LLMVOX_B200_TRACE"))
    fprintf(stderr, "[llmvox_b200] cluster decode: %d co-resident clusters of 16 CTAs, %d of 8 CTAs\n", e->cd_max_clusters, e->cd_max_clusters8);
  std::vector<CdLayerW> lw(c.n_layer);
  for (int l = 0; l < c.n_layer; ++l) {
    auto& L = e->layers[l];
    lw[l] = CdLayerW{L.attn.f32, L.proj.f32, L.fc.f32, L.proj2.f32, L.attn.ld, L.proj.ld, L.fc.ld, L.proj2.ld, L.ln1_w, L.ln2_w};
  }
  std::vector<CdPackDesc> descs;
  e->cd_stream_bytes = cd_build_descs<16>(lw.data(), c.n_layer, e->lm_head.f32, e->lm_head.ld, e->lnf_w, &descs);
  LVX_TRY(dev_alloc_bytes(e, (void**)&e->cd_stream, (size_t)e->cd_stream_bytes * 16));
  CdPackDesc* d_descs = nullptr;
  LVX_CUDA(cudaMalloc(&d_descs, descs.size() * sizeof(CdPackDesc)));
  LVX_CUDA(cudaMemcpy(d_descs, descs.data(), descs.size() * sizeof(CdPackDesc), cudaMemcpyHostToDevice));
  cd_pack_kernel<<<(unsigned)descs.size(), 256>>>(d_descs, e->cd_stream);
  LAUNCHED(e);
  if (e->cd_max_clusters8 > 0) {   // the same weights cut for 8-CTA clusters (another 62.9 MB)
    LVX_CUDA(cudaDeviceSynchronize());
    cudaFree(d_descs);
    descs.clear();
    e->cd_stream8_bytes = cd_build_descs<8>(lw.data(), c.n_layer, e->lm_head.f32, e->lm_head.ld, e->lnf_w, &descs);
    LVX_TRY(dev_alloc_bytes(e, (void**)&e->cd_stream8, (size_t)e->cd_stream8_bytes * 8));
    LVX_CUDA(cudaMalloc(&d_descs, descs.size() * sizeof(CdPackDesc)));
    LVX_CUDA(cudaMemcpy(d_descs, descs.data(), descs.size() * sizeof(CdPackDesc), cudaMemcpyHostToDevice));
    cd_pack_kernel<<<(unsigned)descs.size(), 256>>>(d_descs, e->cd_stream8);
    LAUNCHED(e);
  }
  LVX_TRY(dev_alloc(e, &e->cd_text_ss, (size_t)c.text_vocab));
  LVX_TRY(dev_alloc(e, &e->cd_code_ss, (size_t)c.vocab_size));
  cd_row_ss_kernel<<<ceil_div(c.text_vocab, 8), 256>>>(W(e, "text_table"), c.text_vocab, c.text_dim, e->cd_text_ss);
  LAUNCHED(e);
  cd_row_ss_kernel<<<ceil_div(c.vocab_size, 8), 256>>>(W(e, "feature_extractor.encodec.quantizer.vq.layers.0._codebook.embed"),
                                                      c.vocab_size, c.code_dim, e->cd_code_ss);
  LAUNCHED(e);
  LVX_CUDA(cudaDeviceSynchronize());
  cudaFree(d_descs);
  if (getenv("LLMVOX_B200_CD_DIAG")) {   // timeout records of the bounded spins, readable after the context died
    unsigned long long* h = nullptr;
    LVX_CUDA(cudaHostAlloc(&h, 256 * sizeof(unsigned long long), cudaHostAllocMapped));
    memset(h, 0, 256 * sizeof(unsigned long long));
    unsigned long long* d = nullptr;
    LVX_CUDA(cudaHostGetDevicePointer(&d, h, 0));
    LVX_CUDA(cudaMemcpyToSymbol(cd_diag, &d, sizeof(d)));
    g_cd_diag_host = h;
  }
  e->cd_ready = true;
  return LVX_OK;
}

static bool cluster_applicable(const lvx_engine* e, const SamplerArgs& sa) {
  const lvx_config& c = e->cfg;
  return e->cfg.precision != LVX_PRECISION_FP32 && (sa.greedy || !sa.uniform) && !sa.forced && !sa.out_codes && c.n_embd == CD_C &&
         c.n_head == CD_H && c.vocab_size == CD_V && c.n_layer <= CD_MAX_LAYERS && c.text_dim + c.code_dim == CD_C && !c.bias &&
         c.kv_page_tokens == 16 &&
         (long long)e->pool_pages * c.kv_page_tokens * c.n_embd < (1LL << 31);
}

// One launch = at most `cap` clusters, and never more than `cap` clusters in flight across the engine's streams, where
// `cap` is what cudaOccupancyMaxActiveClusters reports for the kernel (7 clusters of 16 CTAs on a B200): a launch that
// would exceed it waits (cudaStreamWaitEvent, no host blocking) for the oldest launches on other streams.  The bound is
// the device's co-residency and nothing else: clusters beyond it could only queue behind the resident ones, so
// capping costs no throughput, and it keeps every cluster of a wave in step with the others' weight streams (they share
// the 62.9 MB stream through L2).  History (profiles/r01d_cluster_decode.md section 7): early builds died in a bounded
// spin when MORE clusters were in flight than co-resident -- two protocol bugs (spin-waits that did not reconverge the
// warp before .sync.aligned instructions; three MMA issuers over an 8-slot ring, where consecutive uses of a slot
// belonged to different issuers and the parity wait could pass on stale data) that are fixed (cd_wait; CD_NI = 2 with the
// static_assert in CdG).  scripts/cluster_stress.py runs the kernel with the cap lifted (LLMVOX_B200_CD_CAP, a
// measurement knob that does not change results) as a regression test of those fixes.
// variant: 0 = by batch size, 16 / 8 = forced cut
static int cluster_launch(lvx_engine* e, lvx_engine::Lane& ln, const int32_t* h_slots, int n, int n_steps, const SamplerArgs& sa,
                          int variant, cudaStream_t st) {
  LVX_TRY(cluster_init(e));
  const lvx_config& c = e->cfg;
  // LLMVOX_B200_CD_CAP overrides the cap (experiments only: scripts/cluster_stress.py)
  const int cap16 = getenv("LLMVOX_B200_CD_CAP") ? std::max(1, atoi(getenv("LLMVOX_B200_CD_CAP"))) : std::max(1, e->cd_max_clusters);
  // Variant.  Up to one wave of 16-CTA clusters (7 x 16 = 112 sessions on a B200) the latency variant runs: every CTA
  // streams 1/16 of the weights.  Above that decoding switches to 8-CTA clusters: 1/8 of the weights per CTA
  // makes an iteration ~1.4x longer, but 15 clusters are co-resident, so up to 240 sessions advance in ONE wave where the
  // 16-CTA variant needs two or three.  (LLMVOX_B200_CD_NO8=1: measurement knob, results do not depend on the variant's
  // choice beyond the bf16 tolerance both meet.)
  const bool can8 = e->cd_max_clusters8 > 0;
  LVX_CHECK(variant != 8 || can8, LVX_ERR_INVALID, "no 8-CTA cluster fits on this device");
  const bool use8 = variant == 8 || (variant == 0 && can8 && n > cap16 * CD_NB && !getenv("LLMVOX_B200_CD_NO8"));
  const int cl = use8 ? 8 : 16;
  const int cap8 = getenv("LLMVOX_B200_CD_CAP8") ? std::max(1, atoi(getenv("LLMVOX_B200_CD_CAP8"))) : e->cd_max_clusters8;   // (stress only)
  const int cap = use8 ? cap8 : cap16;
  const int unit = use8 ? 1 : 2, unit_cap = use8 ? cap8 : 2 * cap16;   // in-flight accounting in 8-CTA units
  ClusterParams P;
  memset(&P, 0, sizeof(P));
  P.n_iters = n_steps; P.n_layer = c.n_layer;
  P.st = e->st;
  P.text_table = W(e, "text_table");
  P.codebook = W(e, "feature_extractor.encodec.quantizer.vq.layers.0._codebook.embed");
  P.wpe = W(e, "transformer.wpe.weight");
  P.text_ss = e->cd_text_ss; P.code_ss = e->cd_code_ss;
  P.text_dim = c.text_dim; P.code_dim = c.code_dim; P.pad_id = c.pad_token_id;
  P.wstream = use8 ? e->cd_stream8 : e->cd_stream;
  P.stream_bytes = use8 ? e->cd_stream8_bytes : e->cd_stream_bytes;
  P.top_k = sa.top_k; P.temperature = sa.temperature; P.seed = sa.seed;
  P.kv = e->kv; P.pool_pages = e->pool_pages;
  P.page_shift = 0;
  while ((1 << P.page_shift) < c.kv_page_tokens) P.page_shift += 1;
  P.trace = getenv("LLMVOX_B200_TRACE") ? ln.d_trace : nullptr;
  // waves of at most `cap` clusters, balanced (128 sessions: 2 x 64, not 112 + 16)
  const int waves = ceil_div(n, cap * CD_NB);
  const int per_wave = ceil_div(n, waves);
  for (int pos = 0; pos < n; pos += per_wave) {
    const int cnt = std::min(per_wave, n - pos);
    // 8-CTA clusters: spread the wave over all co-resident clusters, 8 sessions or more each (a CTA's eight attention
    // warps take a second pass only for sessions 9..16 of its cluster)
    if (use8) P.per_cluster = ceil_div(cnt, std::min(cap, ceil_div(cnt, 8)));
    else P.per_cluster = e->cd_spread ? ceil_div(cnt, std::min(cap, cnt)) : CD_NB;
    const int clusters = ceil_div(cnt, P.per_cluster) * unit;
    P.n = cnt;
    P.slots = ln.d_slots + pos;
    P.logits = ln.logits + (size_t)pos * c.vocab_size;
    // retire finished launches, then wait for the oldest ones on other streams until this one fits
    while (!e->cd_inflight.empty() && cudaEventQuery(e->cd_inflight.front().ev) == cudaSuccess) {
      e->ev_pool.push_back(e->cd_inflight.front().ev);
      e->cd_inflight.pop_front();
    }
    cudaGetLastError();   // cudaErrorNotReady of the query above is not an error
    int inflight = 0;
    for (auto& f : e->cd_inflight)
      if (f.st != st) inflight += f.clusters;   // launches on this stream are ordered before this one anyway
    for (auto& f : e->cd_inflight) {
      if (inflight + clusters <= unit_cap) break;
      if (f.st == st) continue;
      LVX_CUDA(cudaStreamWaitEvent(st, f.ev, 0));
      inflight -= f.clusters;
    }
    // algorithmic bytes of this launch: the bf16 GEMM weights once per iteration + KV read and append
    // (bf16 cache; fp32 in exact mode: "fp32 variant doubles" the KV term)
    double bytes = (double)n_steps * 2.0 * ((double)c.n_layer * 12.0 * CD_C * CD_C + (double)CD_C * CD_V);
    const double kv_el = (double)dt_size(e->kvdt());
    for (int i = 0; i < cnt; ++i) {
      const double t0 = e->h_len[h_slots[pos + i]];
      bytes += (double)c.n_layer * 2.0 * CD_C * kv_el * ((double)n_steps * (t0 + 1.0) + 0.5 * n_steps * (n_steps - 1.0));
    }
    {
      ProfScope prof_scope(e, "cluster_decode", st, 0.0, bytes);
      LVX_TRY(cluster_decode_launch(e->exact(), !sa.greedy, cl, P, st));
    }
    e->launches += 1;
    lvx_engine::CdInflight rec{e->get_event(), st, clusters};
    LVX_CUDA(cudaEventRecord(rec.ev, st));
    e->cd_inflight.push_back(rec);
  }
  return LVX_OK;
}

extern "C" int lvx_decode_steps_ex(lvx_engine* e, int lane, const int32_t* h_slots, int n, int n_steps, const lvx_sampling* s,
                                   int path, void* stream) {
  LVX_TRY(check_engine(e));
  LVX_CHECK(path >= LVX_PATH_AUTO && path <= LVX_PATH_PER_OP_TAIL, LVX_ERR_INVALID, "bad decode path");
  const bool forced_cluster = path == LVX_PATH_CLUSTER || path == LVX_PATH_CLUSTER16 || path == LVX_PATH_CLUSTER8;
  LVX_CHECK(lane >= 0 && lane < (int)e->lanes.size(), LVX_ERR_INVALID, "decode lane out of range");
  LVX_CHECK(n_steps > 0, LVX_ERR_INVALID, "n_steps must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  lvx_engine::Lane& ln = e->lanes[lane];
  LVX_TRY(check_slots(e, h_slots, n, true));
  std::vector<int> need(n);
  for (int i = 0; i < n; ++i) need[i] = e->h_len[h_slots[i]] + n_steps;
  LVX_TRY(prepare_decode(e, ln, h_slots, n, need, st));
  SamplerArgs sa = sampler_args(s);
  LVX_CHECK(sa.greedy || sa.temperature > 0.f, LVX_ERR_INVALID, "temperature must be positive");
  LVX_CHECK(!(sa.uniform && n_steps > 1), LVX_ERR_INVALID, "d_uniform supplies one draw per session: use n_steps == 1");
  const bool want_cluster = forced_cluster || (path == LVX_PATH_AUTO && e->use_cluster);
  // the tail of a batch that runs beside a resident wave of 8-CTA clusters sizes its grids for the SMs the wave leaves
  // (split-K factors depend on the budget: it is part of the call, never of timing, so results are reproducible)
  ln.cta_budget = 0;
  if (path == LVX_PATH_PER_OP_TAIL) {
    const int free_sms = (e->tcw.num_sms > 0 ? e->tcw.num_sms : 148) - 8 * e->cd_max_clusters8;
    ln.cta_budget = getenv("LLMVOX_B200_TAIL_BUDGET") ? atoi(getenv("LLMVOX_B200_TAIL_BUDGET")) : std::max(8, (2 * free_sms) / 3);
  }
  LVX_CHECK(!forced_cluster || cluster_applicable(e, sa), LVX_ERR_INVALID,
            "the cluster-resident decode kernel does not apply to this engine / sampler");
  if (want_cluster && cluster_applicable(e, sa)) {
    LVX_TRY(cluster_launch(e, ln, h_slots, n, n_steps, sa, path == LVX_PATH_CLUSTER16 ? 16 : path == LVX_PATH_CLUSTER8 ? 8 : 0, st));
  } else if (e->use_graphs && !e->prof_on && !sa.uniform) {
    const int unroll = 10;   // iterations per graph launch for long runs
    int left = n_steps;
    while (left > 0) {
      const int iters = left >= unroll ? unroll : 1;
      lvx_engine::StepGraph* g = nullptr;
      LVX_TRY(lane_graph(e, ln, n, iters, sa, &g));
      LVX_CUDA(cudaGraphLaunch(g->exec, st));
      e->launches += g->launches;
      left -= iters;
    }
  } else {
    for (int t = 0; t < n_steps; ++t) LVX_TRY(decode_one_step(e, ln, n, sa, ln.logits, st));
  }
  for (int i = 0; i < n; ++i) e->h_len[h_slots[i]] += n_steps;
  return LVX_OK;
}

extern "C" int lvx_decode_steps_lane(lvx_engine* e, int lane, const int32_t* h_slots, int n, int n_steps, const lvx_sampling* s,
                                     void* stream) {
  return lvx_decode_steps_ex(e, lane, h_slots, n, n_steps, s, LVX_PATH_AUTO, stream);
}

extern "C" int lvx_session_progress(lvx_engine* e, const int32_t* h_slots, int n, int32_t* h_pinned_out, void* stream) {
  LVX_TRY(check_engine(e));
  LVX_TRY(check_slots(e, h_slots, n, false));
  LVX_CHECK(h_pinned_out, LVX_ERR_INVALID, "NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  LVX_TRY(upload_slots(e, h_slots, n, st));
  gather_progress_kernel<<<ceil_div(n, 256), 256, 0, st>>>(e->d_slots, n, e->st, reinterpret_cast<int2*>(e->d_aux));
  LAUNCHED(e);
  LVX_CUDA(cudaMemcpyAsync(h_pinned_out, e->d_aux, (size_t)n * 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  return LVX_OK;
}

extern "C" int lvx_cluster_capacity(lvx_engine* e, int32_t* wave16, int32_t* wave8) {
  LVX_TRY(check_engine(e));
  LVX_CHECK(wave16 && wave8, LVX_ERR_INVALID, "NULL argument");
  *wave16 = *wave8 = 0;
  SamplerArgs greedy = sampler_args(nullptr);
  if (!cluster_applicable(e, greedy)) return LVX_OK;
  LVX_TRY(cluster_init(e));
  *wave16 = e->cd_max_clusters * CD_NB;
  *wave8 = e->cd_max_clusters8 * CD_NB;
  return LVX_OK;
}

// Host-side choice of the greedy bf16 decode path: 1 = cluster-resident kernel where applicable (default), 0 = kernel-per-op
// chain (the better choice for batches above ~224 sessions: LaneRunner switches per call).
extern "C" int lvx_set_cluster_decode(lvx_engine* e, int on) {
  LVX_TRY(check_engine(e));
  e->use_cluster = on != 0;
  return LVX_OK;
}

extern "C" int lvx_decode_steps(lvx_engine* e, const int32_t* h_slots, int n, int n_steps, const lvx_sampling* s, void* stream) {
  return lvx_decode_steps_lane(e, 0, h_slots, n, n_steps, s, stream);
}

extern "C" int lvx_peek_trace(lvx_engine* e, int lane, long long* h_out, int count) {
  LVX_TRY(check_engine(e));
  LVX_CHECK(lane >= 0 && lane < (int)e->lanes.size() && h_out && count > 0 && count <= 256, LVX_ERR_INVALID, "bad argument");
  LVX_CUDA(cudaDeviceSynchronize());
  LVX_CUDA(cudaMemcpy(h_out, e->lanes[lane].d_trace, count * sizeof(long long), cudaMemcpyDeviceToHost));
  return LVX_OK;
}

// Test hook: raw copy of a lane workspace buffer (0 x fp32, 1 qkv fp32, 2 h, 3 y, 4 g; h/y/g in the activation type)
extern "C" int lvx_peek_buffer(lvx_engine* e, int lane, int which, void* h_out, int64_t bytes) {
  LVX_TRY(check_engine(e));
  LVX_CHECK(lane >= 0 && lane < (int)e->lanes.size() && h_out && bytes > 0, LVX_ERR_INVALID, "bad argument");
  auto& ln = e->lanes[lane];
  const void* src = which == 0 ? (void*)ln.x : which == 1 ? (void*)ln.qkv : which == 2 ? ln.h : which == 3 ? ln.y : which == 4 ? ln.g : nullptr;
  LVX_CHECK(src, LVX_ERR_INVALID, "bad buffer id");
  LVX_CUDA(cudaDeviceSynchronize());
  LVX_CUDA(cudaMemcpy(h_out, src, bytes, cudaMemcpyDeviceToHost));
  return LVX_OK;
}

extern "C" int lvx_peek_logits(lvx_engine* e, int lane, int n, float* d_out, void* stream) {
  LVX_TRY(check_engine(e));
  LVX_CHECK(lane >= 0 && lane < (int)e->lanes.size() && n > 0 && n <= e->cfg.max_batch && d_out, LVX_ERR_INVALID, "bad argument");
  LVX_CUDA(cudaMemcpyAsync(d_out, e->lanes[lane].logits, (size_t)n * e->cfg.vocab_size * sizeof(float), cudaMemcpyDeviceToDevice,
                           (cudaStream_t)stream));
  return LVX_OK;
}

extern "C" int lvx_decode_step_logits(lvx_engine* e, const int32_t* h_slots, int n, const lvx_sampling* s,
                                      const int32_t* d_forced_codes, float* d_logits, int32_t* d_codes, void* stream) {
  LVX_TRY(check_engine(e));
  LVX_CHECK(d_logits, LVX_ERR_INVALID, "d_logits is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  lvx_engine::Lane& ln = e->lanes[0];
  LVX_TRY(check_slots(e, h_slots, n, true));
  std::vector<int> need(n);
  for (int i = 0; i < n; ++i) need[i] = e->h_len[h_slots[i]] + 1;
  LVX_TRY(prepare_decode(e, ln, h_slots, n, need, st));
  SamplerArgs sa = sampler_args(s);
  LVX_CHECK(sa.greedy || sa.temperature > 0.f, LVX_ERR_INVALID, "temperature must be positive");
  sa.forced = d_forced_codes;
  sa.out_codes = d_codes;
  LVX_TRY(decode_one_step(e, ln, n, sa, d_logits, st));
  for (int i = 0; i < n; ++i) e->h_len[h_slots[i]] += 1;
  return LVX_OK;
}

extern "C" int lvx_decode_step_embeds(lvx_engine* e, const int32_t* h_slots, int n, const float* d_emb,
                                      const int32_t* h_positions, float* d_logits, void* stream) {
  LVX_TRY(check_engine(e));
  LVX_CHECK(d_emb && h_positions && d_logits, LVX_ERR_INVALID, "NULL argument");
  cudaStream_t st = (cudaStream_t)stream;
  lvx_engine::Lane& ln = e->lanes[0];
  LVX_TRY(check_slots(e, h_slots, n, true));
  std::vector<int> need(n);
  for (int i = 0; i < n; ++i) {
    // the reference's cache holds exactly T-1 tokens when the caller feeds T rows (src/model.py:74-79)
    LVX_CHECK(h_positions[i] == e->h_len[h_slots[i]], LVX_ERR_STATE, "position != tokens held by the session's cache");
    need[i] = h_positions[i] + 1;
  }
  LVX_TRY(prepare_decode(e, ln, h_slots, n, need, st));
  int* d_pos = ln.d_pos;
  LVX_CUDA(cudaMemcpyAsync(d_pos, h_positions, n * sizeof(int), cudaMemcpyHostToDevice, st));
  add_wpe_kernel<<<n, 192, 0, st>>>(d_emb, d_pos, W(e, "transformer.wpe.weight"), e->cfg.n_embd, ln.x);
  LAUNCHED(e);
  LVX_TRY(gpt_body(e, ln, n, d_pos, st));
  LVX_TRY(lm_head_logits(e, ln, n, d_logits, st));
  set_ctx_kernel<<<ceil_div(n, 256), 256, 0, st>>>(ln.d_slots, d_pos, n, e->st);
  LAUNCHED(e);
  for (int i = 0; i < n; ++i) e->h_len[h_slots[i]] += 1;
  return LVX_OK;
}

extern "C" int lvx_gather_codes(lvx_engine* e, const int32_t* h_slots, int n, int start, int count, int32_t* d_out, void* stream) {
  LVX_TRY(check_engine(e));
  LVX_TRY(check_slots(e, h_slots, n, false));
  LVX_CHECK(d_out && start >= 0 && count > 0 && start + count <= e->cfg.max_context, LVX_ERR_INVALID, "bad range");
  for (int i = 0; i < n; ++i)
    LVX_CHECK(start + count <= e->h_len[h_slots[i]], LVX_ERR_STATE, "range beyond the codes decoded so far");
  cudaStream_t st = (cudaStream_t)stream;
  LVX_TRY(upload_slots(e, h_slots, n, st));
  gather_codes_kernel<<<n, 128, 0, st>>>(e->d_slots, e->st, start, count, d_out);
  LAUNCHED(e);
  return LVX_OK;
}

extern "C" int lvx_gather_code_ranges(lvx_engine* e, const int32_t* h_slots, const int32_t* h_starts, const int32_t* h_counts,
                                      int n, int32_t* d_out, void* stream) {
  LVX_TRY(check_engine(e));
  LVX_TRY(check_slots(e, h_slots, n, false));
  LVX_CHECK(h_starts && h_counts && d_out, LVX_ERR_INVALID, "NULL argument");
  std::vector<int> offs(n + 1, 0);
  for (int i = 0; i < n; ++i) {
    LVX_CHECK(h_starts[i] >= 0 && h_counts[i] >= 0 && h_starts[i] + h_counts[i] <= e->h_len[h_slots[i]], LVX_ERR_STATE,
              "range beyond the codes decoded so far");
    offs[i + 1] = offs[i] + h_counts[i];
  }
  if (offs[n] == 0) return LVX_OK;
  cudaStream_t st = (cudaStream_t)stream;
  LVX_TRY(upload_slots(e, h_slots, n, st));
  LVX_CUDA(cudaMemcpyAsync(e->d_aux, offs.data(), (n + 1) * sizeof(int), cudaMemcpyHostToDevice, st));
  LVX_CUDA(cudaMemcpyAsync(e->d_ids, h_starts, n * sizeof(int), cudaMemcpyHostToDevice, st));
  gather_code_ranges_kernel<<<n, 128, 0, st>>>(e->d_slots, e->d_ids, e->d_aux, e->st, d_out);
  LAUNCHED(e);
  return LVX_OK;
}

extern "C" int lvx_session_length(lvx_engine* e, int slot, int32_t* out_len) {
  LVX_CHECK(e && out_len, LVX_ERR_INVALID, "NULL argument");
  LVX_CHECK(slot >= 0 && slot < e->cfg.max_sessions, LVX_ERR_INVALID, "slot out of range");
  *out_len = e->h_len[slot];
  return LVX_OK;
}

extern "C" int lvx_session_text(lvx_engine* e, int slot, int32_t* h_out, int cap, int32_t* out_n, void* stream) {
  LVX_TRY(check_engine(e));
  LVX_CHECK(h_out && out_n && cap >= 0, LVX_ERR_INVALID, "bad argument");
  LVX_CHECK(slot >= 0 && slot < e->cfg.max_sessions, LVX_ERR_INVALID, "slot out of range");
  const int n = e->h_text_len[slot];
  *out_n = n;
  if (std::min(n, cap) > 0) {
    LVX_CUDA(cudaMemcpyAsync(h_out, e->st.text_ids + (size_t)slot * e->cfg.max_context, std::min(n, cap) * sizeof(int),
                             cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    LVX_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  }
  return LVX_OK;
}

extern "C" int lvx_codes_to_features(lvx_engine* e, const int32_t* d_codes, int n, float* d_out, void* stream) {
  LVX_TRY(check_engine(e));
  LVX_CHECK(d_codes && d_out && n > 0, LVX_ERR_INVALID, "bad argument");
  gather_rows_kernel<float><<<n, 128, 0, (cudaStream_t)stream>>>(
      d_codes, W(e, "feature_extractor.encodec.quantizer.vq.layers.0._codebook.embed"), e->cfg.code_dim, n, d_out,
      e->cfg.code_dim, nullptr);
  LAUNCHED(e);
  return LVX_OK;
}

extern "C" int lvx_text_embed(lvx_engine* e, const int32_t* d_ids, int n, float* d_out, void* stream) {
  LVX_TRY(check_engine(e));
  LVX_CHECK(d_ids && d_out && n > 0, LVX_ERR_INVALID, "bad argument");
  gather_rows_kernel<float><<<n, 64, 0, (cudaStream_t)stream>>>(d_ids, W(e, "text_table"), e->cfg.text_dim, n, d_out,
                                                                 e->cfg.text_dim, nullptr);
  LAUNCHED(e);
  return LVX_OK;
}

// ------------------------------------------------------------------------------------------------ vocoder
struct VocGroup {
  std::vector<ChunkInfo> chunks;
  int R = 0;        // padded rows
  int frames = 0;   // valid rows
  int code0 = 0;    // first code of the group in the caller's packed order
  int max_len = 0;
  int sld() const { return (max_len + 63) & ~63; }   // leading dimension of the group's score / probability matrices
};

static int groupnorm(lvx_engine* e, const VocGroup& g, const float* x, const float* w, const float* b, int swish, void* out,
                     cudaStream_t st) {
  PROF(e, "groupnorm", st);
  dim3 grid(32, (unsigned)g.chunks.size());
  groupnorm_stats_kernel<768><<<grid, 256, 0, st>>>(x, e->d_chunks, 1e-6f, e->v_stats);
  LAUNCHED(e);
  if (e->adt() == F32)
    groupnorm_apply_kernel<float, 768><<<ceil_div(g.R, 8), 256, 0, st>>>(x, g.R, e->row_chunk, e->v_stats, w, b, swish, (float*)out);
  else
    groupnorm_apply_kernel<bf16, 768><<<ceil_div(g.R, 8), 256, 0, st>>>(x, g.R, e->row_chunk, e->v_stats, w, b, swish, (bf16*)out);
  LAUNCHED(e);
  return LVX_OK;
}

static int conv_gemm(lvx_engine* e, const VocGroup& g, const void* A, int Cin, int taps, const GemmW& w, const float* bias,
                     float* out, const float* residual, cudaStream_t st) {
  GemmParams p;
  p.A = A; p.C = out; p.M = g.R; p.lda = Cin; p.ldc = e->cfg.voc_dim; p.a_rows = g.R;
  p.taps = taps; p.tap_K = Cin; p.tap_pad = taps / 2;
  p.bias = bias; p.residual = residual; p.ldr = e->cfg.voc_dim; p.row_chunk = e->row_chunk;
  if (e->prof_detail) p.tag = taps == 7 ? "tc_gemm:embed_k7" : "tc_gemm:resnet_k3";
  return run_gemm(e, p, w, e->adt(), F32, st);
}

static int resnet_block(lvx_engine* e, const VocGroup& g, const lvx_engine::Res& r, cudaStream_t st) {
  const int D = e->cfg.voc_dim;
  LVX_TRY(groupnorm(e, g, e->v_x, r.n1w, r.n1b, 1, e->v_h, st));
  LVX_TRY(conv_gemm(e, g, e->v_h, D, 3, r.c1, r.c1b, e->v_t, nullptr, st));
  LVX_TRY(groupnorm(e, g, e->v_t, r.n2w, r.n2b, 1, e->v_h, st));
  LVX_TRY(conv_gemm(e, g, e->v_h, D, 3, r.c2, r.c2b, e->v_x, e->v_x, st));
  return LVX_OK;
}

static int aux_fork(lvx_engine* e, cudaStream_t st) {
  LVX_CUDA(cudaEventRecord(e->ev_fork, st));
  for (int a = 0; a < lvx_engine::kAux; ++a) LVX_CUDA(cudaStreamWaitEvent(e->aux[a], e->ev_fork, 0));
  return LVX_OK;
}
static int aux_join(lvx_engine* e, cudaStream_t st) {
  for (int a = 0; a < lvx_engine::kAux; ++a) {
    LVX_CUDA(cudaEventRecord(e->ev_join[a], e->aux[a]));
    LVX_CUDA(cudaStreamWaitEvent(st, e->ev_join[a], 0));
  }
  return LVX_OK;
}

// stage: -1 = full pipeline; otherwise stop after that stage and copy the activation to d_stage_out
static int vocode_group(lvx_engine* e, const VocGroup& g, const int32_t* d_codes, const float* d_feats, int bw, float* d_pcm,
                        int stage, float* d_stage_out, cudaStream_t st) {
  const lvx_config& c = e->cfg;
  const int D = c.voc_dim, I = c.voc_inter;
  const DT a = e->adt();
  const int nch = (int)g.chunks.size();
  const int32_t* codes = d_codes ? d_codes + g.code0 : nullptr;
  auto dump = [&](const float* src, int ld, int width) -> int {
    unpad_rows_kernel<<<g.frames, 256, 0, st>>>(src, ld, width, e->code_rows, g.frames, d_stage_out);
    LAUNCHED(e);
    return LVX_OK;
  };
  // layout
  LVX_TRY(e->upload(e->d_chunks, g.chunks.data(), nch * sizeof(ChunkInfo), st));
  LVX_CUDA(cudaMemsetAsync(e->row_chunk, 0xFF, (size_t)g.R * sizeof(int), st));
  build_row_chunk_kernel<<<nch, 128, 0, st>>>(e->d_chunks, nch, e->row_chunk, e->code_rows);
  LAUNCHED(e);
  // a3: codebook gather (pretrained.py:209-239), channels-last
  const float* codebook = W(e, "feature_extractor.encodec.quantizer.vq.layers.0._codebook.embed");
  if (d_feats) {
    const float* f = d_feats + (size_t)g.code0 * c.code_dim;
    if (a == F32)
      copy_feats_padded_kernel<float><<<g.R, 128, 0, st>>>(f, c.code_dim, e->row_chunk, e->d_chunks, (float*)e->v_feats);
    else
      copy_feats_padded_kernel<bf16><<<g.R, 128, 0, st>>>(f, c.code_dim, e->row_chunk, e->d_chunks, (bf16*)e->v_feats);
  } else if (a == F32)
    gather_codebook_padded_kernel<float><<<g.R, 128, 0, st>>>(codes, codebook, c.code_dim, c.n_codes, e->row_chunk, e->d_chunks,
                                                               (float*)e->v_feats);
  else
    gather_codebook_padded_kernel<bf16><<<g.R, 128, 0, st>>>(codes, codebook, c.code_dim, c.n_codes, e->row_chunk, e->d_chunks,
                                                              (bf16*)e->v_feats);
  LAUNCHED(e);
  // backbone.embed: Conv1d 512 -> 768, k7 (models.py:224)
  LVX_TRY(conv_gemm(e, g, e->v_feats, c.code_dim, 7, e->embed, e->embed_b, e->v_x, nullptr, st));
  if (stage == 0) return dump(e->v_x, D, D);
  // pos_net (models.py:203-216)
  LVX_TRY(resnet_block(e, g, e->res[0], st));
  if (stage == 1) return dump(e->v_x, D, D);
  LVX_TRY(resnet_block(e, g, e->res[1], st));
  {
    // AttnBlock (models.py:107-127)
    LVX_TRY(groupnorm(e, g, e->v_x, e->at_nw, e->at_nb, 0, e->v_h, st));
    GemmParams p;
    p.A = e->v_h; p.C = e->v_big; p.M = g.R; p.lda = D; p.ldc = 3 * D; p.bias = e->at_qkv_b; p.row_chunk = e->row_chunk;
    if (e->prof_detail) p.tag = "tc_gemm:attn_qkv";
    LVX_TRY(run_gemm(e, p, e->at_qkv, a, a, st));
    // Scores S (fp32) and probabilities P (GEMM operand type) of the whole launch group live in [padded rows][sld]
    // matrices (row = the frame's padded row, sld = longest chunk rounded up to 64), so every chunk's Q K^T and P V are
    // windows of shared matrices.  Chunks longer than 256 frames (bf16 mode): ONE grouped tcgen05 launch for all Q K^T and
    // one for all P V (TcGroup windows over single tensor maps of v_big / P / V^T).  Shorter chunks: one ragged batch on
    // the FMA-pipe kernel (a 128 x 128 tensor-core tile per tiny chunk would cost more than it saves).
    const int sld = g.sld();
    std::vector<int> small, large;
    for (int i = 0; i < nch; ++i) ((a == B16 && g.chunks[i].len > 256) ? large : small).push_back(i);
    const int ns = (int)small.size(), nl = (int)large.size();
    std::vector<GemmProblem> ps(std::max(ns, 1)), pv(std::max(ns, 1));
    int small_max = 0;
    for (int j = 0; j < ns; ++j) {
      const ChunkInfo& ci = g.chunks[small[j]];
      const int L = ci.len;
      ps[j] = GemmProblem{(long long)ci.row0 * 3 * D, (long long)ci.row0 * 3 * D + D, (long long)ci.row0 * sld, L, L, D, 0, sld};
      pv[j] = GemmProblem{(long long)ci.row0 * sld, (long long)ci.row0 * 3 * D + 2 * D, (long long)ci.row0 * D, L, D, L, sld, 0};
      small_max = std::max(small_max, L);
    }
    const float att_scale = 1.0f / sqrtf((float)D);
    if (ns) {
      LVX_TRY(e->upload(e->d_prob_s, ps.data(), ns * sizeof(GemmProblem), st));
      LVX_TRY(e->upload(e->d_prob_pv, pv.data(), ns * sizeof(GemmProblem), st));
      GemmParams sp;
      sp.A = e->v_big; sp.W = e->v_big; sp.C = e->v_S; sp.batch = e->d_prob_s; sp.n_batch = ns; sp.lda = 3 * D; sp.ldw = 3 * D;
      sp.max_M = small_max; sp.max_N = small_max; sp.alpha = att_scale;
      LVX_TRY(run_gemm_batched(e, sp, a, a, F32, st));
    }
    int large_max = 0;
    if (nl) {
      // V^T for the whole group: the v third of the fused q|k|v output, transposed (Vt[c, r] = V[r, c]; padding rows = 0)
      {
        PROF(e, "transpose_v", st);
        transpose_v_kernel<<<dim3(ceil_div(g.R, 64), D / 64), 256, 0, st>>>((const bf16*)e->v_big, 3 * D, 2 * D, g.R, e->row_chunk, e->v_vt,
                                                                             e->R_max);
        LAUNCHED(e);
      }
      std::vector<TcGroup> gs(nl), gp(nl);
      double fl = 0;
      for (int j = 0; j < nl; ++j) {
        const ChunkInfo& ci = g.chunks[large[j]];
        const int L = ci.len;
        gs[j] = TcGroup{ci.row0, ci.row0, 0, D, L, L, ceil_div(D, TC_BK), sld, (long long)ci.row0 * sld};
        gp[j] = TcGroup{ci.row0, 0, 0, ci.row0, L, D, ceil_div(L, TC_BK), D, (long long)ci.row0 * D};
        large_max = std::max(large_max, L);
        fl += 2.0 * L * (double)L * D;
      }
      LVX_TRY(e->upload(e->d_grp_s, gs.data(), nl * sizeof(TcGroup), st));
      LVX_TRY(e->upload(e->d_grp_pv, gp.data(), nl * sizeof(TcGroup), st));
      CUtensorMap qkm;
      LVX_TRY(tc_act_map(&e->tcw, e->v_big, e->R_max, 3 * D, 3 * D, TC_BM, &qkm));
      GemmParams sp;
      sp.C = e->v_S; sp.alpha = att_scale;
      {
        ProfScope prof(e, "tc_gemm:attn_qk", st, fl, 0);
        LVX_TRY(tc_gemm_grouped(qkm, qkm, sp, e->d_grp_s, nl, large_max, large_max, ceil_div(D, TC_BK), false, st));
        e->launches++;
      }
    }
    if (a == F32)
      attn_softmax_kernel<float><<<ceil_div(g.R, 8), 256, 0, st>>>(e->v_S, e->v_S, e->d_chunks, e->row_chunk, g.R, sld);
    else
      attn_softmax_kernel<bf16><<<ceil_div(g.R, 8), 256, 0, st>>>(e->v_S, (bf16*)e->v_P, e->d_chunks, e->row_chunk, g.R, sld);
    LAUNCHED(e);
    if (ns) {
      GemmParams o;
      o.A = (a == F32) ? (void*)e->v_S : e->v_P; o.W = e->v_big; o.C = e->v_h; o.batch = e->d_prob_pv; o.n_batch = ns;
      o.ldw = 3 * D; o.ldc = D; o.w_kn = 1; o.max_M = small_max; o.max_N = D;
      LVX_TRY(run_gemm_batched(e, o, a, a, a, st));
    }
    if (nl) {
      CUtensorMap pm, vm;
      LVX_TRY(tc_act_map(&e->tcw, e->v_P, (int)std::min<long long>(e->R_max, e->score_cap / sld), sld, sld, TC_BM, &pm));
      LVX_TRY(tc_act_map(&e->tcw, e->v_vt, D, e->R_max, e->R_max, TC_BM, &vm));
      GemmParams o;
      o.C = e->v_h;
      double fl = 0;
      for (int i : large) fl += 2.0 * g.chunks[i].len * (double)g.chunks[i].len * D;
      ProfScope prof(e, "tc_gemm:attn_pv", st, fl, 0);
      LVX_TRY(tc_gemm_grouped(pm, vm, o, e->d_grp_pv, nl, large_max, D, ceil_div(large_max, TC_BK), true, st));
      e->launches++;
    }
    GemmParams q;
    q.A = e->v_h; q.C = e->v_x; q.M = g.R; q.lda = D; q.ldc = D; q.bias = e->at_proj_b; q.residual = e->v_x; q.ldr = D;
    q.row_chunk = e->row_chunk;
    if (e->prof_detail) q.tag = "tc_gemm:attn_proj";
    LVX_TRY(run_gemm(e, q, e->at_proj, a, F32, st));
  }
  if (stage == 2) return dump(e->v_x, D, D);
  LVX_TRY(resnet_block(e, g, e->res[2], st));
  LVX_TRY(resnet_block(e, g, e->res[3], st));
  // pos_net[5] GroupNorm + backbone.norm AdaLayerNorm (models.py:213,226-228)
  {
    dim3 grid(32, (unsigned)nch);
    groupnorm_stats_kernel<768><<<grid, 256, 0, st>>>(e->v_x, e->d_chunks, 1e-6f, e->v_stats);
    LAUNCHED(e);
    groupnorm_adaln_kernel<768><<<ceil_div(g.R, 8), 256, 0, st>>>(e->v_x, g.R, e->row_chunk, e->v_stats, e->pn5_w, e->pn5_b,
                                                                   e->norm_scale + (size_t)bw * D, e->norm_shift + (size_t)bw * D,
                                                                   1e-6f, stage == 3 ? e->v_t : nullptr, e->v_x);
    LAUNCHED(e);
    if (stage == 3) return dump(e->v_t, D, D);
  }
  // 12 x ConvNeXtBlock (modules.py:43-60)
  for (int i = 0; i < c.voc_layers; ++i) {
    auto& X = e->cnx[i];
    {
    PROF(e, "dwconv_adaln", st);
    // strips of DWB_R frames brought in by the bulk-copy engine; LLMVOX_B200_DW_SIMPLE selects the one-warp-per-frame kernel
    // (the cross-check of tests/test_gpu_parity.py; read per call so a test can switch)
    const bool bulk = getenv("LLMVOX_B200_DW_SIMPLE") == nullptr;
    if (bulk && a == F32)
      dwconv_adaln_bulk_kernel<float, 768><<<ceil_div(g.R, DWB_R), 192, DWB_SMEM, st>>>(e->v_x, g.R, e->row_chunk, X.dw_w, X.dw_b,
                                                                                         X.scale + (size_t)bw * D, X.shift + (size_t)bw * D,
                                                                                         1e-6f, (float*)e->v_h);
    else if (bulk)
      dwconv_adaln_bulk_kernel<bf16, 768><<<ceil_div(g.R, DWB_R), 192, DWB_SMEM, st>>>(e->v_x, g.R, e->row_chunk, X.dw_w, X.dw_b,
                                                                                        X.scale + (size_t)bw * D, X.shift + (size_t)bw * D,
                                                                                        1e-6f, (bf16*)e->v_h);
    else if (a == F32)
      dwconv_adaln_kernel<float, 768><<<ceil_div(g.R, 8), 256, 0, st>>>(e->v_x, g.R, e->row_chunk, e->d_chunks, X.dw_w, X.dw_b,
                                                                         X.scale + (size_t)bw * D, X.shift + (size_t)bw * D, 1e-6f,
                                                                         (float*)e->v_h);
    else
      dwconv_adaln_kernel<bf16, 768><<<ceil_div(g.R, 8), 256, 0, st>>>(e->v_x, g.R, e->row_chunk, e->d_chunks, X.dw_w, X.dw_b,
                                                                        X.scale + (size_t)bw * D, X.shift + (size_t)bw * D, 1e-6f,
                                                                        (bf16*)e->v_h);
    LAUNCHED(e);
    }
    GemmParams p;
    p.A = e->v_h; p.C = e->v_big; p.M = g.R; p.lda = D; p.ldc = I; p.bias = X.b1; p.row_chunk = e->row_chunk;
    p.act = (a == B16 && !getenv("LLMVOX_B200_GELU_AS")) ? ACT_GELU_ERF_BF16 : ACT_GELU_ERF;   // bf16 destination: tanh form (tc_gemm.cuh)
    if (e->prof_detail) p.tag = "tc_gemm:pw1_gelu";
    LVX_TRY(run_gemm(e, p, X.pw1, a, a, st));
    GemmParams q;
    q.A = e->v_big; q.C = e->v_x; q.M = g.R; q.lda = I; q.ldc = D; q.bias = X.b2; q.col_scale = X.gamma; q.residual = e->v_x;
    q.ldr = D; q.row_chunk = e->row_chunk;
    if (e->prof_detail) q.tag = "tc_gemm:pw2_res";
    LVX_TRY(run_gemm(e, q, X.pw2, a, F32, st));
  }
  // final_layer_norm (models.py:234) + head.out (heads.py:53)
  const int NF = c.n_fft, bins = NF / 2 + 1;
  if (a == F32) {
    layernorm_kernel<float, 768><<<ceil_div(g.R, 8), 256, 0, st>>>(e->v_x, g.R, e->fln_w, e->fln_b, 1e-6f, e->row_chunk, (float*)e->v_h);
    LAUNCHED(e);
    if (stage == 4) return dump((float*)e->v_h, D, D);
    GemmParams p;
    p.A = e->v_h; p.C = e->v_raw; p.M = g.R; p.lda = D; p.ldc = e->raw_ld; p.bias = e->head_b; p.row_chunk = e->row_chunk;
    LVX_TRY(run_gemm(e, p, e->head, a, F32, st));
    dim3 grid(g.R, ceil_div(e->spec_ld, 256));
    head_activation_kernel<float><<<grid, 256, 0, st>>>(e->v_raw, e->raw_ld, g.R, e->row_chunk, bins, (float*)e->v_spec, e->spec_ld);
    LAUNCHED(e);
    GemmParams q;
    q.A = e->v_spec; q.C = e->v_frames; q.M = g.R; q.lda = e->spec_ld; q.ldc = NF; q.row_chunk = e->row_chunk;
    LVX_TRY(run_gemm(e, q, e->idft, a, F32, st));
  } else {
    if (stage == 4) {   // test hook: the LayerNorm output itself
      layernorm_kernel<float, 768><<<ceil_div(g.R, 8), 256, 0, st>>>(e->v_x, g.R, e->fln_w, e->fln_b, 1e-6f, e->row_chunk, e->v_t);
      LAUNCHED(e);
      return dump(e->v_t, D, D);
    }
    {
      PROF(e, "layernorm_split3", st);
      layernorm_split3_kernel<768><<<ceil_div(g.R, 8), 256, 0, st>>>(e->v_x, g.R, e->fln_w, e->fln_b, 1e-6f, e->row_chunk, (bf16*)e->v_h3);
      LAUNCHED(e);
    }
    GemmParams p;
    p.kdup = 3;
    p.A = e->v_h3; p.C = e->v_raw; p.M = g.R; p.lda = 3 * D; p.ldc = e->raw_ld; p.bias = e->head_b; p.row_chunk = e->row_chunk;
    if (e->prof_detail) p.tag = "tc_gemm:head_x3";
    LVX_TRY(run_gemm(e, p, e->head, a, F32, st));
    {
      PROF(e, "head_act_split3", st);
      LVX_CHECK(16 * e->spec_ld * sizeof(bf16) <= 48 * 1024, LVX_ERR_CAPACITY, "n_fft too large for the head operand builder's shared-memory rows");
      head_act_split3_kernel<<<ceil_div(g.R, 8), 256, 16 * e->spec_ld * sizeof(bf16), st>>>(e->v_raw, e->raw_ld, g.R, e->row_chunk, bins, (bf16*)e->v_spec,
                                                                              e->spec_ld);
      LAUNCHED(e);
    }
    GemmParams q;
    q.kdup = 3;
    q.A = e->v_spec; q.C = e->v_frames; q.M = g.R; q.lda = 3 * e->spec_ld; q.ldc = NF; q.row_chunk = e->row_chunk;
    if (e->prof_detail) q.tag = "tc_gemm:idft_x3";
    LVX_TRY(run_gemm(e, q, e->idft, a, F32, st));
  }
  if (stage == 5) return dump(e->v_frames, NF, NF);
  // overlap-add + trim + envelope (spectral_ops.py:59-73)
  PROF(e, "overlap_add", st);
  float* pcm0 = d_pcm + (size_t)g.code0 * c.hop;
  if ((reinterpret_cast<uintptr_t>(pcm0) & 15) == 0 && c.hop % 4 == 0 && NF % 4 == 0 && ((NF - c.hop) / 2) % 4 == 0) {
    dim3 grid(ceil_div(g.max_len * c.hop, 1024), nch);
    overlap_add_kernel<4><<<grid, 256, 0, st>>>(e->v_frames, NF, e->d_chunks, e->window, NF, c.hop, pcm0);
  } else {
    dim3 grid(ceil_div(g.max_len * c.hop, 256), nch);
    overlap_add_kernel<1><<<grid, 256, 0, st>>>(e->v_frames, NF, e->d_chunks, e->window, NF, c.hop, pcm0);
  }
  LAUNCHED(e);
  return LVX_OK;
}

static int plan_groups(lvx_engine* e, const int32_t* h_cu, int n_chunks, std::vector<VocGroup>* out) {
  LVX_CHECK(h_cu && n_chunks > 0, LVX_ERR_INVALID, "no chunks given");
  LVX_CHECK(h_cu[0] == 0, LVX_ERR_INVALID, "h_cu must start at 0");
  VocGroup g;
  g.R = ROW_PAD;
  g.code0 = 0;
  for (int i = 0; i < n_chunks; ++i) {
    const int L = h_cu[i + 1] - h_cu[i];
    LVX_CHECK(L > 0, LVX_ERR_INVALID, "empty chunk (the reference never decodes zero codes)");
    LVX_CHECK(L <= e->cfg.max_vocode_frames, LVX_ERR_CAPACITY, "chunk longer than max_vocode_frames");
    const int row0_next = (g.R + 7) & ~7;
    // (+128: the last 128-row tile of the grouped attention GEMMs may start at the group's last rows)
    auto fits = [&](int rows, int max_len) { return (long long)(rows + 128) * ((max_len + 63) & ~63) <= e->score_cap; };
    LVX_CHECK(fits(ROW_PAD + L + ROW_PAD + 8, L), LVX_ERR_CAPACITY, "chunk too long for the attention workspace");
    if (!g.chunks.empty() &&
        (g.frames + L > e->cfg.max_vocode_frames || !fits(row0_next + L + ROW_PAD, std::max(g.max_len, L)) ||
         (int)g.chunks.size() >= e->max_chunks || row0_next + L + ROW_PAD > e->R_max - 128)) {
      out->push_back(g);
      g = VocGroup();
      g.R = ROW_PAD;
      g.code0 = h_cu[i];
    }
    ChunkInfo ci;
    ci.row0 = (g.R + 7) & ~7;   // 16-byte aligned bf16 column offset for the transposed-V tensor maps
    ci.len = L;
    ci.out0 = h_cu[i] - g.code0;
    ci.s_off = 0;
    g.chunks.push_back(ci);
    g.R = ci.row0 + L + ROW_PAD;
    g.frames += L;
    g.max_len = std::max(g.max_len, L);
  }
  out->push_back(g);
  return LVX_OK;
}

extern "C" int lvx_vocode(lvx_engine* e, const int32_t* d_codes, const int32_t* h_cu, int n_chunks, int bandwidth_id, float* d_pcm,
                          void* stream) {
  LVX_TRY(check_engine(e));
  LVX_CHECK(d_codes && d_pcm, LVX_ERR_INVALID, "NULL argument");
  LVX_CHECK(bandwidth_id >= 0 && bandwidth_id < e->cfg.voc_ada_rows, LVX_ERR_INVALID, "bandwidth_id out of range");
  std::vector<VocGroup> groups;
  LVX_TRY(plan_groups(e, h_cu, n_chunks, &groups));
  for (const VocGroup& g : groups)
    LVX_TRY(vocode_group(e, g, d_codes, nullptr, bandwidth_id, d_pcm, -1, nullptr, (cudaStream_t)stream));
  return LVX_OK;
}

extern "C" int lvx_vocode_features(lvx_engine* e, const float* d_feats, const int32_t* h_cu, int n_chunks, int bandwidth_id,
                                   float* d_pcm, void* stream) {
  LVX_TRY(check_engine(e));
  LVX_CHECK(d_feats && d_pcm, LVX_ERR_INVALID, "NULL argument");
  LVX_CHECK(bandwidth_id >= 0 && bandwidth_id < e->cfg.voc_ada_rows, LVX_ERR_INVALID, "bandwidth_id out of range");
  std::vector<VocGroup> groups;
  LVX_TRY(plan_groups(e, h_cu, n_chunks, &groups));
  for (const VocGroup& g : groups)
    LVX_TRY(vocode_group(e, g, nullptr, d_feats, bandwidth_id, d_pcm, -1, nullptr, (cudaStream_t)stream));
  return LVX_OK;
}

extern "C" int lvx_vocode_stage(lvx_engine* e, const int32_t* d_codes, int len, int bandwidth_id, int stage, float* d_out,
                                void* stream) {
  LVX_TRY(check_engine(e));
  LVX_CHECK(d_codes && d_out, LVX_ERR_INVALID, "NULL argument");
  LVX_CHECK(stage >= 0 && stage <= 5, LVX_ERR_INVALID, "stage out of range");
  LVX_CHECK(bandwidth_id >= 0 && bandwidth_id < e->cfg.voc_ada_rows, LVX_ERR_INVALID, "bandwidth_id out of range");
  const int32_t cu[2] = {0, len};
  std::vector<VocGroup> groups;
  LVX_TRY(plan_groups(e, cu, 1, &groups));
  return vocode_group(e, groups[0], d_codes, nullptr, bandwidth_id, nullptr, stage, d_out, (cudaStream_t)stream);
}

extern "C" int lvx_test_gemm(lvx_engine* e, const float* d_A, const float* d_W, int M, int N, int K, int taps, float* d_C,
                             void* stream) {
  LVX_CHECK(e, LVX_ERR_INVALID, "engine is NULL");
  LVX_CUDA(cudaSetDevice(e->device));
  LVX_CHECK(d_A && d_W && d_C && M > 0 && N > 0 && K > 0 && taps >= 1 && K % taps == 0, LVX_ERR_INVALID, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const int tap_K = K / taps;
  LVX_CHECK(tap_K % 8 == 0, LVX_ERR_INVALID, "K / taps must be a multiple of 8");
  GemmParams p;
  p.C = d_C; p.M = M; p.lda = tap_K; p.ldc = N; p.a_rows = M; p.taps = taps; p.tap_K = tap_K; p.tap_pad = taps / 2;
  GemmW w;
  w.N = N; w.K = K; w.ld = K;
  int status = LVX_OK;
  if (e->adt() == F32) {
    p.A = d_A;
    p.a_cap = M;
    w.f32 = const_cast<float*>(d_W);
    status = run_gemm(e, p, w, F32, F32, st);
  } else {
    bf16 *a16 = nullptr, *w16 = nullptr;
    const int Mp = ceil_div(M, 128) * 128;
    LVX_CUDA(cudaMalloc(&a16, (size_t)Mp * tap_K * sizeof(bf16)));
    LVX_CUDA(cudaMalloc(&w16, (size_t)N * K * sizeof(bf16)));
    LVX_CUDA(cudaMemsetAsync(a16, 0, (size_t)Mp * tap_K * sizeof(bf16), st));
    cast_bf16_kernel<<<1024, 256, 0, st>>>(d_A, a16, (size_t)M * tap_K);
    cast_bf16_kernel<<<1024, 256, 0, st>>>(d_W, w16, (size_t)N * K);
    p.A = a16;
    p.a_cap = M;   // rows >= M must read as zero for the tap shifts: the tensor map ends at M
    w.b16 = w16;
    status = tc_make_desc(&w.tma, w16, N, K, K);
    if (status == LVX_OK) status = run_gemm(e, p, w, B16, F32, st);
    cudaStreamSynchronize(st);
    e->tcw.act_maps.clear();   // the temporary operand is about to be freed
    cudaFree(a16);
    cudaFree(w16);
  }
  return status;
}

extern "C" int lvx_profile_enable(lvx_engine* e, int on) {
  LVX_CHECK(e, LVX_ERR_INVALID, "engine is NULL");
  e->prof_on = on != 0;
  e->prof_detail = on == 2;
  return LVX_OK;
}

extern "C" int lvx_profile_report(lvx_engine* e, char* buf, int64_t buf_size) {
  LVX_CHECK(e && buf && buf_size > 2, LVX_ERR_INVALID, "bad argument");
  LVX_CUDA(cudaSetDevice(e->device));
  LVX_CUDA(cudaDeviceSynchronize());
  struct Agg { long long n = 0; double ms = 0, flops = 0, bytes = 0; };
  std::map<std::string, Agg> agg;
  for (auto& r : e->prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    Agg& a = agg[r.name];
    a.n++; a.ms += ms; a.flops += r.flops; a.bytes += r.bytes;
    e->ev_pool.push_back(r.a);
    e->ev_pool.push_back(r.b);
  }
  e->prof.clear();
  std::string out = "{";
  bool first = true;
  for (auto& kv : agg) {
    char line[512];
    snprintf(line, sizeof(line), "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}", first ? "" : ", ",
             kv.first.c_str(), kv.second.n, kv.second.ms, kv.second.flops, kv.second.bytes);
    out += line;
    first = false;
  }
  out += "}";
  LVX_CHECK((int64_t)out.size() + 1 <= buf_size, LVX_ERR_CAPACITY, "report buffer too small");
  memcpy(buf, out.c_str(), out.size() + 1);
  return LVX_OK;
}

extern "C" int64_t lvx_kernel_launches(const lvx_engine* e) { return e ? e->launches : 0; }
extern "C" int64_t lvx_device_bytes(const lvx_engine* e) { return e ? e->bytes : 0; }
