/*
 * llmvox_b200 -- C ABI of the B200-native LLMVoX speech-synthesis hot path.
 *
 * One engine per GPU.  Plain pointers and sizes only (no torch types).  Every entry point returns an
 * int status (LVX_OK == 0); lvx_last_error() gives the message of the last failure on the calling
 * thread.  An engine is externally serialised (no internal locking) and re-entrant across engines.
 * "h_" arguments are host pointers, "d_" arguments are device pointers on the engine's GPU, "stream" is
 * a cudaStream_t passed as void* (NULL = the legacy default stream).  All work is enqueued on that
 * stream; nothing in the decode / vocode entry points synchronises the host.
 *
 * The reference (sabbirhossainujjal/LLMVoX) has no FFI: its boundary is the Python object protocol of
 * inference/model_handler.py:45-63 driven by streaming_server.py:250-426.  Each entry point below names
 * the reference interface it replaces; llmvox_b200/model_handler.py is the host-side mirror that binds
 * them with ctypes.
 */
#ifndef LLMVOX_B200_H
#define LLMVOX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LVX_OK 0
#define LVX_ERR_INVALID 1   /* bad argument / unsupported configuration */
#define LVX_ERR_CUDA 2      /* a CUDA runtime call failed */
#define LVX_ERR_STATE 3     /* call order (weights not finalised, slot not open, ...) */
#define LVX_ERR_CAPACITY 4  /* KV pages, context, workspace or slot capacity exceeded */

#define LVX_PRECISION_FP32 0 /* fp32 weights, activations and KV cache; FMA-pipe GEMMs (parity mode)   */
#define LVX_PRECISION_BF16 1 /* bf16 GEMM operands + KV cache, fp32 accumulate/residual; tcgen05 GEMMs */
/* Decode with bf16 WEIGHTS (LayerNorm weights folded into the GEMM that follows, then rounded: the same 62.9 MB
 * stream as BF16) but fp32-class ACTIVATIONS on the same tcgen05 tensor cores: every activation operand is split
 * into bf16 hi + bf16 lo halves whose products are accumulated in fp32, the KV cache is fp32.  Greedy tokens are
 * then identical to the reference's fp32 loop run on those rounded weights (streaming_server.py:342-346; tested).
 * The vocoder runs as in BF16 mode. */
#define LVX_PRECISION_EXACT 2

typedef struct lvx_engine lvx_engine;

/* Architecture + capacity.  Fields mirror the reference's GPTConfig (src/model.py:135-146), the frame75
 * WavTokenizer yaml (WavTokenizer/configs/...frame75...yaml:39-65) and configs/inference_config.py. */
typedef struct lvx_config {
  int32_t n_layer, n_head, n_embd, block_size, vocab_size, bias;       /* GPT: 4, 8, 768, 8192, 4096, 0 */
  int32_t text_vocab, text_dim, code_dim, n_codes;                     /* 386, 256, 512, 4096            */
  int32_t voc_dim, voc_inter, voc_layers, voc_ada_rows, n_fft, hop;    /* 768, 2304, 12, 4, 1280, 320    */
  int32_t max_sessions;     /* session slots resident on this GPU                                      */
  int32_t max_context;      /* decode steps (= KV tokens) a slot may hold, <= block_size               */
  int32_t kv_page_tokens;   /* tokens per KV page (16)                                                 */
  int32_t kv_pages;         /* pages in the pool; 0 = max_sessions * ceil(max_context / page_tokens)   */
  int32_t max_batch;        /* sessions stepped by one lvx_decode_steps call                           */
  int32_t max_vocode_frames;/* codes one lvx_vocode launch group may carry (workspace sizing)          */
  int32_t precision;        /* LVX_PRECISION_*                                                         */
  int32_t pad_token_id;     /* 384: text id fed once a session's text is exhausted (:316-320)          */
  int32_t eoa_token_id;     /* 453                                                                     */
  int32_t decode_lanes;     /* independent decode workspaces (lvx_decode_steps_lane); 0 or 1 = one     */
} lvx_config;

/* Sampler.  greedy != 0 (or top_k == 1): argmax with lowest-index tie break = the hot loop's
 * softmax->argmax (streaming_server.py:342-346).  Otherwise the semantics of GPT.generate
 * (src/model.py:397-406): logits / temperature; keep entries >= the top_k-th largest (ties kept,
 * top_k <= 0 disables); softmax; one draw by inverse CDF in index order against a uniform from
 * Philox4x32-10(seed; slot, step) -- or d_uniform[i] for session i of the call when non-NULL. */
typedef struct lvx_sampling {
  int32_t greedy;
  int32_t top_k;
  float temperature;
  uint64_t seed;
  const float* d_uniform;
} lvx_sampling;

const char* lvx_last_error(void);
int lvx_version(void);

/* Fills *cfg with the english-tiny / frame75 architecture and default capacities. */
int lvx_config_default(lvx_config* cfg);

/* Replaces ModelHandler.__init__ (inference/model_handler.py:48-63): allocates weights, KV pool, session
 * state and workspaces on `device`. */
int lvx_engine_create(const lvx_config* cfg, int device, lvx_engine** out);
int lvx_engine_destroy(lvx_engine* e);

/* Weight upload, one tensor at a time, fp32 host data, using the reference's state-dict key names
 * (e.g. "transformer.h.0.attn.c_attn.weight", "backbone.convnext.3.pwconv1.weight",
 * "feature_extractor.encodec.quantizer.vq.layers.0._codebook.embed", "head.istft.window") plus
 * "text_table" for T5 encoder.embed_tokens (model_handler.py:105).  Replaces load_state_dict at
 * model_handler.py:163 and WavTokenizer/decoder/pretrained.py:113.  Unknown names are an error; keys of
 * the unused SEANet encoder/decoder must be filtered by the caller. */
int lvx_load_tensor(lvx_engine* e, const char* name, const float* h_data, const int64_t* shape, int ndim);
/* Checks every tensor arrived, builds derived tensors (tap-major conv weights, fused q|k|v, windowed iDFT
 * basis, bf16 copies).  Must precede any compute call. */
int lvx_finalize_weights(lvx_engine* e);

/* Session slots.  open == the per-sentence state reset of streaming_server.py:268-282 / :404-416
 * (kvcache = None, speech_gen_index = 0): context 0, no text, KV pages released. */
int lvx_session_open(lvx_engine* e, const int32_t* h_slots, int n, void* stream);
int lvx_session_close(lvx_engine* e, const int32_t* h_slots, int n, void* stream);

/* Appends text ids to sessions: ids of session i are h_ids[h_offsets[i] .. h_offsets[i+1]).  Replaces the
 * per-word tokenizer -> tensor -> .to(device) -> llm_model(ids) of streaming_server.py:306-315 (the
 * embedding gather itself happens inside the decode step). */
int lvx_feed_text(lvx_engine* e, const int32_t* h_slots, const int32_t* h_offsets, const int32_t* h_ids,
                  int n, void* stream);

/* The same from raw text, tokenised on the device: sentence i is the UTF-8 bytes h_bytes[h_offsets[i] ..
 * h_offsets[i+1]).  clean != 0 runs the reference's clean_text (streaming_server.py:106-149) first; the text
 * is then split at spaces, every word stripped and byte-tokenised (byte + 3, "[PAD]" -> 384, "EOS" -> 385)
 * with its own </s> = 1, and 385 follows the last word -- the ids the producer / generator threads of
 * streaming_server.py:184-248, 297-310 feed for a whole sentence.  h_counts[i] receives the number of ids
 * appended to session i.  Synchronous on `stream`.  LVX_ERR_CAPACITY if a sentence is longer than max_context
 * bytes or its ids do not fit the session. */
int lvx_feed_utf8(lvx_engine* e, const int32_t* h_slots, const int32_t* h_offsets, const uint8_t* h_bytes,
                  int n, int clean, int32_t* h_counts, void* stream);

/* n_steps decode steps for n sessions, batched: the loop body of streaming_server.py:323-354 (input
 * assembly a4, GPT.forward a5-a8 with a paged KV cache, pick a9) with no host sync per token.  Step t of a
 * session consumes text id t (pad_token_id beyond its text), the previous code's codebook row (zeros at
 * t = 0) and position t; the new code is appended to the session's device-side code history. */
int lvx_decode_steps(lvx_engine* e, const int32_t* h_slots, int n, int n_steps, const lvx_sampling* s,
                     void* stream);

/* Same on decode lane `lane` (0 <= lane < decode_lanes).  A decode iteration is a chain of ~35 dependent,
 * latency-bound kernels; sessions are independent, so disjoint groups of sessions can run their chains
 * CONCURRENTLY: each lane owns its own workspace and CUDA graphs, and calls on different lanes may be enqueued
 * on different streams without any ordering between them.  lvx_decode_steps == lane 0.  All other entry points
 * (open / feed / gather / vocode) share one staging area and must be issued on one stream at a time. */
int lvx_decode_steps_lane(lvx_engine* e, int lane, const int32_t* h_slots, int n, int n_steps, const lvx_sampling* s,
                          void* stream);

/* Same with the decode path chosen PER CALL (no engine state): LVX_PATH_AUTO = the engine's default (cluster-resident
 * kernel where it applies, see lvx_set_cluster_decode), LVX_PATH_CLUSTER = the cluster-resident kernel (an error when it
 * does not apply: fp32 mode, sampled decoding), LVX_PATH_PER_OP = the kernel-per-op chain.  Calls with different paths
 * may be in flight on different lanes / streams at the same time (e.g. 224 sessions on the cluster kernel and the rest
 * of a 256-session batch on the kernel-per-op lanes). */
#define LVX_PATH_AUTO 0
#define LVX_PATH_CLUSTER 1
#define LVX_PATH_PER_OP 2
/* The cluster-resident kernel exists in two cuts of the same weights: 16-CTA clusters (lowest latency, one wave holds
 * 7 x 16 sessions on a B200) and 8-CTA clusters (one wave holds 15 x 16 sessions: the 256-streams
 * operating point).  LVX_PATH_CLUSTER picks by batch size (16-CTA up to one wave of them, 8-CTA above); the two values
 * below force one cut (an error when it does not apply).  Within a cut a session's codes do not depend on how calls are
 * composed; across cuts they agree within the bf16 contract (teacher-forced logits within 2e-2), not bit for bit. */
#define LVX_PATH_CLUSTER16 3
#define LVX_PATH_CLUSTER8 4
/* The kernel-per-op chain sized for the SMs a resident wave of 8-CTA clusters leaves free: the tail of a batch slightly
 * larger than one wave (256 streams = 240 on the clusters + 16 here, at the same time). */
#define LVX_PATH_PER_OP_TAIL 5
int lvx_decode_steps_ex(lvx_engine* e, int lane, const int32_t* h_slots, int n, int n_steps, const lvx_sampling* s,
                        int path, void* stream);

/* Sessions ONE wave of the cluster-resident kernel advances together: *wave16 for 16-CTA clusters, *wave8 for 8-CTA
 * clusters (0 where the cut does not exist: fp32 precision has neither, exact precision only the first).  Host
 * schedulers use it to size batches (streaming.py: LaneRunner.plan). */
int lvx_cluster_capacity(lvx_engine* e, int32_t* wave16, int32_t* wave8);

/* Per-session progress for the host's chunk scheduler (streaming_server.py:357-422): writes (eoa_pos, ctx_len) pairs
 * -- the position of the sentence's first end-of-audio code (eoa_token_id; -1 = none yet) and the codes decoded so far
 * -- of n sessions to h_pinned_out (2 n int32, pinned host memory) with an asynchronous copy on `stream`.  The decode
 * kernels maintain eoa_pos on the device, so no code VALUE has to visit the host to detect the end of a sentence. */
int lvx_session_progress(lvx_engine* e, const int32_t* h_slots, int n, int32_t* h_pinned_out, void* stream);

/* Greedy bf16 decode path of lvx_decode_steps[_lane]: on != 0 (default) = the cluster-resident kernel (one 16-CTA
 * cluster per 16 sessions runs whole iterations; at most 7 clusters in flight per engine), 0 = the kernel-per-op chain
 * (CUDA graphs + programmatic dependent launch), which is the better choice for batches above ~224 sessions.  Both
 * compute src/model.py:201-237 with the same numerics class; fp32 mode and sampled decoding always
 * use the kernel-per-op chain.  Host-side flag: takes effect for the calls that follow. */
int lvx_set_cluster_decode(lvx_engine* e, int on);

/* Test hook: ONE step that also returns the logits (n x vocab fp32, device) and the picked codes (n,
 * device, may be NULL).  With d_forced_codes != NULL the code stored in the history (and therefore fed
 * back at the next step) is d_forced_codes[i] instead of the pick: teacher forcing. */
int lvx_decode_step_logits(lvx_engine* e, const int32_t* h_slots, int n, const lvx_sampling* s,
                           const int32_t* d_forced_codes, float* d_logits, int32_t* d_codes, void* stream);

/* Test hook: copies the logits of the LAST decode iteration run on `lane` (n x vocab fp32) to d_out. */
int lvx_peek_logits(lvx_engine* e, int lane, int n, float* d_out, void* stream);

/* Test hook: raw host copy of a decode-lane workspace buffer (which: 0 x, 1 qkv (fp32); 2 h, 3 y, 4 g (activation type)). */
int lvx_peek_buffer(lvx_engine* e, int lane, int which, void* h_out, int64_t bytes);

/* Test hook: with LLMVOX_B200_TRACE set at engine creation, CTA 0 of the fused decode kernel stamps clock64() at every
 * phase boundary of the last iteration of a launch; this copies the first `count` (<= 256) stamps to h_out. */
int lvx_peek_trace(lvx_engine* e, int lane, long long* h_out, int count);

/* Drop-in for `model(emb, kvcache)` (streaming_server.py:341 -> src/model.py:201-237): the caller supplies
 * the assembled, normalised input row of each session (n x n_embd fp32, device) and its position (the
 * reference's T-1); returns logits (n x vocab fp32, device) and appends K/V.  No code is recorded. */
int lvx_decode_step_embeds(lvx_engine* e, const int32_t* h_slots, int n, const float* d_emb,
                           const int32_t* h_positions, float* d_logits, void* stream);

/* Copies codes [start, start+count) of each session's history to d_out (n x count int32, device). */
int lvx_gather_codes(lvx_engine* e, const int32_t* h_slots, int n, int start, int count, int32_t* d_out,
                     void* stream);
/* Ragged form: codes [h_starts[i], h_starts[i] + h_counts[i]) of session i, packed back to back in d_out
 * (sum of counts int32, device) -- the chunk cut `speech_outputs[:dump_size]` of streaming_server.py:359-360
 * for many sessions at once. */
int lvx_gather_code_ranges(lvx_engine* e, const int32_t* h_slots, const int32_t* h_starts, const int32_t* h_counts,
                           int n, int32_t* d_out, void* stream);
/* Host mirror of a slot's context length (codes decoded so far). */
int lvx_session_length(lvx_engine* e, int slot, int32_t* out_len);
/* The text ids a slot holds, read back from the device (synchronous on `stream`): up to `cap` ids into h_out,
 * *out_n = the slot's text length. */
int lvx_session_text(lvx_engine* e, int slot, int32_t* h_out, int cap, int32_t* out_n, void* stream);

/* Replaces wavtokenizer.codes_to_features(codes) (pretrained.py:209-239): n codes -> n x code_dim fp32
 * rows (channels-last; the reference returns the transpose (1, 512, n)). */
int lvx_codes_to_features(lvx_engine* e, const int32_t* d_codes, int n, float* d_out, void* stream);
/* Replaces llm_model(ids) (model_handler.py:105): n ids -> n x text_dim fp32 rows. */
int lvx_text_embed(lvx_engine* e, const int32_t* d_ids, int n, float* d_out, void* stream);

/* Replaces codes_to_features + wavtokenizer.decode(features, bandwidth_id) (streaming_server.py:363-365 ->
 * pretrained.py:192-207 -> models.py:223-235, heads.py:53-66, spectral_ops.py:33-75) for a ragged batch of
 * INDEPENDENT chunks: chunk i holds codes d_codes[h_cu[i] .. h_cu[i+1]) and yields hop * len_i samples at
 * d_pcm[hop * h_cu[i]].  Conv padding, GroupNorm statistics, the pos_net attention and the iSTFT edge
 * envelope are per chunk, exactly as a separate reference call per chunk. */
int lvx_vocode(lvx_engine* e, const int32_t* d_codes, const int32_t* h_cu, int n_chunks, int bandwidth_id,
               float* d_pcm, void* stream);

/* Same, from features: replaces wavtokenizer.decode(features, bandwidth_id) alone (pretrained.py:192-207) for
 * callers that already hold codes_to_features' output.  d_feats is channels-last: row h_cu[i] + t = frame t of
 * chunk i, code_dim fp32 values (the reference tensor is (1, code_dim, L); the host mirror transposes). */
int lvx_vocode_features(lvx_engine* e, const float* d_feats, const int32_t* h_cu, int n_chunks, int bandwidth_id,
                        float* d_pcm, void* stream);

/* Test hook: runs lvx_vocode's pipeline up to `stage` for ONE chunk and copies that activation
 * (len x width fp32, channels-last) to d_out.  Stages: 0 embed conv, 1 pos_net[0], 2 pos_net[0..2] (after
 * attention), 3 pos_net output (after final GroupNorm), 4 backbone output (after final LayerNorm),
 * 5 windowed iDFT frames (len x n_fft). */
int lvx_vocode_stage(lvx_engine* e, const int32_t* d_codes, int len, int bandwidth_id, int stage,
                     float* d_out, void* stream);

/* Test hook: C (M x N fp32) = A (a_rows x tap_K fp32, row-major) . W (N x K fp32, row-major)^T through the
 * engine's own GEMM path for its precision (FMA-pipe fp32, or tcgen05 with operands rounded to bf16).  With
 * taps > 1 (K = taps * tap_K) column k = tap * tap_K + c of row m reads A[m + tap - taps / 2, c], rows outside
 * [0, M) reading as zero: the conv-as-GEMM form of the vocoder's k = 3 / k = 7 convolutions. */
int lvx_test_gemm(lvx_engine* e, const float* d_A, const float* d_W, int M, int N, int K, int taps, float* d_C,
                  void* stream);

/* Per-kernel timing for bench.py's roofline: while enabled, every launch is bracketed by a pair of CUDA events
 * on the launching stream (this serialises nothing but adds ~2 us of host work per launch, so the profiled pass
 * is run beside the timed region, never inside it).  lvx_profile_report synchronises the device and writes one
 * JSON object {"<kernel>": {"launches": n, "ms": t, "flops": f, "bytes": b}, ...} (algorithmic flops / bytes,
 * GEMMs only) into buf, then clears the records. */
int lvx_profile_enable(lvx_engine* e, int on);
int lvx_profile_report(lvx_engine* e, char* buf, int64_t buf_size);

/* Counters for bench.py's "gpu_launches": kernels launched by this engine since creation. */
int64_t lvx_kernel_launches(const lvx_engine* e);
/* Bytes of device memory the engine allocated. */
int64_t lvx_device_bytes(const lvx_engine* e);

#ifdef __cplusplus
}
#endif
#endif /* LLMVOX_B200_H */
