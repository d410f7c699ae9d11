"""Python face of the C-ABI engine (include/llmvox_b200.h).  torch is used for device memory and streams
only; every computation happens in libllmvox_b200.so.  One Engine per GPU."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import LvxConfig, LvxSampling, PRECISION_BF16, PRECISION_EXACT, PRECISION_FP32, check, i32_array
from .weights import CODEBOOK_KEY, GPTArch, VocoderArch

_UNUSED_PREFIXES = ("feature_extractor.",)   # SEANet encoder etc.: loaded by the reference, never executed


@dataclass
class Sampling:
    """Sampler of GPT.generate (src/model.py:397-406); greedy = the hot loop's argmax (streaming_server.py:342-346)."""
    greedy: bool = True
    top_k: int = 0
    temperature: float = 1.0
    seed: int = 0

    def to_c(self, d_uniform: Optional[torch.Tensor] = None) -> LvxSampling:
        return LvxSampling(int(self.greedy), int(self.top_k), float(self.temperature), int(self.seed),
                           d_uniform.data_ptr() if d_uniform is not None else None)


class Engine:
    def __init__(self, weights: Dict[str, torch.Tensor], device: int = 0, precision: str = "fp32",
                 gpt_arch: Optional[GPTArch] = None, voc_arch: Optional[VocoderArch] = None,
                 max_sessions: int = 256, max_context: int = 1024, max_batch: Optional[int] = None,
                 max_vocode_frames: int = 32768, kv_page_tokens: int = 16, kv_pages: int = 0,
                 pad_token_id: int = 384, eoa_token_id: int = 453, decode_lanes: int = 1):
        if not torch.cuda.is_available():
            raise RuntimeError("llmvox_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        ga, va = gpt_arch or GPTArch(), voc_arch or VocoderArch()
        cfg = LvxConfig()
        check(self.lib.lvx_config_default(C.byref(cfg)))
        cfg.n_layer, cfg.n_head, cfg.n_embd = ga.n_layer, ga.n_head, ga.n_embd
        cfg.block_size, cfg.vocab_size, cfg.bias = ga.block_size, ga.vocab_size, int(ga.bias)
        cfg.code_dim, cfg.n_codes = va.input_channels, va.vq_bins
        cfg.voc_dim, cfg.voc_inter, cfg.voc_layers = va.dim, va.intermediate_dim, va.num_layers
        cfg.voc_ada_rows, cfg.n_fft, cfg.hop = va.adanorm_num_embeddings, va.n_fft, va.hop_length
        cfg.max_sessions, cfg.max_context = max_sessions, max_context
        cfg.max_batch = max_batch or max_sessions
        cfg.max_vocode_frames, cfg.kv_page_tokens, cfg.kv_pages = max_vocode_frames, kv_page_tokens, kv_pages
        cfg.precision = {"fp32": PRECISION_FP32, "bf16": PRECISION_BF16, "exact": PRECISION_EXACT}[precision]
        cfg.pad_token_id, cfg.eoa_token_id = pad_token_id, eoa_token_id
        cfg.decode_lanes = decode_lanes
        self.decode_lanes = max(1, decode_lanes)
        self.cfg = cfg
        self.precision = precision
        self.device = torch.device("cuda", device)
        self._h = C.c_void_p()
        check(self.lib.lvx_engine_create(C.byref(cfg), device, C.byref(self._h)))
        try:
            self._load(weights)
        except Exception:
            self.close()
            raise

    # ------------------------------------------------------------------ lifecycle
    def _load(self, weights: Dict[str, torch.Tensor]):
        for name, t in weights.items():
            if name != CODEBOOK_KEY and name.startswith(_UNUSED_PREFIXES):
                continue
            t = t.detach().to(device="cpu", dtype=torch.float32).contiguous()
            if name == "transformer.wpe.weight" and t.shape[0] > self.cfg.max_context:
                t = t[: self.cfg.max_context].contiguous()      # only positions < max_context are ever read
            shape = (C.c_int64 * t.dim())(*t.shape)
            check(self.lib.lvx_load_tensor(self._h, name.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()))
        check(self.lib.lvx_finalize_weights(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.lvx_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def kernel_launches(self) -> int:
        return int(self.lib.lvx_kernel_launches(self._h))

    @property
    def device_bytes(self) -> int:
        return int(self.lib.lvx_device_bytes(self._h))

    def profile(self, on):
        """True / 1: per kernel class; 2: vocoder GEMMs labelled by role as well."""
        check(self.lib.lvx_profile_enable(self._h, int(on)))

    def profile_report(self) -> dict:
        import json
        buf = C.create_string_buffer(1 << 16)
        check(self.lib.lvx_profile_report(self._h, buf, len(buf)))
        return json.loads(buf.value.decode())

    def _stream(self, stream) -> C.c_void_p:
        s = stream if stream is not None else torch.cuda.current_stream(self.device)
        return C.c_void_p(s.cuda_stream)

    # ------------------------------------------------------------------ sessions
    def open(self, slots: Sequence[int], stream=None):
        check(self.lib.lvx_session_open(self._h, i32_array(slots), len(slots), self._stream(stream)))

    def release(self, slots: Sequence[int], stream=None):
        check(self.lib.lvx_session_close(self._h, i32_array(slots), len(slots), self._stream(stream)))

    def feed_text(self, slots: Sequence[int], ids: Sequence[Sequence[int]], stream=None):
        import itertools
        import numpy as np
        offs = np.zeros((len(ids) + 1,), dtype=np.int32)
        np.cumsum([len(seq) for seq in ids], out=offs[1:])
        flat = np.fromiter(itertools.chain.from_iterable(ids), dtype=np.int32, count=int(offs[-1]))
        check(self.lib.lvx_feed_text(self._h, i32_array(slots), i32_array(offs), i32_array(flat), len(slots),
                                     self._stream(stream)))

    def feed_sentences(self, slots: Sequence[int], sentences: Sequence[str], clean: bool = True, stream=None) -> List[int]:
        """Raw sentences -> [clean_text ->] ByT5 ids appended to the sessions' text, tokenised on the device
        (lvx_feed_utf8): the ids `tokenizer.sentence_ids(protocol.clean_text(s))` would give.  Returns the id count
        per sentence.  Synchronous on the stream."""
        import numpy as np
        raw = [s.encode("utf-8") for s in sentences]
        offs = np.zeros((len(raw) + 1,), dtype=np.int32)
        np.cumsum([len(b) for b in raw], out=offs[1:])
        counts = np.zeros((len(raw),), dtype=np.int32)
        check(self.lib.lvx_feed_utf8(self._h, i32_array(slots), i32_array(offs), b"".join(raw), len(slots), int(bool(clean)),
                                     counts.ctypes.data_as(C.POINTER(C.c_int32)), self._stream(stream)))
        return counts.tolist()

    def session_text(self, slot: int, stream=None) -> List[int]:
        """The text ids the slot holds, read back from the device."""
        import numpy as np
        out = np.zeros((self.cfg.max_context,), dtype=np.int32)
        n = C.c_int32()
        check(self.lib.lvx_session_text(self._h, slot, out.ctypes.data_as(C.POINTER(C.c_int32)), out.size, C.byref(n), self._stream(stream)))
        return out[:n.value].tolist()

    def session_length(self, slot: int) -> int:
        out = C.c_int32()
        check(self.lib.lvx_session_length(self._h, slot, C.byref(out)))
        return out.value

    # ------------------------------------------------------------------ decode
    def decode_steps(self, slots: Sequence[int], n_steps: int, sampling: Optional[Sampling] = None, stream=None, lane: int = 0,
                     path: int = _lib.PATH_AUTO):
        """n_steps decode iterations for `slots` on decode lane `lane`.  Calls on different lanes may be enqueued on
        different streams and run concurrently (sessions are independent).  `path`: _lib.PATH_AUTO / PATH_CLUSTER /
        PATH_PER_OP / PATH_CLUSTER16 / PATH_CLUSTER8 (lvx_decode_steps_ex)."""
        s = (sampling or Sampling()).to_c()
        check(self.lib.lvx_decode_steps_ex(self._h, lane, i32_array(slots), len(slots), n_steps, C.byref(s), int(path),
                                           self._stream(stream)))

    def session_progress(self, slots: Sequence[int], out: torch.Tensor, stream=None):
        """Asynchronously writes (eoa_pos, ctx_len) of each slot into the pinned int32 host tensor `out` (n, 2)."""
        assert out.dtype == torch.int32 and out.is_pinned() and out.numel() >= 2 * len(slots)
        check(self.lib.lvx_session_progress(self._h, i32_array(slots), len(slots), C.c_void_p(out.data_ptr()),
                                            self._stream(stream)))

    def set_cluster_decode(self, on: bool):
        """Greedy bf16 decode path for the calls that follow: the cluster-resident kernel (default) or the kernel-per-op
        chain (better above ~224 sessions per batch)."""
        check(self.lib.lvx_set_cluster_decode(self._h, int(bool(on))))

    def cluster_capacity(self) -> Tuple[int, int]:
        """Sessions one wave of the cluster-resident kernel advances together: (16-CTA clusters, 8-CTA clusters); 0 where
        the cut does not exist (lvx_cluster_capacity).  (112, 240) for bf16 on a B200."""
        a, b = C.c_int32(), C.c_int32()
        check(self.lib.lvx_cluster_capacity(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def cluster_decode_applicable(self, sampling: Optional[Sampling] = None) -> bool:
        """Host-side mirror of the engine's own test (engine.cu: cluster_applicable): would a decode_steps call with this
        sampler run on the cluster-resident kernel?"""
        c = self.cfg
        return (self.precision in ("bf16", "exact") and c.n_embd == 768 and c.n_head == 8 and c.vocab_size == 4096 and not c.bias
                and c.kv_page_tokens == 16 and c.text_dim + c.code_dim == 768)

    def decode_step_logits(self, slots: Sequence[int], forced: Optional[torch.Tensor] = None,
                           sampling: Optional[Sampling] = None, uniform: Optional[torch.Tensor] = None,
                           stream=None) -> Tuple[torch.Tensor, torch.Tensor]:
        n = len(slots)
        logits = torch.empty((n, self.cfg.vocab_size), dtype=torch.float32, device=self.device)
        codes = torch.empty((n,), dtype=torch.int32, device=self.device)
        s = (sampling or Sampling()).to_c(uniform)
        fp = C.c_void_p(forced.data_ptr()) if forced is not None else None
        if forced is not None:
            assert forced.dtype == torch.int32 and forced.is_cuda and forced.numel() == n
        check(self.lib.lvx_decode_step_logits(self._h, i32_array(slots), n, C.byref(s), fp, C.c_void_p(logits.data_ptr()),
                                              C.c_void_p(codes.data_ptr()), self._stream(stream)))
        return logits, codes

    def peek_buffer(self, which: str, rows: int, lane: int = 0):
        """Test hook: host copy of a decode-lane workspace buffer ('x', 'qkv', 'h', 'y', 'g') as float32 (rows, width)."""
        import numpy as np
        C_ = self.cfg.n_embd
        idx, width, f32 = {"x": (0, C_, True), "qkv": (1, 3 * C_, True), "h": (2, C_, False), "y": (3, C_, False),
                           "g": (4, 4 * C_, False)}[which]
        f32 = f32 or self.precision == "fp32"
        buf = np.empty((rows, width), dtype=np.float32 if f32 else np.uint16)
        check(self.lib.lvx_peek_buffer(self._h, lane, idx, C.c_void_p(buf.ctypes.data), buf.nbytes))
        if not f32:
            buf = (buf.astype(np.uint32) << 16).view(np.float32)
        return buf

    def peek_trace(self, count: int = 128, lane: int = 0):
        """Test hook: clock64 stamps of the fused decode kernel's phase boundaries (LLMVOX_B200_TRACE=1)."""
        buf = (C.c_int64 * count)()
        check(self.lib.lvx_peek_trace(self._h, lane, buf, count))
        return list(buf)

    def peek_logits(self, n: int, lane: int = 0, stream=None) -> torch.Tensor:
        """Test hook: logits of the last decode iteration run on `lane`."""
        out = torch.empty((n, self.cfg.vocab_size), dtype=torch.float32, device=self.device)
        check(self.lib.lvx_peek_logits(self._h, lane, n, C.c_void_p(out.data_ptr()), self._stream(stream)))
        return out

    def decode_step_embeds(self, slots: Sequence[int], emb: torch.Tensor, positions: Sequence[int], stream=None) -> torch.Tensor:
        n = len(slots)
        emb = emb.to(device=self.device, dtype=torch.float32).contiguous()
        assert emb.shape == (n, self.cfg.n_embd)
        logits = torch.empty((n, self.cfg.vocab_size), dtype=torch.float32, device=self.device)
        check(self.lib.lvx_decode_step_embeds(self._h, i32_array(slots), n, C.c_void_p(emb.data_ptr()), i32_array(positions),
                                              C.c_void_p(logits.data_ptr()), self._stream(stream)))
        return logits

    def gather_codes(self, slots: Sequence[int], start: int, count: int, stream=None) -> torch.Tensor:
        out = torch.empty((len(slots), count), dtype=torch.int32, device=self.device)
        check(self.lib.lvx_gather_codes(self._h, i32_array(slots), len(slots), start, count, C.c_void_p(out.data_ptr()),
                                        self._stream(stream)))
        return out

    def gather_code_ranges(self, slots: Sequence[int], starts: Sequence[int], counts: Sequence[int], stream=None) -> torch.Tensor:
        out = torch.empty((max(1, int(sum(counts))),), dtype=torch.int32, device=self.device)
        check(self.lib.lvx_gather_code_ranges(self._h, i32_array(slots), i32_array(starts), i32_array(counts), len(slots),
                                              C.c_void_p(out.data_ptr()), self._stream(stream)))
        return out[: int(sum(counts))]

    # ------------------------------------------------------------------ embeddings
    def codes_to_features(self, codes: torch.Tensor, stream=None) -> torch.Tensor:
        codes = codes.to(device=self.device, dtype=torch.int32).contiguous().view(-1)
        out = torch.empty((codes.numel(), self.cfg.code_dim), dtype=torch.float32, device=self.device)
        check(self.lib.lvx_codes_to_features(self._h, C.c_void_p(codes.data_ptr()), codes.numel(), C.c_void_p(out.data_ptr()),
                                             self._stream(stream)))
        return out

    def text_embed(self, ids: torch.Tensor, stream=None) -> torch.Tensor:
        ids = ids.to(device=self.device, dtype=torch.int32).contiguous().view(-1)
        out = torch.empty((ids.numel(), self.cfg.text_dim), dtype=torch.float32, device=self.device)
        check(self.lib.lvx_text_embed(self._h, C.c_void_p(ids.data_ptr()), ids.numel(), C.c_void_p(out.data_ptr()),
                                      self._stream(stream)))
        return out

    # ------------------------------------------------------------------ vocoder
    def vocode(self, codes: torch.Tensor, cu: Sequence[int], bandwidth_id: int = 0, out: Optional[torch.Tensor] = None,
               stream=None) -> torch.Tensor:
        """codes: packed int32 device tensor; chunk i = codes[cu[i]:cu[i+1]].  -> packed PCM (hop * total,) fp32."""
        assert codes.dtype == torch.int32 and codes.is_cuda and codes.is_contiguous()
        total = int(cu[-1])
        assert codes.numel() >= total
        if out is None:
            out = torch.empty((total * self.cfg.hop,), dtype=torch.float32, device=self.device)
        check(self.lib.lvx_vocode(self._h, C.c_void_p(codes.data_ptr()), i32_array(cu), len(cu) - 1, bandwidth_id,
                                  C.c_void_p(out.data_ptr()), self._stream(stream)))
        return out

    def vocode_features(self, feats: torch.Tensor, cu: Sequence[int], bandwidth_id: int = 0, stream=None) -> torch.Tensor:
        """feats: packed channels-last (total, code_dim) fp32 device tensor."""
        feats = feats.to(device=self.device, dtype=torch.float32).contiguous()
        total = int(cu[-1])
        assert feats.shape == (total, self.cfg.code_dim)
        out = torch.empty((total * self.cfg.hop,), dtype=torch.float32, device=self.device)
        check(self.lib.lvx_vocode_features(self._h, C.c_void_p(feats.data_ptr()), i32_array(cu), len(cu) - 1, bandwidth_id,
                                           C.c_void_p(out.data_ptr()), self._stream(stream)))
        return out

    def test_gemm(self, A: torch.Tensor, Wt: torch.Tensor, taps: int = 1, stream=None) -> torch.Tensor:
        """Test hook (lvx_test_gemm): A (M, K/taps) . Wt (N, K)^T -> (M, N) through the engine's GEMM path."""
        A = A.to(device=self.device, dtype=torch.float32).contiguous()
        Wt = Wt.to(device=self.device, dtype=torch.float32).contiguous()
        M, N, K = A.shape[0], Wt.shape[0], Wt.shape[1]
        assert A.shape[1] * taps == K
        out = torch.empty((M, N), dtype=torch.float32, device=self.device)
        check(self.lib.lvx_test_gemm(self._h, C.c_void_p(A.data_ptr()), C.c_void_p(Wt.data_ptr()), M, N, K, taps,
                                     C.c_void_p(out.data_ptr()), self._stream(stream)))
        return out

    _STAGE_WIDTH = {0: "voc_dim", 1: "voc_dim", 2: "voc_dim", 3: "voc_dim", 4: "voc_dim", 5: "n_fft"}

    def vocode_stage(self, codes: torch.Tensor, stage: int, bandwidth_id: int = 0, stream=None) -> torch.Tensor:
        codes = codes.to(device=self.device, dtype=torch.int32).contiguous().view(-1)
        width = getattr(self.cfg, self._STAGE_WIDTH[stage])
        out = torch.empty((codes.numel(), width), dtype=torch.float32, device=self.device)
        check(self.lib.lvx_vocode_stage(self._h, C.c_void_p(codes.data_ptr()), codes.numel(), bandwidth_id, stage,
                                        C.c_void_p(out.data_ptr()), self._stream(stream)))
        return out
