"""Host-side mirror of the reference's `inference/model_handler.py:45-166`.

`ModelHandler(config, device_id)` exposes the same four attributes the reference's hot loop touches
(`streaming_server.py:250-426`): `.tokenizer`, `.llm_model`, `.model`, `.wavtokenizer` (+ `.device`), each
call-compatible with the reference call shapes, and every one of them computes through the CUDA engine
(libllmvox_b200.so).  It adds the batched generate / stream-chunk API (`synthesize`, `stream`) that drives many
sessions at once.  There is no CPU path: constructing it without a CUDA device raises."""
from __future__ import annotations

import os
import weakref
from typing import Dict, Iterable, Iterator, List, Optional, Sequence

import numpy as np
import torch

from . import weights as W
from .engine import Engine, Sampling
from .scheduler import INITIAL_DUMP_SIZE_1
from .streaming import BatchSynthesizer, Chunk, synthesize
from .tokenizer import ByT5Tokenizer, sentence_ids

# configs/inference_config.py keys the loader reads (model_handler.py:66-106,140-166) plus ours
DEFAULT_CONFIG = {
    "wav_config_path": None, "wav_model_path": None, "encoder_model_path": None, "tokenizer_path": None,
    "llmvox_checkpoint_path": None,
    "initial_dump_size_1": 10, "initial_dump_size_2": 160, "max_dump_size": 1280, "max_audio_length": 8000,
    "pad_token_id": 384, "eoa_token_id": 453,
    # llmvox_b200 extensions
    "random_init_seed": None,        # seeded random-init weights when no checkpoints are available
    # "exact" (default): bf16 weights + fp32-class activations on the tcgen05 tensor cores, greedy tokens identical to the
    # reference's fp32 loop on the rounded weights; "bf16": bf16 activations too (2e-2 logits); "fp32": FMA-pipe parity mode
    "precision": "exact",
    "max_sessions": 256, "max_context": 1024, "max_vocode_frames": 32768,
    "interactive_slots": 16,         # slots reserved for the reference-style `.model(emb, kvcache)` calls
}


class KVCacheHandle:
    """The opaque `kvcache` object round-tripped by the reference loop (streaming_server.py:341).  Holds a
    session slot; the slot returns to the pool when the handle is dropped (the loop's `kvcache = None`, :412)."""

    def __init__(self, owner: "GPTModule", slot: int):
        self.slot = slot
        self.length = 0
        self._fin = weakref.finalize(self, owner._release, slot)

    def __bool__(self):
        return True

    def __len__(self):            # the reference's cache is a list of n_layer [k, v] pairs
        return 1


class GPTModule:
    """`.model(emb, kvcache=None) -> (logits, None, kvcache)` == GPT.forward (src/model.py:201-237)."""

    def __init__(self, engine: Engine, slots: Sequence[int]):
        self.e = engine
        self._free = list(slots)

    def _release(self, slot: int):
        self._free.append(slot)

    def eval(self):
        return self

    def __call__(self, emb: torch.Tensor, targets=None, kvcache: Optional[KVCacheHandle] = None):
        if targets is not None:
            raise NotImplementedError("training forward (targets) is out of scope of the inference path")
        if emb.dim() != 3 or emb.shape[0] != 1:
            raise ValueError("expected emb of shape (1, T, n_embd) as fed by streaming_server.py:341")
        t = emb.shape[1]
        if not kvcache:
            if t != 1:
                raise NotImplementedError("a cache-less forward over T > 1 rows is not on the inference path")
            if not self._free:
                raise RuntimeError("no free interactive session slot (drop old kvcache handles)")
            kvcache = KVCacheHandle(self, self._free.pop())
            self.e.open([kvcache.slot])
        if kvcache.length != t - 1:
            raise AssertionError(f"cache holds {kvcache.length} tokens but {t} input rows were given")
        if t > self.e.cfg.max_context:     # src/model.py:205
            raise AssertionError(f"Cannot forward sequence of length {t}, block size is only {self.e.cfg.max_context}")
        logits = self.e.decode_step_embeds([kvcache.slot], emb[0, -1:, :], [t - 1])
        kvcache.length = t
        return logits.view(1, 1, -1), None, kvcache


class TextEmbedding:
    """`.llm_model(ids)` == T5 encoder.embed_tokens (model_handler.py:105): (1, n) int64 -> (1, n, 256)."""

    def __init__(self, engine: Engine):
        self.e = engine

    def __call__(self, ids: torch.Tensor) -> torch.Tensor:
        shape = tuple(ids.shape)
        return self.e.text_embed(ids.reshape(-1)).view(*shape, self.e.cfg.text_dim)


class WavTokenizerModule:
    """`.wavtokenizer.codes_to_features` / `.decode` (WavTokenizer/decoder/pretrained.py:192-239)."""

    def __init__(self, engine: Engine):
        self.e = engine

    def eval(self):
        return self

    def codes_to_features(self, codes: torch.Tensor) -> torch.Tensor:
        if codes.dim() == 2:
            codes = codes.unsqueeze(1)                       # (K, B, L), pretrained.py:230-231
        if codes.shape[0] != 1:
            raise NotImplementedError("the frame75 config has a single codebook (nq = 1)")
        _, b, l = codes.shape
        feats = self.e.codes_to_features(codes.reshape(-1))  # (B*L, 512) channels-last
        return feats.view(b, l, -1).transpose(1, 2)          # (B, 512, L) like the reference

    def decode(self, features_input: torch.Tensor, bandwidth_id: Optional[torch.Tensor] = None, **kwargs) -> torch.Tensor:
        if bandwidth_id is None:
            raise ValueError("the AdaLayerNorm backbone needs bandwidth_id (modules.py:81-86)")
        b, c, l = features_input.shape
        bw = int(bandwidth_id.reshape(-1)[0].item())
        feats = features_input.transpose(1, 2).reshape(b * l, c)
        pcm = self.e.vocode_features(feats, [i * l for i in range(b + 1)], bw)
        return pcm.view(b, l * self.e.cfg.hop)


def resize_text_table(table: torch.Tensor, new_rows: int = 386) -> torch.Tensor:
    """smart_tokenizer_and_embedding_resize (inference/model_handler.py:22-41), applied once per added token as the
    reference does (:92-102: "[PAD]" -> row 384, then "EOS" -> row 385): every new row is the MEAN of all rows that exist
    before it is added, so row 384 = mean(rows[:384]) and row 385 = mean(rows[:385]) -- deterministic, unlike
    `resize_token_embeddings`' version-dependent random initialisation.  Row 384 is the text row fed on every step
    after a sentence's text runs out (:316-320)."""
    table = table.detach().float()
    rows = [table[i] for i in range(min(table.shape[0], new_rows))]
    while len(rows) < new_rows:
        rows.append(torch.stack(rows).mean(dim=0))
    return torch.stack(rows)


def _load_text_table(path: Optional[str]) -> torch.Tensor:
    """The reference pulls `encoder.embed_tokens` out of a HF T5 (model_handler.py:80-106).  Accepts a local HF
    directory or a tensor file holding the (386, 256) table."""
    if path and os.path.isfile(path):
        obj = torch.load(path, map_location="cpu")
        t = obj["weight"] if isinstance(obj, dict) and "weight" in obj else obj
        return resize_text_table(torch.as_tensor(t, dtype=torch.float32))
    if path and os.path.isdir(path):
        from transformers import T5ForConditionalGeneration
        m = T5ForConditionalGeneration.from_pretrained(path)
        return resize_text_table(m.encoder.embed_tokens.weight.detach().float())
    raise FileNotFoundError(f"text embedding table not found at {path!r} (no network: pass a local path)")


class ModelHandler:
    def __init__(self, config: Dict, device_id: Optional[int] = None):
        cfg = dict(DEFAULT_CONFIG)
        cfg.update(config or {})
        self.config = cfg
        if device_id is None or not torch.cuda.is_available():
            raise RuntimeError("llmvox_b200.ModelHandler needs a CUDA device id (sm_100a); there is no CPU fallback")
        self.device = torch.device(f"cuda:{device_id}")
        gpt_arch = W.GPTArch()
        if cfg.get("random_init_seed") is not None:
            sd = W.make_random_weights(int(cfg["random_init_seed"]))
        else:
            gpt_arch, sd = W.load_llmvox_checkpoint(cfg["llmvox_checkpoint_path"])
            sd.update(W.load_wavtokenizer_checkpoint(cfg["wav_model_path"]))
            sd["text_table"] = _load_text_table(cfg["encoder_model_path"])
        n_inter = int(cfg["interactive_slots"])
        self.engine = Engine(sd, device=device_id, precision=cfg["precision"], gpt_arch=gpt_arch,
                             max_sessions=cfg["max_sessions"] + n_inter, max_batch=cfg["max_sessions"],
                             max_context=cfg["max_context"], max_vocode_frames=cfg["max_vocode_frames"],
                             pad_token_id=cfg["pad_token_id"], eoa_token_id=cfg["eoa_token_id"])
        self._batch_slots = list(range(cfg["max_sessions"]))
        self.tokenizer = ByT5Tokenizer()
        self.llm_model = TextEmbedding(self.engine)
        self.model = GPTModule(self.engine, range(cfg["max_sessions"], cfg["max_sessions"] + n_inter))
        self.wavtokenizer = WavTokenizerModule(self.engine)

    # ------------------------------------------------------------------ batched generate / stream-chunk API
    def text_to_ids(self, sentence: str) -> List[int]:
        return sentence_ids(sentence)

    def stream(self, sentences: Sequence[str], max_steps: Optional[int] = None, replica: int = 0, stop_on_eoa: bool = True,
               flush_tail: bool = True, sampling: Optional[Sampling] = None, clean: bool = False) -> Iterator[List[Chunk]]:
        """Yields lists of ready chunks (float32 PCM) as the sessions advance; replica 0 / 1 selects
        initial_dump_size_1 / _2 (streaming_server.py:521-531).  The sentences are tokenised on the device (the ids
        `text_to_ids` gives); `clean=True` runs the reference's clean_text (:106-149) there first."""
        ids = list(sentences)
        if len(ids) > len(self._batch_slots):
            raise ValueError("more sentences than max_sessions")
        dump = self.config["initial_dump_size_1" if replica == 0 else "initial_dump_size_2"]
        # like the reference (:397), a sentence is decoded until its EOA code or max_audio_length pending codes, never by a
        # text-length heuristic; the engine's max_context is the only other bound (an explicit max_steps is the caller's)
        steps = max_steps if max_steps is not None else self.config["max_context"]
        bs = BatchSynthesizer(self.engine, len(ids), dump, self.config["max_dump_size"], stop_on_eoa, sampling,
                              slots=self._batch_slots[: len(ids)], max_audio_length=self.config["max_audio_length"])
        bs.start(ids, clean=clean)
        yield from bs.run(steps, flush_tail=flush_tail)
        self.truncated = [i for i, sc in enumerate(bs.sched) if stop_on_eoa and not sc.done]   # sessions cut by the step bound

    def synthesize(self, sentences: Sequence[str], **kw) -> List[np.ndarray]:
        """One float32 waveform per sentence (its chunks concatenated)."""
        per: List[List[np.ndarray]] = [[] for _ in sentences]
        for chunks in self.stream(sentences, **kw):
            for ch in chunks:
                per[ch.session].append(ch.pcm)
        return [np.concatenate(p) if p else np.zeros((0,), np.float32) for p in per]
