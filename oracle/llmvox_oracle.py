"""CPU oracle for the LLMVoX speech-synthesis hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import it.  Nothing under ``llmvox_b200/`` imports it and the product
path has no CPU fallback.

What it is: a self-contained fp32 restatement, on torch-CPU ATen ops, of the
reference's algorithm for the path (the reference is pure PyTorch and cannot
travel to the GPU box, where ``/root/reference`` does not exist).  Every
function cites the reference ``file:line`` it follows.  Floating-point work is
restated with the same ATen calls the reference makes (``F.layer_norm``,
``F.linear``, SDPA, ``F.conv1d``, ``F.group_norm``, ``torch.bmm``,
``torch.fft.irfft``, ``F.fold``) so that, on the same torch build, it agrees
with the reference to rounding.

Parity pin: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so the pin is the reference ITSELF executed in the build
container: ``oracle/make_golden.py`` imports ``/root/reference`` unmodified,
loads the seeded weights of ``llmvox_b200.weights`` into the reference modules,
runs the reference's own ``audio_generator_sync`` / ``GPT.forward`` /
``WavTokenizer.decode`` and commits the outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this oracle against those fixtures.

Weights are passed as a flat dict ``sd`` using the reference's own state-dict
key names (``transformer.h.0.attn.c_attn.weight`` ..., ``backbone.embed.weight``
..., plus ``text_table`` for the T5 ``encoder.embed_tokens`` table).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

CODEBOOK_KEY = "feature_extractor.encodec.quantizer.vq.layers.0._codebook.embed"

# configs/inference_config.py:30-33,40-41
PAD_TOKEN_ID = 384
EOS_TEXT_ID = 385          # streaming_server.py:310
EOA_TOKEN_ID = 453
MAX_DUMP_SIZE = 1280
MAX_AUDIO_LENGTH = 8000


@dataclass
class GPTArch:
    """src/model.py:135-146 (GPTConfig); english-tiny values from configs/train_config.py:17-22."""
    n_layer: int = 4
    n_head: int = 8
    n_embd: int = 768
    block_size: int = 8192
    vocab_size: int = 4096
    bias: bool = False


# --------------------------------------------------------------------------- a1
def byt5_tokenize(text: str) -> List[int]:
    """``model_handler.tokenizer(word)["input_ids"]`` (streaming_server.py:306).

    google/byt5-small semantics: utf-8 byte + 3, trailing ``</s>`` = 1.  The two
    special tokens added at inference/model_handler.py:92-102 ("[PAD]" -> 384,
    "EOS" -> 385) are matched as literal substrings before byte encoding.
    """
    specials = (("[PAD]", 384), ("EOS", 385))
    out: List[int] = []
    i = 0
    while i < len(text):
        for lit, tid in specials:
            if text.startswith(lit, i):
                out.append(tid)
                i += len(lit)
                break
        else:
            out.extend(b + 3 for b in text[i].encode("utf-8"))
            i += 1
    out.append(1)
    return out


def word_ids(word: str, sentence_end: bool) -> List[int]:
    """streaming_server.py:305-310: strip, tokenise, ``+ [385]`` at sentence end."""
    ids = byt5_tokenize(word.strip())
    if sentence_end:
        ids = ids + [EOS_TEXT_ID]
    return ids


# --------------------------------------------------------------------------- a2/a3
def text_embed(sd: Dict[str, Tensor], ids: Tensor) -> Tensor:
    """``llm_model(ids)`` = T5 ``encoder.embed_tokens`` (model_handler.py:105; call
    streaming_server.py:315,319).  (1,n) int64 -> (1,n,256)."""
    return F.embedding(ids, sd["text_table"])


def codes_to_features(sd: Dict[str, Tensor], codes: Tensor) -> Tensor:
    """WavTokenizer/decoder/pretrained.py:209-239.  (K=1,L) or (K,B,L) -> (B,512,L)."""
    if codes.dim() == 2:
        codes = codes.unsqueeze(1)
    n_bins = sd[CODEBOOK_KEY].shape[0]
    offsets = torch.arange(0, n_bins * len(codes), n_bins)
    idx = codes + offsets.view(-1, 1, 1)
    feats = F.embedding(idx, sd[CODEBOOK_KEY]).sum(dim=0)
    return feats.transpose(1, 2)


# --------------------------------------------------------------------------- a5-a8
def new_gelu(x: Tensor) -> Tensor:
    """src/model.py:21-26 (tanh GELU)."""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def _ln(sd, prefix: str, x: Tensor) -> Tensor:
    """src/model.py:29-38: eps 1e-5, weight, optional bias."""
    w = sd[prefix + ".weight"]
    return F.layer_norm(x, w.shape, w, sd.get(prefix + ".bias"), 1e-5)


def _attn(sd, arch: GPTArch, i: int, x: Tensor, kv):
    """src/model.py:68-98 with is_train=False (model_handler.py:151): non-causal SDPA
    of the new rows against [cache ; new rows]."""
    p = f"transformer.h.{i}.attn."
    B, T, C = x.shape
    q, k, v = F.linear(x, sd[p + "c_attn.weight"], sd.get(p + "c_attn.bias")).split(arch.n_embd, dim=2)
    if kv:
        k = torch.cat([kv[0], k], dim=1)
        v = torch.cat([kv[1], v], dim=1)
    new_kv = [k, v]
    Tc = k.shape[1]
    hs = C // arch.n_head
    k = k.view(B, Tc, arch.n_head, hs).transpose(1, 2)
    q = q.view(B, T, arch.n_head, hs).transpose(1, 2)
    v = v.view(B, Tc, arch.n_head, hs).transpose(1, 2)
    y = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False)
    y = y.transpose(1, 2).contiguous().view(B, T, C)
    y = F.linear(y, sd[p + "c_proj.weight"], sd.get(p + "c_proj.bias"))
    return y, new_kv


def _mlp(sd, i: int, x: Tensor) -> Tensor:
    """src/model.py:101-116."""
    p = f"transformer.h.{i}.mlp."
    x = F.linear(x, sd[p + "c_fc.weight"], sd.get(p + "c_fc.bias"))
    x = new_gelu(x)
    return F.linear(x, sd[p + "c_proj.weight"], sd.get(p + "c_proj.bias"))


def gpt_forward(sd, arch: GPTArch, emb: Tensor, kvcache=None):
    """src/model.py:201-237 (inference branch, targets=None).

    emb (B,T,768).  Adds wpe[0..T) then, when a cache is given, keeps only the
    last row (:216-217).  Returns (logits (B,1,V), new_kvcache).
    """
    b, t, _ = emb.shape
    assert t <= arch.block_size
    pos = torch.arange(0, t, dtype=torch.long).unsqueeze(0)
    x = emb + F.embedding(pos, sd["transformer.wpe.weight"])
    if not kvcache:
        kvcache = [None] * arch.n_layer
    else:
        x = x[:, [-1], :]
    new_cache = []
    for i in range(arch.n_layer):
        a, ce = _attn(sd, arch, i, _ln(sd, f"transformer.h.{i}.ln_1", x), kvcache[i])
        x = x + a
        x = x + _mlp(sd, i, _ln(sd, f"transformer.h.{i}.ln_2", x))
        new_cache.append(ce)
    x = _ln(sd, "transformer.ln_f", x)
    logits = F.linear(x[:, [-1], :], sd["lm_head.weight"])
    return logits, new_cache


# --------------------------------------------------------------------------- a4 + a9
def assemble_input(sd, text_id: int, prev_code: Optional[int]) -> Tensor:
    """streaming_server.py:325-334: [text256 || speech512] L2-normalised (eps 1e-8);
    speech = zeros at the first step of a sentence."""
    te = text_embed(sd, torch.tensor([[text_id]]))
    if prev_code is None:
        se = torch.zeros((1, 1, 512))
    else:
        se = codes_to_features(sd, torch.tensor([[prev_code]])).permute(0, 2, 1)
    x = torch.cat([te, se], dim=2)
    return F.normalize(x, p=2, dim=2, eps=1e-8)


@torch.inference_mode()
def decode_steps(sd, arch: GPTArch, text_ids: Sequence[int], n_steps: Optional[int] = None,
                 forced_codes: Optional[Sequence[int]] = None, return_logits: bool = False):
    """The per-byte loop of streaming_server.py:323-354 for ONE sentence.

    Step t feeds ``cat([all previous inputs, x_t], dim=1)`` (``:337-338`` -- note
    ``speech_decoder_input_prev`` is assigned the *concatenated* tensor at ``:353``,
    so the input grows to (1,t+1,768) and the last row gets ``wpe[t]``).  Text ids
    beyond ``len(text_ids)`` are PAD=384 (``:316-320``).  Greedy pick =
    ``softmax(logits).argmax()`` (``:342-346``).  With ``forced_codes`` the fed-back
    code is taken from that list (teacher forcing) instead of the argmax.
    """
    n = len(text_ids) if n_steps is None else n_steps
    codes: List[int] = []
    all_logits = []
    kv = None
    prev = None
    hist = None
    for t in range(n):
        tid = text_ids[t] if t < len(text_ids) else PAD_TOKEN_ID
        x = assemble_input(sd, tid, prev)
        hist = x if hist is None else torch.cat([hist, x], dim=1)
        logits, kv = gpt_forward(sd, arch, hist, kv)
        lg = logits[:, -1, :]
        tok = int(F.softmax(lg, dim=-1).argmax(dim=-1).item())
        codes.append(tok)
        if return_logits:
            all_logits.append(lg[0].clone())
        prev = tok if forced_codes is None else int(forced_codes[t])
    if return_logits:
        return codes, torch.stack(all_logits)
    return codes


def sample_from_logits(logits: Tensor, temperature: float, top_k: Optional[int], u: Tensor) -> Tensor:
    """Sampler semantics of ``GPT.generate`` (src/model.py:397-406): ``logits / T``;
    keep entries ``>=`` the k-th largest (ties kept); softmax; one multinomial draw.

    torch's generator stream is not reproducible outside torch, so the draw is
    restated as inverse-CDF sampling in index order against a supplied uniform
    ``u`` in [0,1): the first index whose inclusive cumulative probability
    exceeds ``u``.  logits (B,V) fp32, u (B,) -> (B,) int64.
    """
    lg = logits / temperature
    if top_k is not None:
        v, _ = torch.topk(lg, min(top_k, lg.size(-1)))
        lg = lg.masked_fill(lg < v[:, [-1]], -float("inf"))
    probs = F.softmax(lg, dim=-1)
    cdf = torch.cumsum(probs.double(), dim=-1)
    idx = (cdf <= u.double().unsqueeze(1) * cdf[:, [-1]]).sum(dim=1)
    return idx.clamp(max=lg.size(-1) - 1)


# --------------------------------------------------------------------------- a11
def _swish(x):
    """WavTokenizer/decoder/models.py:10-12."""
    return x * torch.sigmoid(x)


def _gn(sd, prefix, x):
    """models.py:15-16: GroupNorm(32, C, eps=1e-6, affine)."""
    return F.group_norm(x, 32, sd[prefix + ".weight"], sd[prefix + ".bias"], 1e-6)


def _resnet_block(sd, p, x):
    """models.py:58-78 (temb=None, dropout is identity in eval, in==out channels)."""
    h = _gn(sd, p + ".norm1", x)
    h = _swish(h)
    h = F.conv1d(h, sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = _gn(sd, p + ".norm2", h)
    h = _swish(h)
    h = F.conv1d(h, sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    return x + h


def _attn_block(sd, p, x):
    """models.py:107-127: single-head non-causal attention over all L frames."""
    h_ = _gn(sd, p + ".norm", x)
    q = F.conv1d(h_, sd[p + ".q.weight"], sd[p + ".q.bias"])
    k = F.conv1d(h_, sd[p + ".k.weight"], sd[p + ".k.bias"])
    v = F.conv1d(h_, sd[p + ".v.weight"], sd[p + ".v.bias"])
    b, c, h = q.shape
    q = q.permute(0, 2, 1)
    w_ = torch.bmm(q, k)
    w_ = w_ * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    w_ = w_.permute(0, 2, 1)
    h_ = torch.bmm(v, w_)
    h_ = F.conv1d(h_, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])
    return x + h_


def _ada_ln(sd, p, x, bw_id):
    """modules.py:81-86: LN(eps 1e-6, no affine) * scale[bw] + shift[bw]."""
    scale = F.embedding(bw_id, sd[p + ".scale.weight"])
    shift = F.embedding(bw_id, sd[p + ".shift.weight"])
    x = F.layer_norm(x, (x.shape[-1],), eps=1e-6)
    return x * scale + shift


def _convnext_block(sd, p, x, bw_id):
    """modules.py:43-60: dwconv k7 -> AdaLN -> Linear -> erf GELU -> Linear -> gamma -> +res."""
    res = x
    x = F.conv1d(x, sd[p + ".dwconv.weight"], sd[p + ".dwconv.bias"], padding=3, groups=x.shape[1])
    x = x.transpose(1, 2)
    x = _ada_ln(sd, p + ".norm", x, bw_id)
    x = F.linear(x, sd[p + ".pwconv1.weight"], sd[p + ".pwconv1.bias"])
    x = F.gelu(x)
    x = F.linear(x, sd[p + ".pwconv2.weight"], sd[p + ".pwconv2.bias"])
    x = sd[p + ".gamma"] * x
    x = x.transpose(1, 2)
    return res + x


def vocos_backbone(sd, feats: Tensor, bw_id: Tensor, n_convnext: int = 12) -> Tensor:
    """models.py:223-235.  (B,512,L) -> (B,L,768)."""
    x = F.conv1d(feats, sd["backbone.embed.weight"], sd["backbone.embed.bias"], padding=3)
    x = _resnet_block(sd, "backbone.pos_net.0", x)
    x = _resnet_block(sd, "backbone.pos_net.1", x)
    x = _attn_block(sd, "backbone.pos_net.2", x)
    x = _resnet_block(sd, "backbone.pos_net.3", x)
    x = _resnet_block(sd, "backbone.pos_net.4", x)
    x = _gn(sd, "backbone.pos_net.5", x)
    x = _ada_ln(sd, "backbone.norm", x.transpose(1, 2), bw_id)
    x = x.transpose(1, 2)
    for i in range(n_convnext):
        x = _convnext_block(sd, f"backbone.convnext.{i}", x, bw_id)
    w = sd["backbone.final_layer_norm.weight"]
    return F.layer_norm(x.transpose(1, 2), w.shape, w, sd["backbone.final_layer_norm.bias"], 1e-6)


# --------------------------------------------------------------------------- a12/a13
def istft_head(sd, x: Tensor) -> Tensor:
    """heads.py:53-66: Linear 768->1282, exp, clip(max=100), cos/sin, complex S."""
    x = F.linear(x, sd["head.out.weight"], sd["head.out.bias"]).transpose(1, 2)
    mag, p = x.chunk(2, dim=1)
    mag = torch.exp(mag)
    mag = torch.clip(mag, max=1e2)
    S = mag * (torch.cos(p) + 1j * torch.sin(p))
    return istft_same(S, sd["head.istft.window"])


def istft_same(spec: Tensor, window: Tensor, n_fft: int = 1280, hop: int = 320) -> Tensor:
    """spectral_ops.py:33-75, padding="same".  (B,641,T) complex -> (B, hop*T)."""
    win_length = n_fft
    pad = (win_length - hop) // 2
    B, N, T = spec.shape
    ifft = torch.fft.irfft(spec, n_fft, dim=1, norm="backward")
    ifft = ifft * window[None, :, None]
    output_size = (T - 1) * hop + win_length
    y = F.fold(ifft, output_size=(1, output_size), kernel_size=(1, win_length), stride=(1, hop))[:, 0, 0, pad:-pad]
    window_sq = window.square().expand(1, T, -1).transpose(1, 2)
    env = F.fold(window_sq, output_size=(1, output_size), kernel_size=(1, win_length),
                 stride=(1, hop)).squeeze()[pad:-pad]
    assert (env > 1e-11).all()
    return y / env


@torch.inference_mode()
def vocoder_decode(sd, codes: Sequence[int], bw: int = 0) -> Tensor:
    """streaming_server.py:363-365: one INDEPENDENT decode of one chunk of codes
    (codes_to_features -> decode with bandwidth_id=[bw]).  -> (320*L,) fp32."""
    feats = codes_to_features(sd, torch.tensor([list(codes)], dtype=torch.long))
    bw_id = torch.tensor([bw])
    x = vocos_backbone(sd, feats, bw_id)
    return istft_head(sd, x).squeeze(0)


# --------------------------------------------------------------------------- a10
def chunk_schedule(codes: Sequence[int], dump_size: int, max_dump: int = MAX_DUMP_SIZE,
                   eoa: int = EOA_TOKEN_ID, max_audio_len: int = MAX_AUDIO_LENGTH,
                   stop_on_eoa: bool = True) -> Tuple[List[List[int]], int, int]:
    """Chunk emission of streaming_server.py:357-422 replayed over one sentence's codes.

    After every code: if ``len(pending) >= dump_size`` emit ``pending[:dump_size]``
    and triple ``dump_size`` (capped); ``elif eoa in pending`` flush everything
    (the EOA code included) and triple; then if the code was EOA (or pending
    exceeds ``max_audio_len``) the sentence ends: pending is dropped and
    ``dump_size`` is tripled once more (:418-422).  Returns (chunks,
    n_codes_consumed, dump_size_after).
    """
    pending: List[int] = []
    chunks: List[List[int]] = []
    used = 0
    for tok in codes:
        pending.append(tok)
        used += 1
        if len(pending) >= dump_size:
            chunks.append(pending[:dump_size])
            pending = pending[dump_size:]
            if dump_size < max_dump:
                dump_size = min(dump_size * 3, max_dump)
        elif stop_on_eoa and eoa in pending:
            chunks.append(pending)
            pending = []
            if dump_size < max_dump:
                dump_size = min(dump_size * 3, max_dump)
        if (stop_on_eoa and tok == eoa) or len(pending) > max_audio_len:
            pending = []
            if dump_size < max_dump:
                dump_size = min(dump_size * 3, max_dump)
            break
    return chunks, used, dump_size


@torch.inference_mode()
def synthesize_sentence(sd, arch: GPTArch, text_ids: Sequence[int], n_steps: int, dump_size: int,
                        stop_on_eoa: bool = False, flush_tail: bool = True):
    """Whole path for one session: decode ``n_steps`` codes, cut them into chunks with
    the reference schedule, vocode each chunk independently.  ``flush_tail`` also
    vocodes the codes left pending at the step cap (the bench's fixed-length
    utterances; the reference only flushes on EOA).  Returns (codes, chunks, pcm list)."""
    codes = decode_steps(sd, arch, text_ids, n_steps)
    chunks, used, _ = chunk_schedule(codes, dump_size, stop_on_eoa=stop_on_eoa)
    emitted = sum(len(c) for c in chunks)
    if flush_tail and emitted < used:
        chunks.append(list(codes[emitted:used]))
    pcm = [vocoder_decode(sd, c) for c in chunks]
    return codes, chunks, pcm
