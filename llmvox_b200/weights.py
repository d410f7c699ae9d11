"""Weights for the hot path: seeded random-init and the reference's two checkpoint formats.

State-dict key names and shapes are the reference's own:
  * LLMVoX ``torch.save`` dict ``{'model_args', 'model'}`` -- inference/model_handler.py:140-166
    (keys may carry the ``_orig_mod.`` prefix of a ``torch.compile``d model, :158-161);
  * WavTokenizer Lightning ckpt ``['state_dict']`` filtered to ``backbone.|head.|feature_extractor.``
    -- WavTokenizer/decoder/pretrained.py:95-114;
  * ``text_table``: the T5 ``encoder.embed_tokens`` table (386,256) -- model_handler.py:80-106.

There is no network in the build or bench environment, so benchmarks and parity tests use
``make_random_weights(seed)``: a self-contained torch-CPU-generator recipe (it does not import the
reference) that yields identical tensors on every machine with the same torch build.  The init
statistics follow the reference's (src/model.py:171-199, WavTokenizer/decoder/models.py:196-221) but
every affine/bias/scale that the reference initialises to a constant is perturbed so that no term of
the computation is trivially 0 or 1 in the parity tests.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch

CODEBOOK_KEY = "feature_extractor.encodec.quantizer.vq.layers.0._codebook.embed"


@dataclass
class GPTArch:
    """Same fields as the reference's ``GPTConfig`` (src/model.py:135-146) that the checkpoint's
    ``model_args`` carries (model_handler.py:149-150)."""
    n_layer: int = 4
    n_head: int = 8
    n_embd: int = 768
    block_size: int = 8192
    vocab_size: int = 4096
    bias: bool = False


@dataclass
class VocoderArch:
    """WavTokenizer frame75 yaml ``model.init_args`` (:39-65)."""
    input_channels: int = 512
    dim: int = 768
    intermediate_dim: int = 2304
    num_layers: int = 12
    adanorm_num_embeddings: int = 4
    n_fft: int = 1280
    hop_length: int = 320
    vq_bins: int = 4096


def _normal(g, shape, std):
    return torch.empty(shape, dtype=torch.float32).normal_(0.0, std, generator=g)


def _uniform(g, shape, bound):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2.0 - 1.0) * bound


def make_random_gpt(seed: int = 1234, arch: Optional[GPTArch] = None, wpe_rows: Optional[int] = None,
                    lm_head_std: float = 0.02) -> Dict[str, torch.Tensor]:
    """Seeded GPT weights.  ``wpe_rows`` limits the generated position table (default: block_size).
    ``lm_head_std`` defaults to the reference init (src/model.py:193-199)."""
    arch = arch or GPTArch()
    g = torch.Generator().manual_seed(seed)
    C = arch.n_embd
    sd: Dict[str, torch.Tensor] = {}
    rows = arch.block_size if wpe_rows is None else wpe_rows
    sd["transformer.wpe.weight"] = _normal(g, (rows, C), 0.02)
    proj_std = 0.02 / math.sqrt(2 * arch.n_layer)
    for i in range(arch.n_layer):
        p = f"transformer.h.{i}."
        sd[p + "ln_1.weight"] = 1.0 + _normal(g, (C,), 0.05)
        sd[p + "attn.c_attn.weight"] = _normal(g, (3 * C, C), 0.02)
        sd[p + "attn.c_proj.weight"] = _normal(g, (C, C), proj_std)
        sd[p + "ln_2.weight"] = 1.0 + _normal(g, (C,), 0.05)
        sd[p + "mlp.c_fc.weight"] = _normal(g, (4 * C, C), 0.02)
        sd[p + "mlp.c_proj.weight"] = _normal(g, (C, 4 * C), proj_std)
        if arch.bias:
            sd[p + "ln_1.bias"] = _normal(g, (C,), 0.02)
            sd[p + "ln_2.bias"] = _normal(g, (C,), 0.02)
            sd[p + "attn.c_attn.bias"] = _normal(g, (3 * C,), 0.02)
            sd[p + "attn.c_proj.bias"] = _normal(g, (C,), 0.02)
            sd[p + "mlp.c_fc.bias"] = _normal(g, (4 * C,), 0.02)
            sd[p + "mlp.c_proj.bias"] = _normal(g, (C,), 0.02)
    sd["transformer.ln_f.weight"] = 1.0 + _normal(g, (C,), 0.05)
    if arch.bias:
        sd["transformer.ln_f.bias"] = _normal(g, (C,), 0.02)
    sd["lm_head.weight"] = _normal(g, (arch.vocab_size, C), lm_head_std)
    return sd


def make_random_vocoder(seed: int = 4321, arch: Optional[VocoderArch] = None) -> Dict[str, torch.Tensor]:
    arch = arch or VocoderArch()
    g = torch.Generator().manual_seed(seed)
    D, I, Cin = arch.dim, arch.intermediate_dim, arch.input_channels
    sd: Dict[str, torch.Tensor] = {}
    sd[CODEBOOK_KEY] = _normal(g, (arch.vq_bins, Cin), 1.0)
    sd["backbone.embed.weight"] = _normal(g, (D, Cin, 7), 0.02)
    sd["backbone.embed.bias"] = _normal(g, (D,), 0.02)
    for i in (0, 1, 3, 4):
        p = f"backbone.pos_net.{i}."
        b = 1.0 / math.sqrt(D * 3)
        for n in ("1", "2"):
            sd[p + f"norm{n}.weight"] = 1.0 + _normal(g, (D,), 0.05)
            sd[p + f"norm{n}.bias"] = _normal(g, (D,), 0.05)
            sd[p + f"conv{n}.weight"] = _uniform(g, (D, D, 3), b)
            sd[p + f"conv{n}.bias"] = _uniform(g, (D,), b)
    p = "backbone.pos_net.2."
    sd[p + "norm.weight"] = 1.0 + _normal(g, (D,), 0.05)
    sd[p + "norm.bias"] = _normal(g, (D,), 0.05)
    b = 1.0 / math.sqrt(D)
    for n in ("q", "k", "v", "proj_out"):
        # q/k a little larger than default so the L x L softmax is not flat
        s = 3.0 if n in ("q", "k") else 1.0
        sd[p + f"{n}.weight"] = _uniform(g, (D, D, 1), b * s)
        sd[p + f"{n}.bias"] = _uniform(g, (D,), b)
    sd["backbone.pos_net.5.weight"] = 1.0 + _normal(g, (D,), 0.05)
    sd["backbone.pos_net.5.bias"] = _normal(g, (D,), 0.05)
    nE = arch.adanorm_num_embeddings
    sd["backbone.norm.scale.weight"] = 1.0 + _normal(g, (nE, D), 0.05)
    sd["backbone.norm.shift.weight"] = _normal(g, (nE, D), 0.05)
    for i in range(arch.num_layers):
        p = f"backbone.convnext.{i}."
        sd[p + "dwconv.weight"] = _normal(g, (D, 1, 7), 0.2)
        sd[p + "dwconv.bias"] = _normal(g, (D,), 0.02)
        sd[p + "norm.scale.weight"] = 1.0 + _normal(g, (nE, D), 0.05)
        sd[p + "norm.shift.weight"] = _normal(g, (nE, D), 0.05)
        sd[p + "pwconv1.weight"] = _normal(g, (I, D), 0.02)
        sd[p + "pwconv1.bias"] = _normal(g, (I,), 0.02)
        sd[p + "pwconv2.weight"] = _normal(g, (D, I), 0.02)
        sd[p + "pwconv2.bias"] = _normal(g, (D,), 0.02)
        sd[p + "gamma"] = (1.0 / arch.num_layers) * (1.0 + _normal(g, (D,), 0.2))
    sd["backbone.final_layer_norm.weight"] = 1.0 + _normal(g, (D,), 0.05)
    sd["backbone.final_layer_norm.bias"] = _normal(g, (D,), 0.05)
    b = 1.0 / math.sqrt(D)
    sd["head.out.weight"] = _uniform(g, (arch.n_fft + 2, D), b)
    hb = _uniform(g, (arch.n_fft + 2,), b)
    hb[5:9] = 5.0  # four magnitude bins sit above the exp clip (heads.py:55) so the clamp is exercised
    sd["head.out.bias"] = hb
    sd["head.istft.window"] = torch.hann_window(arch.n_fft)
    return sd


def make_random_weights(seed: int = 1234, gpt_arch: Optional[GPTArch] = None,
                        voc_arch: Optional[VocoderArch] = None, wpe_rows: Optional[int] = None,
                        lm_head_std: float = 0.02) -> Dict[str, torch.Tensor]:
    """GPT + vocoder + the (386,256) text table, all seeded from ``seed``."""
    sd = make_random_gpt(seed, gpt_arch, wpe_rows, lm_head_std)
    sd.update(make_random_vocoder(seed + 1, voc_arch))
    g = torch.Generator().manual_seed(seed + 2)
    sd["text_table"] = _normal(g, (386, 256), 1.0)
    return sd


def round_weights_to_bf16(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """fp32 tensors whose GEMM operands (ndim>=2 weights) carry bf16-representable values -- what the
    bf16 engine computes with; used to run the fp32 oracle on the same rounded weights."""
    out = {}
    for k, v in sd.items():
        if v.ndim >= 2 and k not in ("text_table", CODEBOOK_KEY, "transformer.wpe.weight") \
                and not k.endswith("scale.weight") and not k.endswith("shift.weight") \
                and "dwconv" not in k:
            out[k] = v.to(torch.bfloat16).to(torch.float32)
        else:
            out[k] = v
    return out


def fold_round_gpt_weights(sd: Dict[str, torch.Tensor], arch: Optional[GPTArch] = None) -> Dict[str, torch.Tensor]:
    """The GPT weights the engine's "exact" precision computes with, as a reference-format state dict: every LayerNorm
    weight (bias=False) is folded into the columns of the Linear that follows it (c_attn, c_fc, lm_head), every Linear
    weight is rounded to bf16, the LayerNorm weights become ones.  `LN(x) * g @ W^T == LN(x) @ (W * g)^T`, so before
    rounding this is the same model; the reference's fp32 loop run on this dict is what the exact mode's greedy
    tokens are compared with (tests/test_gpu_parity.py).  Vocoder tensors are passed through."""
    arch = arch or GPTArch()
    assert not arch.bias, "folding assumes bias=False (english-tiny)"
    out = dict(sd)
    r = lambda w: w.to(torch.bfloat16).to(torch.float32)
    for i in range(arch.n_layer):
        p = f"transformer.h.{i}."
        out[p + "attn.c_attn.weight"] = r(sd[p + "attn.c_attn.weight"] * sd[p + "ln_1.weight"][None, :])
        out[p + "mlp.c_fc.weight"] = r(sd[p + "mlp.c_fc.weight"] * sd[p + "ln_2.weight"][None, :])
        out[p + "attn.c_proj.weight"] = r(sd[p + "attn.c_proj.weight"])
        out[p + "mlp.c_proj.weight"] = r(sd[p + "mlp.c_proj.weight"])
        out[p + "ln_1.weight"] = torch.ones_like(sd[p + "ln_1.weight"])
        out[p + "ln_2.weight"] = torch.ones_like(sd[p + "ln_2.weight"])
    out["lm_head.weight"] = r(sd["lm_head.weight"] * sd["transformer.ln_f.weight"][None, :])
    out["transformer.ln_f.weight"] = torch.ones_like(sd["transformer.ln_f.weight"])
    return out


# ------------------------------------------------------------------ checkpoint formats (reference)
def load_llmvox_checkpoint(path: str):
    """inference/model_handler.py:147-163 -> (GPTArch, state dict without ``_orig_mod.``)."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    margs = ckpt["model_args"]
    arch = GPTArch(**{k: margs[k] for k in ("n_layer", "n_head", "n_embd", "block_size", "bias", "vocab_size")})
    sd = {}
    for k, v in ckpt["model"].items():
        if k.startswith("_orig_mod."):
            k = k[len("_orig_mod."):]
        sd[k] = v.float()
    return arch, sd


def load_wavtokenizer_checkpoint(path: str) -> Dict[str, torch.Tensor]:
    """WavTokenizer/decoder/pretrained.py:100-105: keep ``backbone.|head.|feature_extractor.`` keys;
    of the feature extractor only the codebook buffer is on the decode path."""
    raw = torch.load(path, map_location="cpu", weights_only=False)["state_dict"]
    sd = {}
    for k, v in raw.items():
        if k.startswith("backbone.") or k.startswith("head.") or k == CODEBOOK_KEY:
            sd[k] = v.float()
    return sd


def save_llmvox_checkpoint(path: str, arch: GPTArch, sd: Dict[str, torch.Tensor], compiled_prefix: bool = False):
    """Write the reference's LLMVoX checkpoint format (src/utils.py:143-165) -- used by tests."""
    pre = "_orig_mod." if compiled_prefix else ""
    model = {pre + k: v for k, v in sd.items() if k.startswith("transformer.") or k.startswith("lm_head.")}
    torch.save({"model_args": dict(n_layer=arch.n_layer, n_head=arch.n_head, n_embd=arch.n_embd,
                                   block_size=arch.block_size, bias=arch.bias, vocab_size=arch.vocab_size),
                "model": model, "iter_num": 0}, path)


def save_wavtokenizer_checkpoint(path: str, sd: Dict[str, torch.Tensor]):
    torch.save({"state_dict": {k: v for k, v in sd.items()
                               if k.startswith("backbone.") or k.startswith("head.") or k.startswith("feature_extractor.")}},
               path)
