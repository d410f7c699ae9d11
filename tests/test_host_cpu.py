"""CPU-only tests of the host-side logic and of the C-ABI surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from llmvox_b200 import _lib
from llmvox_b200 import weights as W
from llmvox_b200.scheduler import ChunkScheduler
from llmvox_b200.tokenizer import ByT5Tokenizer, sentence_ids

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tokenizer_matches_reference_byt5(gold):
    g = gold("tokenizer.npz")
    tok = ByT5Tokenizer()
    for text, ids in zip(g["texts"], g["ids"]):
        assert tok(str(text).strip())["input_ids"] == list(ids), text
    assert len(tok) == 386


def test_sentence_ids_match_config0_fixture(gold):
    g = gold("config0_loop.npz")
    sent = ("the quick brown fox jumps over the lazy dog while seven small birds "
            "sing a very old song near the river.")
    assert sentence_ids(sent) == g["text_ids"].tolist()


@pytest.mark.parametrize("name", ["no_eoa_r0", "no_eoa_r1", "eoa_mid", "eoa_on_boundary", "eoa_first"])
def test_chunk_scheduler_matches_reference_loop(gold, name):
    """The scripted token streams went through the reference's own audio_generator_sync (oracle/make_golden.py):
    positive event = chunk of that many codes, -2 = end of sentence."""
    g = gold("scheduler.npz")
    script = g[name + "_script"].tolist()
    sc = ChunkScheduler(dump_size=int(g[name + "_dump"]))
    events = []
    for code in script:
        for (_, length) in sc.push(code):
            events.append(length)
        if sc.done:
            events.append(-2)
            sc.new_sentence()
    assert events == g[name + "_events"].tolist()


def test_chunk_scheduler_ranges_are_contiguous_and_flush():
    sc = ChunkScheduler(dump_size=10, stop_on_eoa=False)
    out = []
    for _ in range(200):
        out += sc.push(None)
    assert out == [(0, 10), (10, 30), (40, 90)]
    assert sc.steps_to_next_event() == 270 - 70
    assert sc.flush() == [(130, 70)]
    assert sc.flush() == []


def test_c_abi_library_exports_every_declared_symbol():
    """Every function include/llmvox_b200.h declares must be exported by the built library, and the ctypes
    table must cover the header."""
    hdr = open(os.path.join(ROOT, "include", "llmvox_b200.h")).read()
    declared = set(re.findall(r"\b(lvx_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"lvx_engine", "lvx_config", "lvx_sampling"}
    assert len(declared) >= 20
    lib = _lib.load()
    bound = {n for n, _, _ in _lib.SYMBOLS}
    assert declared == bound, declared ^ bound
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.lvx_version() >= 100
    cfg = _lib.LvxConfig()
    assert lib.lvx_config_default(ctypes.byref(cfg)) == 0
    assert (cfg.n_layer, cfg.n_head, cfg.n_embd, cfg.vocab_size) == (4, 8, 768, 4096)
    assert (cfg.n_fft, cfg.hop, cfg.pad_token_id, cfg.eoa_token_id) == (1280, 320, 384, 453)
    assert lib.lvx_config_default(None) != 0
    assert b"NULL" in lib.lvx_last_error()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "llmvox_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|import_module\(.oracle|oracle/", src, re.M), f


def test_engine_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("needs a CPU-only box")
    from llmvox_b200.engine import Engine
    from llmvox_b200.model_handler import ModelHandler
    with pytest.raises(RuntimeError):
        Engine({}, device=0)
    with pytest.raises(RuntimeError):
        ModelHandler({"random_init_seed": 1}, None)


def test_checkpoint_formats_roundtrip(tmp_path):
    arch = W.GPTArch()
    sd = W.make_random_weights(7, wpe_rows=64)
    W.save_llmvox_checkpoint(str(tmp_path / "gpt.pt"), arch, sd, compiled_prefix=True)
    W.save_wavtokenizer_checkpoint(str(tmp_path / "wav.ckpt"), sd)
    arch2, g = W.load_llmvox_checkpoint(str(tmp_path / "gpt.pt"))
    v = W.load_wavtokenizer_checkpoint(str(tmp_path / "wav.ckpt"))
    assert arch2 == arch
    assert all(not k.startswith("_orig_mod.") for k in g)
    assert torch.equal(g["lm_head.weight"], sd["lm_head.weight"])
    assert torch.equal(v[W.CODEBOOK_KEY], sd[W.CODEBOOK_KEY])
    assert set(v) == {k for k in sd if k.startswith(("backbone.", "head.")) or k == W.CODEBOOK_KEY}


def test_seeded_weights_are_reproducible():
    a = W.make_random_weights(1234, wpe_rows=16)
    b = W.make_random_weights(1234, wpe_rows=16)
    assert all(torch.equal(a[k], b[k]) for k in a)
    r = W.round_weights_to_bf16(a)
    k = "transformer.h.0.attn.c_attn.weight"
    assert torch.equal(r[k], a[k].bfloat16().float()) and not torch.equal(r[k], a[k])


def test_chunk_scheduler_properties_against_the_oracle_schedule():
    """Random code streams (with and without EOA codes): the incremental scheduler must cut exactly the chunks the
    oracle's replay of streaming_server.py:357-422 cuts, ranges must tile the sentence without gaps, and the dump
    size may only grow up to max_dump."""
    from hypothesis import given, settings, strategies as st
    from oracle import llmvox_oracle as O

    @settings(max_examples=150, deadline=None)
    @given(st.lists(st.sampled_from([453, 7, 8, 9, 1000]), min_size=1, max_size=400), st.sampled_from([10, 160]),
           st.sampled_from([1280, 90]))
    def check(codes, dump, max_dump):
        sc = ChunkScheduler(dump_size=dump, max_dump=max_dump)
        pos, got = 0, []
        last_dump = dump
        while pos < len(codes):
            chunks, used, new_dump = O.chunk_schedule(codes[pos:], sc.dump_size, max_dump=max_dump)
            mine = []
            for k in range(used):
                mine += sc.push(codes[pos + k])
                assert last_dump <= sc.dump_size <= max(max_dump, dump)
                last_dump = sc.dump_size
            assert [c for (_, c) in mine] == [len(c) for c in chunks]
            start = 0
            for (s, c) in mine:
                assert s == start and c > 0
                start += c
            assert sc.dump_size == new_dump
            got.append(mine)
            pos += used
            if sc.done:
                sc.new_sentence()
            else:
                assert pos == len(codes)
    check()


def test_lane_runner_picks_the_decode_path_by_batch_size():
    """LaneRunner's host-side policy (no GPU needed).  Greedy bf16 / exact: the cluster-resident kernel takes everything up to one
    wave of 8-CTA clusters (240 sessions on a B200; the engine itself uses 16-CTA clusters up to 112), a batch slightly
    above a wave keeps the wave on the cluster kernel and puts the rest on the kernel-per-op lanes, two waves are still
    cluster work, beyond that the lanes take all; sampled decoding follows the same plan.  Without the 8-CTA cut: one or two
    waves of 7 16-CTA clusters (up to 112 and 140-224 sessions)."""
    from llmvox_b200.engine import Engine, Sampling
    from llmvox_b200.streaming import LaneRunner

    class Cfg:
        n_embd, n_head, vocab_size, bias, kv_page_tokens, max_context, text_dim, code_dim = 768, 8, 4096, 0, 16, 512, 256, 512

    class FakeEngine:
        cfg = Cfg()
        precision = "bf16"
        cluster_decode_applicable = Engine.cluster_decode_applicable

        def cluster_capacity(self):
            return (112, 240)

    r = object.__new__(LaneRunner)
    r.e = FakeEngine()
    r.G = 4
    r.cluster_default = True
    greedy, sampled = Sampling(), Sampling(greedy=False, top_k=50, temperature=0.8)
    # (sessions on the cluster kernel, sessions on the kernel-per-op lanes)
    assert [r.plan(n, greedy) for n in (1, 64, 112, 113, 139, 240, 241, 256, 272, 273, 480, 481, 2048)] == \
        [(1, 0), (64, 0), (112, 0), (113, 0), (139, 0), (240, 0), (240, 1), (240, 16), (240, 32), (273, 0), (480, 0), (0, 481), (0, 2048)]
    r.G = 1                                                        # no second lane to run the tail beside the wave
    assert r.plan(256, greedy) == (256, 0)
    r.G = 4
    # sampled decoding runs inside the cluster kernel too (S = 1), on both cuts
    assert [r.plan(n, sampled) for n in (64, 112, 113, 240, 256, 481)] == [(64, 0), (112, 0), (113, 0), (240, 0), (240, 16), (0, 481)]
    # without the 8-CTA cut (measurement knob / a device with no room for it): one or two waves of 16-CTA clusters
    r._caps = (112, 0)
    assert [r.plan(n, greedy) for n in (64, 112, 113, 139, 140, 224, 225, 256)] == \
        [(64, 0), (112, 0), (0, 113), (0, 139), (140, 0), (224, 0), (0, 225), (0, 256)]
    r.HYBRID_ABOVE_MAX_BATCH = True                                # opt-in split of one call between both paths
    assert r.plan(256, greedy) == (224, 32)
    r.HYBRID_ABOVE_MAX_BATCH = False
    r._caps = (112, 240)
    r.e.precision = "fp32"
    assert r.plan(64, greedy) == (0, 64)                          # fp32 parity mode: FMA-pipe GEMMs
    r.e.precision = "exact"
    assert r.plan(64, greedy) == (64, 0) and r.plan(120, greedy) == (120, 0)   # exact mode: hi | lo cluster kernel, both cuts
    assert r.plan(256, greedy) == (240, 16) and r.plan(256, sampled) == (240, 16)
    r.e.precision = "bf16"
    r.e.cfg.max_context = 8192
    assert r.plan(64, greedy) == (64, 0)                          # the reference's block_size: page-table windows of 64 pages
    r.e.cfg.max_context = 512
    r.cluster_default = False                                     # LLMVOX_B200_CLUSTER=0
    assert r.plan(64, greedy) == (0, 64)


def test_fold_round_gpt_weights_is_the_same_model_before_rounding(weights):
    """weights.fold_round_gpt_weights (what the engine's exact precision computes with): folding the LayerNorm weights into
    the following Linear is an identity on the model; only the bf16 rounding of the Linear weights changes logits."""
    import torch
    from llmvox_b200 import weights as W
    from oracle import llmvox_oracle as O
    g = torch.Generator().manual_seed(3)
    sd = dict(weights)
    for k in list(sd):
        if k.startswith("transformer.") and (".ln_" in k or "ln_f" in k) and k.endswith(".weight"):
            sd[k] = 1.0 + 0.2 * torch.randn(sd[k].shape, generator=g)
    folded = W.fold_round_gpt_weights(sd)
    ids = O.word_ids("fold", True)
    forced = [7, 99, 1234, 4000, 17, 5]
    _, a = O.decode_steps(sd, O.GPTArch(), ids, 6, forced_codes=forced, return_logits=True)
    _, b = O.decode_steps(folded, O.GPTArch(), ids, 6, forced_codes=forced, return_logits=True)
    assert float((a - b).abs().max()) < 2e-2           # bf16 weight rounding only
    for k, v in folded.items():
        if k.startswith("transformer.h.") and k.endswith("weight") and v.ndim == 2:
            assert torch.equal(v, v.to(torch.bfloat16).to(torch.float32)), k
        if ".ln_" in k or "ln_f" in k:
            assert torch.equal(v, torch.ones_like(v))
    # unrounded fold == original model to fp32 rounding
    unrounded = dict(sd)
    for i in range(4):
        p = f"transformer.h.{i}."
        unrounded[p + "attn.c_attn.weight"] = sd[p + "attn.c_attn.weight"] * sd[p + "ln_1.weight"][None, :]
        unrounded[p + "mlp.c_fc.weight"] = sd[p + "mlp.c_fc.weight"] * sd[p + "ln_2.weight"][None, :]
        unrounded[p + "ln_1.weight"] = torch.ones(768)
        unrounded[p + "ln_2.weight"] = torch.ones(768)
    unrounded["lm_head.weight"] = sd["lm_head.weight"] * sd["transformer.ln_f.weight"][None, :]
    unrounded["transformer.ln_f.weight"] = torch.ones(768)
    _, c = O.decode_steps(unrounded, O.GPTArch(), ids, 6, forced_codes=forced, return_logits=True)
    assert float((a - c).abs().max()) < 1e-4


def test_resize_text_table_rows_are_running_means():
    """inference/model_handler.py:22-41, applied per added token (:92-102): row 384 = mean(rows[:384]), row 385 =
    mean(rows[:385]); existing rows untouched; an already 386-row table passes through."""
    import torch
    from llmvox_b200.model_handler import resize_text_table
    g = torch.Generator().manual_seed(0)
    t = torch.randn(384, 256, generator=g)
    r = resize_text_table(t)
    assert r.shape == (386, 256) and torch.equal(r[:384], t)
    assert torch.allclose(r[384], t.mean(0), atol=1e-6)
    assert torch.allclose(r[385], torch.cat([t, r[384:385]]).mean(0), atol=1e-6)
    assert torch.equal(resize_text_table(r), r)


def test_scheduler_push_many_equals_repeated_push():
    """ChunkScheduler.push_many(k) == k x push(None): same ranges, same state, for random round lengths, dump sizes and
    length caps (including the shapes where the cap can trigger)."""
    import random
    from llmvox_b200.scheduler import ChunkScheduler
    rnd = random.Random(5)
    for _ in range(300):
        kw = dict(dump_size=rnd.choice([1, 3, 10, 160]), max_dump=rnd.choice([10, 90, 1280]), stop_on_eoa=rnd.random() < 0.5,
                  max_audio_len=rnd.choice([5, 50, 8000]))
        a, b = ChunkScheduler(**kw), ChunkScheduler(**kw)
        for _ in range(rnd.randint(1, 12)):
            if a.done:
                break
            k = rnd.randint(1, 400)
            got = a.push_many(k)
            want = []
            for _ in range(k):
                if b.done:
                    break
                want.extend(b.push(None))
            assert got == want
            assert (a.seen, a.emitted, a.dump_size, a.done, a.chunks) == (b.seen, b.emitted, b.dump_size, b.done, b.chunks)
