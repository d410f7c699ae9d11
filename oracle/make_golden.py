"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

TEST INFRASTRUCTURE (see oracle/llmvox_oracle.py header).  Run here, where /root/reference exists:

    python oracle/make_golden.py            # writes tests/golden/

What runs is the reference's own code, imported from /root/reference (nothing is copied into this
repo): ``src.model.GPT``, ``decoder.pretrained.WavTokenizer.from_hparams0802`` and the body of
``streaming_server.audio_generator_sync``, whose source text is cut out of the reference file with
``ast`` at run time and exec'd against a stub ``model_handler`` (importing ``streaming_server`` itself
would pull FastAPI / whisper / soundfile, which are absent).  The seeded weights of
``llmvox_b200.weights`` are loaded into the reference modules with ``load_state_dict``.

The GPU box has no /root/reference; tests there regenerate the same seeded weights and compare the CUDA
path (and the oracle) with these fixtures.
"""
from __future__ import annotations

import ast
import contextlib
import io
import os
import queue
import sys
import threading

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, ROOT)

from llmvox_b200 import weights as W  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
YAML = os.path.join(REF, "WavTokenizer/configs/wavtokenizer_smalldata_frame75_3s_nq1_code4096_dim512_kmeans200_attn.yaml")

# config 0 of BASELINE.json: one 20-word sentence, greedy, batch 1.
SENTENCE_20 = ("the quick brown fox jumps over the lazy dog while seven small birds "
               "sing a very old song near the river.")
SEED = 1234
VOC_LENGTHS_FULL = (1, 2, 3, 5, 10, 30, 90, 160)
VOC_LENGTHS_SLICED = (270, 480, 810, 1280)


class _Stop(Exception):
    pass


def build_reference(sd, arch):
    sys.path[:0] = [REF, os.path.join(REF, "WavTokenizer")]
    from src.model import GPT, GPTConfig
    from decoder.pretrained import WavTokenizer
    with contextlib.redirect_stdout(io.StringIO()):
        gpt = GPT(GPTConfig(block_size=arch.block_size, vocab_size=arch.vocab_size, n_layer=arch.n_layer,
                            n_head=arch.n_head, n_embd=arch.n_embd, dropout=0.0, bias=arch.bias, is_train=False))
        wav = WavTokenizer.from_hparams0802(YAML)
    missing, unexpected = gpt.load_state_dict({k: v for k, v in sd.items()
                                               if k.startswith("transformer.") or k.startswith("lm_head.")}, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    wsd = wav.state_dict()
    n = 0
    for k, v in sd.items():
        if k.startswith("backbone.") or k.startswith("head.") or k.startswith("feature_extractor."):
            assert wsd[k].shape == v.shape, (k, wsd[k].shape, v.shape)
            wsd[k] = v
            n += 1
    wav.load_state_dict(wsd)
    # every decode-path key of the reference must have come from our generator
    ours = {k for k in sd if k.startswith(("backbone.", "head."))}
    theirs = {k for k in wav.state_dict() if k.startswith(("backbone.", "head."))}
    assert ours == theirs, (ours ^ theirs)
    gpt.eval()
    wav.eval()
    return gpt, wav


def reference_tokenizer():
    from transformers import ByT5Tokenizer
    tok = ByT5Tokenizer()
    tok.add_special_tokens(dict(pad_token="[PAD]"))   # inference/model_handler.py:92-102
    tok.add_special_tokens(dict(pad_token="EOS"))
    assert len(tok) == 386
    return tok


def load_audio_generator_sync(cfg):
    """exec the reference's audio_generator_sync (streaming_server.py:250-426) from its source text."""
    src = open(os.path.join(REF, "streaming_server.py")).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "audio_generator_sync")
    code = ast.get_source_segment(src, fn)
    import time
    import torch.nn.functional as F
    ns = dict(torch=torch, F=F, time=time, Queue=queue.Queue, ModelHandler=object, config=cfg)
    exec(compile(code, "streaming_server.py::audio_generator_sync", "exec"), ns)
    return ns["audio_generator_sync"]


class StubHandler:
    """The four attributes audio_generator_sync touches (model_handler.py:61-63) + .device."""

    def __init__(self, gpt, wav, tok, table, max_steps, scripted=None):
        self.device = torch.device("cpu")
        self.wavtokenizer = wav
        self.tokenizer = tok
        self.llm_model = torch.nn.Embedding.from_pretrained(table, freeze=True)
        self._gpt = gpt
        self.max_steps = max_steps
        self.scripted = scripted
        self.logits = []
        self.in_T = []
        self.calls = 0

    def model(self, emb, kvcache=None):
        if self.calls >= self.max_steps:
            raise _Stop()
        self.in_T.append(int(emb.shape[1]))
        if self.scripted is None:
            logits, loss, kv = self._gpt(emb, kvcache=kvcache)
        else:  # scripted token stream: pins the chunk scheduler only
            logits = torch.zeros(1, 1, 4096)
            logits[0, 0, self.scripted[self.calls]] = 1.0
            loss, kv = None, [1]
        self.logits.append(logits[0, -1].clone())
        self.calls += 1
        return logits, loss, kv


class _CountingQueue(queue.Queue):
    """Audio queue that stops the reference loop (which otherwise runs forever) after `limit` control tokens."""

    def __init__(self, limit):
        super().__init__()
        self.limit, self.seen = limit, 0

    def put(self, item, *a, **kw):
        super().put(item, *a, **kw)
        if not isinstance(item, (bytes, bytearray)) and item is not None:
            self.seen += 1
            if self.seen >= self.limit:
                raise _Stop()


def run_reference_loop(fn, handler, words, dump_size, index=0, stop_after_controls=None):
    tq, aq = queue.Queue(), (queue.Queue() if stop_after_controls is None else _CountingQueue(stop_after_controls))
    for w in words:
        tq.put(w)

    class _Exhausted(queue.Queue):
        pass
    err = []

    def target():
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                fn(index, dump_size, handler, tq, aq)
        except _Stop:
            pass
        except Exception as e:  # pragma: no cover
            err.append(e)
    # the reference blocks forever on an empty text queue (:291); feed a sentinel that raises
    th = threading.Thread(target=target, daemon=True)
    th.start()
    th.join(timeout=3600)
    assert not th.is_alive(), "reference loop did not stop"
    if err:
        raise err[0]
    out = []
    while not aq.empty():
        out.append(aq.get())
    return out


def sliced(pcm: np.ndarray):
    """first 2560, middle 2560, last 2560 samples of a long chunk (keeps fixtures small)."""
    n = len(pcm)
    m = n // 2
    return np.concatenate([pcm[:2560], pcm[m - 1280:m + 1280], pcm[-2560:]])


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    arch = W.GPTArch()
    sd = W.make_random_weights(SEED)
    gpt, wav = build_reference(sd, arch)
    tok = reference_tokenizer()
    import importlib.util
    spec = importlib.util.spec_from_file_location("refcfg", os.path.join(REF, "configs/inference_config.py"))
    refcfg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(refcfg)
    cfg = dict(refcfg.config)
    fn = load_audio_generator_sync(cfg)

    # ---- tokenizer (a1)
    tok_cases = ["hello", "hello.", "", " a b ", "héllo wörld", "a[PAD]bEOS c", "naïve—dash", "EOS", "日本語", "x" * 40]
    np.savez(os.path.join(GOLD, "tokenizer.npz"),
             texts=np.array(tok_cases, dtype=object),
             ids=np.array([np.array(tok(t.strip())["input_ids"], dtype=np.int64) for t in tok_cases], dtype=object),
             allow_pickle=True)

    # ---- config 0: the real loop, real modules, 20-word sentence + 24 PAD steps after the text ends
    words = SENTENCE_20.split(" ")
    words = [w + " " for w in words[:-1]] + [words[-1]]
    text_ids = []
    for i, w in enumerate(words):
        ids = tok(w.strip())["input_ids"]
        if i == len(words) - 1:
            ids = ids + [385]
        text_ids.extend(ids)
    n_steps = len(text_ids) + 24
    h = StubHandler(gpt, wav, tok, sd["text_table"], n_steps)
    out = run_reference_loop(fn, h, words, dump_size=10, index=0)
    logits = torch.stack(h.logits).numpy()
    codes = logits.argmax(axis=1).astype(np.int32)
    chunks = [np.frombuffer(b, dtype=np.float32) for b in out if isinstance(b, (bytes, bytearray))]
    print("config0: steps", n_steps, "text ids", len(text_ids), "input T per step", h.in_T[:4], "...", h.in_T[-1],
          "chunks", [len(c) // 320 for c in chunks], "unique codes", len(set(codes.tolist())),
          "EOA seen", bool((codes == 453).any()))
    top2 = np.sort(logits, axis=1)[:, -2:]
    margin = top2[:, 1] - top2[:, 0]
    print("  logits std %.3f  margin min %.2e p10 %.2e median %.2e" % (logits.std(), margin.min(),
                                                                      np.percentile(margin, 10), np.median(margin)))
    np.savez_compressed(os.path.join(GOLD, "config0_loop.npz"),
                        text_ids=np.array(text_ids, dtype=np.int32), n_steps=n_steps, codes=codes,
                        logits_rows=logits[[0, 1, 2, 10, 50, n_steps - 1]],
                        logits_row_idx=np.array([0, 1, 2, 10, 50, n_steps - 1]),
                        logits_top_margin=margin.astype(np.float32),
                        input_T=np.array(h.in_T, dtype=np.int32),
                        chunk_lens=np.array([len(c) // 320 for c in chunks], dtype=np.int32),
                        **{f"pcm{i}": c for i, c in enumerate(chunks)})

    # ---- replica-1 schedule (initial_dump_size_2 = 160), same sentence, 200 steps
    h = StubHandler(gpt, wav, tok, sd["text_table"], 200)
    out = run_reference_loop(fn, h, words, dump_size=160, index=1)
    chunks = [np.frombuffer(b, dtype=np.float32) for b in out if isinstance(b, (bytes, bytearray))]
    codes1 = torch.stack(h.logits).numpy().argmax(axis=1).astype(np.int32)
    np.savez_compressed(os.path.join(GOLD, "replica1_loop.npz"), codes=codes1, n_steps=200,
                        chunk_lens=np.array([len(c) // 320 for c in chunks], dtype=np.int32),
                        pcm0_head=chunks[0][:3200], pcm0_tail=chunks[0][-3200:])

    # ---- teacher-forced logits straight from GPT.forward (src/model.py:201-237), full-sequence form:
    # one call with T rows and no cache returns the last row's logits; row t of the cached loop must match.
    g = torch.Generator().manual_seed(99)
    forced = torch.randint(0, 4096, (48,), generator=g)
    tids = torch.randint(3, 259, (48,), generator=g)
    import torch.nn.functional as F
    with torch.inference_mode(), contextlib.redirect_stdout(io.StringIO()):
        xs, kv, tf_logits, hist = [], None, [], None
        for t in range(48):
            te = F.embedding(tids[t].view(1, 1), sd["text_table"])
            se = torch.zeros(1, 1, 512) if t == 0 else wav.codes_to_features(forced[t - 1].view(1, 1)).permute(0, 2, 1)
            x = F.normalize(torch.cat([te, se], dim=2), p=2, dim=2, eps=1e-8)
            hist = x if hist is None else torch.cat([hist, x], dim=1)
            lg, _, kv = gpt(hist, kvcache=kv)
            tf_logits.append(lg[0, -1].clone())
    np.savez_compressed(os.path.join(GOLD, "teacher_forced.npz"), text_ids=tids.numpy().astype(np.int32),
                        forced_codes=forced.numpy().astype(np.int32), logits=torch.stack(tf_logits).numpy())

    # ---- scheduler pins: scripted token streams through the real loop (a10)
    sched = {}
    rng = np.random.RandomState(7)
    cases = {
        "no_eoa_r0": (10, [int(x) for x in rng.randint(0, 4096, 500) if x != 453][:450], ["hello there friend."]),
        "no_eoa_r1": (160, [int(x) for x in rng.randint(0, 4096, 700) if x != 453][:650], ["hello there friend."]),
        "eoa_mid": (10, [int(x) for x in rng.randint(454, 4096, 57)] + [453] + [7] * 30, ["hi.", "yo."]),
        "eoa_on_boundary": (10, [5] * 9 + [453] + [9] * 45 + [453] + [11] * 12, ["a.", "b.", "c."]),
        "eoa_first": (10, [453] + [3] * 35, ["a.", "b."]),
    }
    for name, (dump, script, wds) in cases.items():
        h = StubHandler(None, wav, tok, sd["text_table"], len(script), scripted=script)
        out = run_reference_loop(fn, h, wds + ["filler."] * 40, dump_size=dump, index=0)
        events = []
        for o in out:
            if isinstance(o, (bytes, bytearray)):
                events.append(len(o) // (4 * 320))
            elif o == "end":
                events.append(-3)
            elif o is None:
                events.append(-4)
            else:
                events.append(-1 - int(o))       # switch signal 1 -> -2, 0 -> -1
        sched[name + "_dump"] = dump
        sched[name + "_script"] = np.array(script, dtype=np.int32)
        sched[name + "_events"] = np.array(events, dtype=np.int32)
        print("sched", name, events)
    np.savez_compressed(os.path.join(GOLD, "scheduler.npz"), **sched)

    # ---- vocoder: WavTokenizer.codes_to_features + decode on independent chunks (a3, a11-a13)
    voc = {}
    g = torch.Generator().manual_seed(5)
    with torch.inference_mode():
        for L in VOC_LENGTHS_FULL + VOC_LENGTHS_SLICED:
            cds = torch.randint(0, 4096, (1, L), generator=g)
            feats = wav.codes_to_features(cds)
            pcm = wav.decode(feats, bandwidth_id=torch.tensor([0])).squeeze(0).numpy()
            assert pcm.shape == (320 * L,)
            voc[f"codes_{L}"] = cds[0].numpy().astype(np.int32)
            voc[f"pcm_{L}"] = pcm if L in VOC_LENGTHS_FULL else sliced(pcm)
            voc[f"rms_{L}"] = np.float32(np.sqrt((pcm.astype(np.float64) ** 2).mean()))
            print("vocoder L=%d rms %.4f max %.3f" % (L, voc[f"rms_{L}"], np.abs(pcm).max()))
        # bandwidth_id != 0 exercises the AdaLayerNorm row select (modules.py:81-86)
        cds = torch.randint(0, 4096, (1, 12), generator=g)
        voc["codes_bw2"] = cds[0].numpy().astype(np.int32)
        voc["pcm_bw2"] = wav.decode(wav.codes_to_features(cds), bandwidth_id=torch.tensor([2])).squeeze(0).numpy()
        # intermediate activations at L=30 for kernel-level tests
        cds = torch.from_numpy(voc["codes_30"]).long().view(1, -1)
        feats = wav.codes_to_features(cds)
        bb = wav.backbone
        x = bb.embed(feats)
        voc["act30_embed"] = x[0].T.numpy().copy()
        x = bb.pos_net[0](x)
        voc["act30_res0"] = x[0].T.numpy().copy()
        x = bb.pos_net[1](x)
        x = bb.pos_net[2](x)
        voc["act30_attn"] = x[0].T.numpy().copy()
        x = bb.pos_net[5](bb.pos_net[4](bb.pos_net[3](x)))
        voc["act30_posnet"] = x[0].T.numpy().copy()
        full = bb(feats, bandwidth_id=torch.tensor([0]))
        voc["act30_backbone"] = full[0].numpy().copy()
    np.savez_compressed(os.path.join(GOLD, "vocoder.npz"), **voc)
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)) // 1024, "KiB")


def _extract(src, tree, name):
    fn = next(n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.AsyncFunctionDef)) and n.name == name)
    return ast.get_source_segment(src, fn)


def protocol_golden():
    """Queue-protocol fixtures from the reference's own clean_text,
    text_streamer_producer and audio_generator_async (streaming_server.py:106-149, :184-248, :428-469)."""
    import asyncio
    import json
    import re
    import importlib.util
    from queue import Empty
    src = open(os.path.join(REF, "streaming_server.py")).read()
    tree = ast.parse(src)
    spec = importlib.util.spec_from_file_location("refcfg", os.path.join(REF, "configs/inference_config.py"))
    refcfg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(refcfg)
    cfg = dict(refcfg.config)
    cfg["chat_type"] = "text"
    ns = dict(torch=torch, re=re, config=cfg, Queue=queue.Queue, StreamModel=object, asyncio=asyncio, Empty=Empty, asr_model=None)
    for name in ("clean_text", "text_streamer_producer", "audio_generator_async"):
        exec(compile(_extract(src, tree, name), f"streaming_server.py::{name}", "exec"), ns)
    out = {"eos": cfg["eos_token"]}
    texts = ["  Hello **world** - it's 5. #1 & co @home ... 1,000/2 \\ ", "plain words", "a--b", "3. 4.5 x", "wait.... what", "###",
             "C:\\dir//file", "tabs\tand\nnewlines", "1,234,567 and 1, 2"]
    out["clean_text"] = [[t, ns["clean_text"](t)] for t in texts]

    class Req:
        def __init__(self, text):
            self.text = text

        def __contains__(self, k):
            return k == "text"

    class Stream:
        def __init__(self, outs):
            self.outs = outs

        def predict(self, _):
            return iter(self.outs)
    streams = [
        ["Hello", " there.", "", "-", " How", " are", " you.", " Fine", " thanks." + cfg["eos_token"]],
        ["One.", "Two.", "Three.", cfg["eos_token"]],
        ["**bold**", " a-b", " 5.", " #tag", " x & y.", " ", "tail"],
    ]
    out["router"] = []
    for outs in streams:
        q1, q2 = queue.Queue(), queue.Queue()
        with contextlib.redirect_stdout(io.StringIO()):
            ns["text_streamer_producer"](Req("prompt"), Stream(outs), q1, q2)
        out["router"].append({"outputs": outs, "q0": list(q1.queue), "q1": list(q2.queue)})

    def enc(x):
        return {"b": x.decode()} if isinstance(x, bytes) else x

    async def drain(a, b):
        qa, qb = queue.Queue(), queue.Queue()
        for x in a:
            qa.put(x)
        for x in b:
            qb.put(x)
        gen = ns["audio_generator_async"](qa, qb)
        got = []
        try:
            while True:
                got.append(await asyncio.wait_for(gen.__anext__(), timeout=2.5))
        except (asyncio.TimeoutError, StopAsyncIteration):
            pass
        return got
    mux_cases = [
        ([b"a", b"b", 1, b"e", "end", None], [b"c", b"d", 0, None]),
        ([1, b"x"], [b"y", "end", b"z", 0, b"never"]),
        ([b"p", None, b"q", "end"], []),
    ]
    out["mux"] = []
    for a, b in mux_cases:
        with contextlib.redirect_stdout(io.StringIO()):
            got = asyncio.run(drain(a, b))
        out["mux"].append({"q0": [enc(x) for x in a], "q1": [enc(x) for x in b], "yield": [enc(x) for x in got]})
    with open(os.path.join(GOLD, "protocol.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("protocol.json written:", len(out["clean_text"]), "clean_text,", len(out["router"]), "router,", len(out["mux"]), "mux cases")


def replica_stream_golden():
    """Multi-sentence answers through the reference's OWN producer, generator loops and consumer (SURVEY.md section 8f
    rows 1-2; streaming_server.py:184-248, :250-426, :428-469) with EOA-terminated sentences: four sentences for replica
    0 and three for replica 1, scripted code streams (StubHandler) whose EOA codes come after each sentence's text --
    sentence lengths chosen to hit every branch of the schedule (flush below the dump size, EOA exactly on a dump
    boundary, several dumps within one sentence, dump sizes 10/30/90/270/810 and 160/480/1280 carried across sentences).
    Records, per replica, everything its loop put on the audio queue (chunk lengths + control tokens) and the playback
    order audio_generator_async produces."""
    import asyncio
    import re
    import importlib.util
    from queue import Empty
    sd = W.make_random_weights(SEED)
    gpt, wav = build_reference(sd, W.GPTArch())
    tok = reference_tokenizer()
    spec = importlib.util.spec_from_file_location("refcfg", os.path.join(REF, "configs/inference_config.py"))
    refcfg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(refcfg)
    cfg = dict(refcfg.config)
    cfg["chat_type"] = "text"
    fn = load_audio_generator_sync(cfg)
    src = open(os.path.join(REF, "streaming_server.py")).read()
    tree = ast.parse(src)
    ns = dict(torch=torch, re=re, config=cfg, Queue=queue.Queue, StreamModel=object, asyncio=asyncio, Empty=Empty, asr_model=None)
    for name in ("clean_text", "text_streamer_producer", "audio_generator_async"):
        exec(compile(_extract(src, tree, name), f"streaming_server.py::{name}", "exec"), ns)
    eos = cfg["eos_token"]
    text = ("the quick brown fox jumps. over the lazy dog. while seven birds sing. a very old song. near the river bank. "
            "and then they rest. good night.")
    words = [w for w in text.split(" ") if w]
    outputs = [(" " if i else "") + w for i, w in enumerate(words)]
    outputs[-1] += eos
    # codes per sentence, in answer order (sentence i -> replica i % 2); the last code of each is the EOA
    lengths = [58, 170, 31, 25, 205, 500, 40]
    rng = np.random.RandomState(11)
    scripts = [[int(x) for x in rng.randint(454, 4096, n - 1)] + [cfg["eoa_token_id"]] for n in lengths]

    class Req:
        text = "prompt"

        def __contains__(self, k):
            return k == "text"

    class Stream:
        def predict(self, _):
            return iter(outputs)
    q = [queue.Queue(), queue.Queue()]
    with contextlib.redirect_stdout(io.StringIO()):
        ns["text_streamer_producer"](Req(), Stream(), q[0], q[1])
    routed = [list(q[0].queue), list(q[1].queue)]
    n_sent = [sum(1 for w in routed[r] if w.endswith(".") or eos in w) for r in (0, 1)]
    assert n_sent == [4, 3]
    out = {"outputs": np.array(outputs, dtype=object), "eoa_id": cfg["eoa_token_id"], "eos": eos,
           "lengths": np.array(lengths, dtype=np.int32),
           "scripts": np.array([np.array(x, dtype=np.int32) for x in scripts], dtype=object)}
    queues = []
    for r in (0, 1):
        script = [c for i, sc in enumerate(scripts) if i % 2 == r for c in sc] + [7] * 64
        h = StubHandler(None, wav, tok, sd["text_table"], len(script), scripted=script)
        items = run_reference_loop(fn, h, routed[r] + ["filler."] * 4, dump_size=(10, 160)[r], index=r, stop_after_controls=n_sent[r])
        events = []
        for o in items:
            if isinstance(o, (bytes, bytearray)):
                events.append(len(o) // (4 * 320))
            elif o == "end":
                events.append(-3)
            else:
                events.append(-1 - int(o))       # switch signal 1 -> -2, 0 -> -1
        out[f"events_r{r}"] = np.array(events, dtype=np.int32)
        queues.append(items + [None])
        print(f"replica {r}: {n_sent[r]} sentences, events {events}")

    async def drain(a, b):
        qa, qb = queue.Queue(), queue.Queue()
        for x in a:
            qa.put(x)
        for x in b:
            qb.put(x)
        gen = ns["audio_generator_async"](qa, qb)
        got = []
        try:
            while True:
                got.append(await asyncio.wait_for(gen.__anext__(), timeout=2.5))
        except (asyncio.TimeoutError, StopAsyncIteration):
            pass
        return got
    with contextlib.redirect_stdout(io.StringIO()):
        played = asyncio.run(drain(queues[0], queues[1]))
    out["playback"] = np.array([-3 if x is None else len(x) // (4 * 320) for x in played], dtype=np.int32)
    print("playback order:", out["playback"].tolist())
    np.savez_compressed(os.path.join(GOLD, "replica_stream.npz"), **out)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "replica_stream":
        replica_stream_golden()
    elif len(sys.argv) > 1 and sys.argv[1] == "protocol":
        sys.path[:0] = [REF]
        protocol_golden()
    else:
        main()
        protocol_golden()
        replica_stream_golden()
