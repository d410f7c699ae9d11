"""Text front-end: host pipeline (clean_text regexes + sentence_ids + feed_text) against the device op (feed_sentences)."""
import os, sys, time, random
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmvox_b200 import weights as W
from llmvox_b200.engine import Engine
from llmvox_b200.protocol import clean_text
from llmvox_b200.tokenizer import sentence_ids
sd = W.make_random_weights(1234, wpe_rows=1024)
rng = random.Random(0)
words = ["hello", "world,", "the", "price", "is", "1,250", "#3", "A&B", "e-mail", "wait...", "5.", "path/to", "**bold**", "café", "x@y"]
for n in (1, 64, 256, 1024):
    e = Engine(sd, device=0, precision="bf16", max_sessions=n, max_context=1024, max_vocode_frames=256)
    slots = list(range(n))
    sents = [" ".join(rng.choice(words) for _ in range(20)) for _ in slots]
    for fn in ("host", "device"):
        ts = []
        for rep in range(12):
            e.open(slots)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if fn == "host":
                e.feed_text(slots, [sentence_ids(clean_text(s)) for s in sents])
            else:
                e.feed_sentences(slots, sents, clean=True)
            torch.cuda.synchronize()
            ts.append(1e6 * (time.perf_counter() - t0))
        ts.sort()
        print(f"{n:5d} sentences ({sum(len(s.encode()) for s in sents) / n:.0f} bytes each) {fn:6s}: median {ts[len(ts) // 2]:9.1f} us", flush=True)
    e.close()
