import os, sys
os.environ["LLMVOX_B200_FUSED"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmvox_b200 import weights as W
from llmvox_b200.engine import Engine
sd = W.make_random_weights(1234, wpe_rows=64)
for n in (97, 64, 33):
    e = Engine(sd, device=0, precision="bf16", max_sessions=n, max_context=48, max_vocode_frames=256)
    rng = np.random.RandomState(n)
    texts = [rng.randint(3, 259, size=rng.randint(0, 40)).tolist() for _ in range(n)]
    slots = list(range(n))
    e.open(slots); e.feed_text(slots, texts)
    L1 = []
    for t in range(6):
        e.decode_steps(slots, 1); L1.append(e.peek_logits(n).cpu())
    c1 = e.gather_codes(slots, 0, 6).cpu().numpy()
    for k in (2, 3, 6):
        e.open(slots); e.feed_text(slots, texts)
        e.decode_steps(slots, k)
        d = (e.peek_logits(n).cpu() - L1[k - 1]).abs().max(dim=1).values
        bad = [int(i) for i in torch.nonzero(d > 0).view(-1)]
        print(f"n={n} launch of {k}: differing sessions {bad} diffs {[round(float(d[i]),4) for i in bad]} textlen {[len(texts[i]) for i in bad]} codes {[c1[i,:k].tolist() for i in bad]}")
    e.close()
