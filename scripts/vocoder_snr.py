"""PCM SNR (dB) of the engine's vocoder against the reference fixtures (tests/golden/vocoder.npz), per chunk length."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmvox_b200 import weights as W
from llmvox_b200.engine import Engine

g = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "vocoder.npz"))
sd = W.make_random_weights(1234, wpe_rows=64)
e = Engine(sd, device=0, precision=os.environ.get("PROBE_PRECISION", "bf16"), max_sessions=4, max_context=64, max_vocode_frames=4096)


def snr_db(ref, x):
    ref, x = np.asarray(ref, np.float64), np.asarray(x, np.float64)
    return 10 * np.log10((ref ** 2).sum() / max(((ref - x) ** 2).sum(), 1e-300))


out = []
for L in (1, 5, 10, 30, 90, 160, 270, 480, 810, 1280):
    codes = torch.from_numpy(g[f"codes_{L}"]).to("cuda", torch.int32)
    pcm = e.vocode(codes, [0, L]).cpu().numpy()
    ref = g[f"pcm_{L}"]
    if len(ref) != len(pcm):
        m = len(pcm) // 2
        pcm = np.concatenate([pcm[:2560], pcm[m - 1280:m + 1280], pcm[-2560:]])
    out.append(f"L={L}: {snr_db(ref, pcm):.1f}")
print("PCM SNR dB:", ", ".join(out))
