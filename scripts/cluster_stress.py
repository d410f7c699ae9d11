"""Stress of the cluster-resident decode kernel at and above its co-residency: batches of 144..256 sessions forced
through the cluster path, repeated open / feed / decode / vocode rounds through BatchSynthesizer, 1 and 4 lanes.
bf16 greedy batches above 112 sessions run on the 8-CTA cut (15 co-resident clusters; LLMVOX_B200_CD_CAP8=64 lifts the
cap, LLMVOX_B200_CD_NO8=1 selects the 16-CTA cut with its cap of 7, LLMVOX_B200_CD_CAP=64 lifts that one); exact
precision always runs the 16-CTA cut.  Prints the timeout records of the bounded spins (LLMVOX_B200_CD_DIAG=1) if a
launch dies."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("LLMVOX_B200_CD_DIAG", "1")
from llmvox_b200 import weights as W
from llmvox_b200.engine import Engine
from llmvox_b200.streaming import BatchSynthesizer, LaneRunner

LaneRunner.CLUSTER_DECODE_MAX_BATCH = 1 << 30       # force the cluster path at every batch size
LaneRunner.HYBRID_TAIL_MAX = 0
LaneRunner.CLUSTER8_MAX_WAVES = 1 << 20
REPS = int(os.environ.get("STRESS_REPS", "40"))
sd = W.make_random_weights(1234, wpe_rows=256)
e = Engine(sd, device=0, precision=os.environ.get("STRESS_PRECISION", "bf16"), max_sessions=256, max_batch=256, max_context=256, max_vocode_frames=256 * 170, decode_lanes=8)
rng = np.random.RandomState(0)
for B in [int(x) for x in os.environ.get("STRESS_B", "144,192,256").split(",")]:
    for lanes in (1, 4):
        texts = [rng.randint(3, 259, size=200).tolist() for _ in range(B)]
        bs = BatchSynthesizer(e, B, 160, stop_on_eoa=False, lanes=lanes)
        ts = []
        for rep in range(REPS):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            try:
                bs.start(texts)
                for _ in bs.run(160, flush_tail=False, copy=False):
                    pass
            except Exception as ex:
                print("FAILED", B, lanes, rep, f"{1e3 * (time.perf_counter() - t0):.1f} ms", type(ex).__name__, flush=True)
                f = e.lib.lvx_cluster_diag
                f.restype = ctypes.POINTER(ctypes.c_ulonglong)
                d = f()
                if d:
                    print("diag records:", d[0])
                    names = {16: "act", 17: "tmem_full", 18: "xbar0", 19: "xbar1"}
                    for i in range(min(24, d[0])):
                        r = [d[1 + 10 * i + k] for k in range(10)]
                        v = r[0]
                        blk = (v >> 12) & 0x3ffff
                        bidx = ((v >> 32) & 511) // 8
                        nm = names.get(bidx, f"full{bidx}" if bidx < 8 else f"empty{bidx - 8}" if bidx < 16 else f"tile{bidx - 20}")
                        def fmt(w):
                            return f"{w >> 40}.{(w >> 32) & 0xff:02x}/{(w >> 16) & 0xffff}.{w & 0xffff}"
                        print(f"  cluster {blk // 16} rank {blk % 16} warp {(v & 0xfff) // 32} waits {nm} parity {(v >> 30) & 1}{' (cluster scope)' if (v >> 31) & 1 else ''}"
                              f" | barrier word {r[9]:#x} | workers {fmt(r[1])} issuers {fmt(r[2])} {fmt(r[3])} {fmt(r[4])} producer iter {r[5] >> 32} gi {r[5] & 0xffffffff}")
                os._exit(3)
            ts.append(1e3 * (time.perf_counter() - t0))
        print(f"B={B} lanes={lanes}: {REPS} rounds ok, median {np.median(ts):.1f} ms", flush=True)
print("stress ok")
