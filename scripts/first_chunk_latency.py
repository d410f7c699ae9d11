"""First-chunk latency sweep (BASELINE config 2): wall time from handing B sentences (host text ids) to the engine until
the first chunk's PCM is in host memory, for the two reference schedules (replica 0: 10 codes, replica 1: 160 codes)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from llmvox_b200 import weights as W
from llmvox_b200.engine import Engine
from llmvox_b200.streaming import BatchSynthesizer

sd = W.make_random_weights(1234, wpe_rows=256)
e = Engine(sd, device=0, precision=os.environ.get("PROBE_PRECISION", "bf16"), max_sessions=256, max_batch=256, max_context=256,
           max_vocode_frames=256 * 170, decode_lanes=8)
rng = np.random.RandomState(0)
print("| streams | first chunk | lanes | p50 ms | p99 ms | min ms |")
print("|---:|---:|---:|---:|---:|---:|")
for dump in [int(x) for x in os.environ.get("PROBE_DUMPS", "10,160").split(",")]:
    for B in [int(x) for x in os.environ.get("PROBE_B", "1,2,4,8,16,32,64,128,256").split(",")]:
        texts = [rng.randint(3, 259, size=200).tolist() for _ in range(B)]
        best = None
        for lanes in ([int(x) for x in os.environ["PROBE_LANES"].split(",")] if "PROBE_LANES" in os.environ else (1,) if B < 32 else (1, 2, 4)):
            bs = BatchSynthesizer(e, B, dump, stop_on_eoa=False, lanes=lanes)
            ts = []
            for rep in range(12):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                bs.start(texts)
                gen = bs.run(dump, flush_tail=False, copy=False)
                try:
                    chunks = next(gen)
                except Exception as ex:
                    print("FAILED after", f"{1e3 * (time.perf_counter() - t0):.1f} ms", type(ex).__name__, flush=True)
                    import ctypes
                    f = e.lib.lvx_cluster_diag
                    f.restype = ctypes.POINTER(ctypes.c_ulonglong)
                    d = f()
                    if d:
                        print("diag records:", d[0])
                        for i in range(min(60, d[0])):
                            v = d[1 + i]
                            print(f"  bar 0x{v >> 32:x} parity {(v >> 31) & 1} block {(v >> 12) & 0x7ffff} (cluster {((v >> 12) & 0x7ffff) // 16} rank {((v >> 12) & 0x7ffff) % 16}) thread {v & 0xfff} (warp {(v & 0xfff) // 32})")
                    os._exit(3)
                if os.environ.get("PROBE_VERBOSE"):
                    print("rep", dump, B, lanes, rep, f"{1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
                ts.append(1e3 * (time.perf_counter() - t0))
                assert len(chunks) == B and chunks[0].length == dump
                for _ in gen:
                    pass
            ts = np.array(ts[2:])
            row = (float(np.percentile(ts, 50)), float(np.percentile(ts, 99)), float(ts.min()), lanes)
            if best is None or row[0] < best[0]:
                best = row
        print(f"| {B} | {dump} codes | {best[3]} | {best[0]:.2f} | {best[1]:.2f} | {best[2]:.2f} |", flush=True)
