#!/usr/bin/env python
"""Benchmark of the LLMVoX speech-synthesis hot path (BASELINE.json: audio-seconds generated per second).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision bf16|fp32]

One "step" = one pass of the hot path over one batch of synthetic input = BASELINE config 1:
64 concurrent streams x 200 codes each (KV-cached decode, greedy) + the chunked vocoder (reference schedule
10 / 30 / 90 + 70-code tail per stream) = 12,800 codes = 170.67 s of 24 kHz audio per GPU.

  value : whole-job audio-s/s with the text ids already resident on the device (sessions opened and fed
          before the timed region), PCM left on the device.
  e2e   : the same work through the public host API (BatchSynthesizer): text ids start in host memory, are
          copied to the device inside the timed region, and every chunk's PCM is copied to pinned host memory.
  roofline : per-launch CUDA-event timing of the dominant kernel (lvx_profile_*), taken on a replay of the same
          step beside the timed region.
  cpu_baseline : the CPU oracle (oracle/llmvox_oracle.py, a restatement of the reference pinned to reference
          fixtures) timed on this box's host cores on a bounded sample of the same workload.

Sessions are independent, so N GPUs = N engines each running its own 64 streams (weak scaling, no collective on
the data path); NCCL is used for the barrier and the max-over-ranks reduction of the elapsed time only."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STREAMS = 64
TOKENS = 200
SCHEDULE = [10, 30, 90, 70]          # reference chunk schedule for 200 codes (replica 0) + flushed tail
CODES_PER_SEC = 75.0                 # 24 kHz / hop 320
SEED = 1234
# dram__bytes_read.sum + dram__bytes_write.sum per launch of cluster_decode_kernel, from the committed `ncu --set full`
# capture of one step's four decode launches (profiles/r01d_cluster_decode.md section 2: mean of 0.48 / 2.00 / 10.53 / 12.71 GB)
CLUSTER_TRAFFIC = 6.43e9


def synthetic_text(n_streams: int, seed: int):
    """Uniform random bytes in [a-z ] formed into words, ByT5 ids (byte + 3) with </s> = 1 per word
    (SURVEY.md section 8d)."""
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n_streams):
        ids = []
        while len(ids) < TOKENS:
            w = rng.randint(ord("a"), ord("z") + 1, size=rng.randint(1, 9))
            ids.extend((w + 3).tolist())
            ids.append(1)
        out.append(ids[:TOKENS])
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def mark(self):
        """Start of the timed region: earlier samples (warm-up) are dropped."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_mark = getattr(self, "t_mark", 0.0)
        rows = [r for (t, r) in self.rows if t >= t_mark] or [r for (_, r) in self.rows[-3:]]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# --------------------------------------------------------------------------------------------- CPU reference arm
def cpu_sample(n_streams: int, threads: int):
    """The oracle on `n_streams` streams of the workload, one after the other (the reference is batch 1 per
    thread): 200 greedy codes + the four chunk decodes each.  Returns (audio seconds, wall seconds)."""
    from llmvox_b200 import weights as W
    from oracle import llmvox_oracle as O
    torch.set_num_threads(threads)
    sd = cpu_sample.sd if hasattr(cpu_sample, "sd") else W.make_random_weights(SEED, wpe_rows=TOKENS)
    cpu_sample.sd = sd
    texts = synthetic_text(n_streams, 99)
    t0 = time.perf_counter()
    codes_total = 0
    for ids in texts:
        codes = O.decode_steps(sd, O.GPTArch(), ids, TOKENS)
        pos = 0
        for L in SCHEDULE:
            O.vocoder_decode(sd, codes[pos:pos + L])
            pos += L
        codes_total += len(codes)
    dt = time.perf_counter() - t0
    return codes_total / CODES_PER_SEC, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    per_step = 2                                   # streams per step: ~2-4 s of CPU work
    for _ in range(args.warmup):
        cpu_sample(1, threads)
    t0 = time.perf_counter()
    audio = 0.0
    for _ in range(args.steps):
        a, _ = cpu_sample(per_step, threads)
        audio += a
    dt = time.perf_counter() - t0
    v = audio / dt
    sample = f"{per_step} of the {STREAMS} streams per step, sequential batch-1 decode of {TOKENS} codes + chunks {SCHEDULE}"
    line = {"impl": "reference", "metric": "audio-sec generated/sec", "value": v, "unit": "audio-s/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "config1: 64 streams x 200 codes, KV-cached greedy decode + chunked vocoder (10/30/90/70)",
                       "sample": sample},
            "cpu_baseline": {"value": v, "unit": "audio-s/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- GPU arm
def device_step(e, runner, slots, pcm_out, alone=False):
    """One pass over one batch whose text already sits on the device: decode to each chunk boundary (on the decode
    lanes), vocode the 64 chunks that became ready as one ragged batch (control stream; overlaps the lanes' next
    iterations).  No host synchronisation, except with `alone` (the per-kernel profile leg: every decode launch is timed
    without the vocoder running beside it)."""
    pos = 0
    for L in SCHEDULE:
        runner.decode(slots, L)
        if alone:
            torch.cuda.synchronize()
        codes = e.gather_codes(slots, pos, L)                       # (64, L) int32 on the device
        cu = list(range(0, (len(slots) + 1) * L, L))
        e.vocode(codes.view(-1), cu, 0, out=pcm_out[pos * len(slots) * 320:(pos + L) * len(slots) * 320])
        pos += L


def run_gpu(args):
    import torch.distributed as dist
    from llmvox_b200 import build as B
    from llmvox_b200 import weights as W
    from llmvox_b200.engine import Engine
    from llmvox_b200.streaming import BatchSynthesizer, LaneRunner

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B.build()
    K, Wm = args.steps, max(args.warmup, 3)
    sd = W.make_random_weights(SEED, wpe_rows=256)
    n_groups = Wm + K
    e = Engine(sd, device=local, precision=args.precision, max_sessions=STREAMS * (n_groups + 1), max_batch=STREAMS,
               max_context=208, max_vocode_frames=STREAMS * 96, decode_lanes=args.lanes)
    runner = LaneRunner(e, args.lanes)
    texts = synthetic_text(STREAMS, 1000 + rank)
    groups = [list(range(g * STREAMS, (g + 1) * STREAMS)) for g in range(n_groups + 1)]
    pcm = torch.empty((STREAMS * TOKENS * 320,), dtype=torch.float32, device=e.device)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident (sessions opened + text fed before the clock starts)
    for g in groups[:n_groups]:
        e.open(g)
        e.feed_text(g, texts)
    runner.sync_from_control()
    clocks = ClockSampler(local)
    for g in groups[:Wm]:
        device_step(e, runner, g, pcm)
    barrier()
    clocks.mark()
    l0 = e.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for g in groups[Wm:Wm + K]:
        device_step(e, runner, g, pcm)
    ev1.record()
    barrier()
    launches = e.kernel_launches - l0
    ms = ev0.elapsed_time(ev1)
    time.sleep(0.15)      # let nvidia-smi emit the sample that covers the end of the region
    clk = clocks.stop()
    assert torch.isfinite(pcm[:: 4099]).all()

    # ---- e2e: host text ids in, PCM out to pinned host memory, through the public batched API
    bs = BatchSynthesizer(e, STREAMS, SCHEDULE[0], stop_on_eoa=False, slots=groups[n_groups], lanes=args.lanes)
    e2e_steps = max(3, min(K, 10))

    def e2e_step():
        bs.start(texts)                                               # H2D of slots + text ids
        n = 0
        for chunks in bs.run(TOKENS, flush_tail=True, copy=False):     # D2H of every chunk into pinned memory
            n += sum(c.length for c in chunks)
        return n
    e2e_step()
    # p50 first-chunk latency (the metric's second half): host text ids -> first 10-code chunk's PCM in host memory
    fc = []
    for _ in range(9):
        torch.cuda.synchronize()
        tf = time.perf_counter()
        bs.start(texts)
        gen = bs.run(SCHEDULE[0], flush_tail=False, copy=False)
        first = next(gen)
        fc.append(1e3 * (time.perf_counter() - tf))
        assert len(first) == STREAMS
        for _ in gen:
            pass
    first_chunk_ms = float(np.median(fc[2:]))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        assert e2e_step() == STREAMS * TOKENS
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    barrier()

    # ---- per-kernel profile of one step (events around every launch), beside the timed region
    e.open(groups[0])
    e.feed_text(groups[0], texts)
    runner.sync_from_control()
    torch.cuda.synchronize()
    e.profile(True)
    device_step(e, LaneRunner(e, 1), groups[0], pcm, alone=True)
    prof = e.profile_report()
    e.profile(False)

    # ---- vocoder bulk leg (BASELINE config 3 shape, scaled to one launch group): tensor-pipe roofline of the GEMM-bound half
    voc = None
    if rank == 0 and not args.short:
        vb_streams, vb_len = 48, 1280
        g = torch.Generator().manual_seed(3)
        vcodes = torch.randint(0, 4096, (vb_streams * vb_len,), generator=g).to(e.device, torch.int32)
        from llmvox_b200.engine import Engine as _E
        ve = _E(sd, device=local, precision=args.precision, max_sessions=2, max_context=32, max_vocode_frames=vb_streams * vb_len + 64)
        vcu = list(range(0, (vb_streams + 1) * vb_len, vb_len))
        vout = torch.empty((vb_streams * vb_len * 320,), dtype=torch.float32, device=e.device)
        for _ in range(2):
            ve.vocode(vcodes, vcu, out=vout)
        torch.cuda.synchronize()
        v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        v0.record()
        for _ in range(3):
            ve.vocode(vcodes, vcu, out=vout)
        v1.record()
        torch.cuda.synchronize()
        vms = v0.elapsed_time(v1) / 3
        vfl = vb_streams * vb_len * 132.78e6          # algorithmic FLOPs per frame at L = 1280
        ve.profile(2)
        ve.vocode(vcodes, vcu, out=vout)
        vrep = ve.profile_report()
        ve.profile(False)
        ve.close()
        pk = measured_peaks()
        voc = {"workload": f"{vb_streams} streams x {vb_len} codes, one chunk each (config 3 shape)", "ms": vms,
               "audio_s_per_s": vb_streams * vb_len / CODES_PER_SEC / (vms / 1e3), "tflops": vfl / (vms / 1e3) / 1e12,
               "frac_of_sustained_peak": vfl / (vms / 1e3) / 1e12 / pk["bf16_tflops_sustained"],
               "gemm_tflops": {k: round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1) for k, v in vrep.items() if v["flops"] > 0},
               "share_ms": {k: round(v["ms"], 3) for k, v in sorted(vrep.items(), key=lambda kv: -kv[1]["ms"])}}

    t = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=e.device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])
    audio_per_step = world * STREAMS * TOKENS / CODES_PER_SEC
    value = audio_per_step * K / (ms / 1e3)
    e2e_value = audio_per_step * e2e_steps / (e2e_ms / 1e3)

    if rank == 0:
        peaks = measured_peaks()
        total_ms = sum(v["ms"] for v in prof.values())
        top = max(prof.items(), key=lambda kv: kv[1]["ms"])
        name, r = top
        if name.startswith("tc_gemm") and name != "tc_gemm_swap":
            ach = r["flops"] / (r["ms"] / 1e3) / 1e12
            roof = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                    "frac": ach / peaks["bf16_tflops_sustained"], "traffic": None}
        else:
            ach = r["bytes"] / (r["ms"] / 1e3) / 1e9 if r["bytes"] else 0.0
            # traffic: dram__bytes_read+write per launch from the committed `ncu --set full` capture of the decode GEMMs
            # (profiles/r01c_*: 68.8 MB over the 17 GEMMs of one iteration, cold cache) -- 1.09x the algorithmic bytes
            roof = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"], "traffic": CLUSTER_TRAFFIC if name == "cluster_decode" else 4.05e6 if name == "tc_gemm_swap" else None,
                    "algorithmic_bytes_per_launch": r["bytes"] / max(1, r["launches"])}
        roof.update({"launches_per_step": r["launches"], "avg_launch_us": 1e3 * r["ms"] / max(1, r["launches"]),
                     "share_of_step": r["ms"] / total_ms, "peak_source": peaks["source"] + " (MEASURED_PEAKS.json)"})
        threads = os.cpu_count() or 1
        cpu_streams = 6
        cpu_sample(1, threads)
        ca, cdt = cpu_sample(cpu_streams, threads)
        bytes_in = STREAMS * TOKENS * 4 + (STREAMS + 1) * 4 + STREAMS * 4
        line = {"metric": "audio-sec generated/sec", "value": value, "unit": "audio-s/s", "n_gpus": world, "steps": K,
                "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.precision, "data": "synthetic",
                "config": {"workload": f"config1: {STREAMS} streams x {TOKENS} codes per GPU, KV-cached greedy decode + chunked vocoder ({'/'.join(map(str, SCHEDULE))})",
                           "weights": "random-init english-tiny GPT + frame75 WavTokenizer decoder, seed 1234",
                           "decode_lanes": args.lanes,
                           "decode_path": ("cluster-resident kernel: one call per round on a side stream, %d clusters of 16 CTAs" % ((STREAMS + 15) // 16)
                                           if runner._cluster_call(STREAMS, None) else "kernel-per-op chain on %d lanes" % args.lanes),
                           "l2": "no flush: per-step working set (189 MB bf16 weights + KV + activations) exceeds the 126 MB L2"},
                "x_realtime_per_gpu": value / world,
                # whole-step view: bf16 weights once per iteration + KV read/append, against measured HBM peak
                "roofline_step": (lambda by: {"bound": "hbm", "algorithmic_bytes": by, "achieved": by / (ms / K / 1e3) / 1e9,
                                              "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": by / (ms / K / 1e3) / 1e9 / peaks["hbm_gbs"],
                                              "note": "decode bytes of one step per GPU / step time; the step is latency-bound (DESIGN.md 5)"})(
                    TOKENS * 62914560.0 + STREAMS * 12288.0 * (TOKENS * (TOKENS + 1) / 2)),
                "roofline": roof,
                "kernels": {k: {"launches": v["launches"], "ms": round(v["ms"], 4)} for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])},
                "cpu_baseline": {"value": ca / cdt, "unit": "audio-s/s", "cores": threads, "kind": "port",
                                 "sample": f"{cpu_streams} of the {STREAMS} streams, sequential batch-1: {TOKENS} codes + chunks {SCHEDULE} each"},
                "e2e": {"value": e2e_value, "unit": "audio-s/s", "steps": e2e_steps, "h2d_bytes_per_step": bytes_in,
                        "d2h_bytes_per_step": STREAMS * TOKENS * 320 * 4},
                "first_chunk_latency": {"p50_ms": first_chunk_ms, "streams": STREAMS, "codes": SCHEDULE[0],
                                        "what": "host text ids -> PCM of every stream's first chunk in pinned host memory"},
                "vocoder_bulk": voc,
                "gpu_launches": int(launches), "clocks": clk}
        print(json.dumps(line))
    e.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="llmvox_b200", choices=["llmvox_b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--lanes", type=int, default=4, help="decode lanes: groups of sessions whose dependent chains run concurrently")
    ap.add_argument("--streams", type=int, default=64, help="concurrent streams per GPU (BASELINE config 1: 64)")
    ap.add_argument("--short", action="store_true", help="40-code utterances (chunks 10/30): a short run for ncu captures")
    args = ap.parse_args()
    global TOKENS, SCHEDULE, STREAMS
    STREAMS = args.streams
    if args.short:
        TOKENS, SCHEDULE = 40, [10, 30]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
