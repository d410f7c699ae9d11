#!/usr/bin/env python
"""Benchmark of the LLMVoX speech-synthesis hot path (BASELINE.json: audio-seconds generated per second).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision exact|bf16|fp32]

One "step" = one pass of the hot path over one batch of synthetic input = BASELINE config 1:
64 concurrent streams x 200 codes each (KV-cached decode, greedy) + the chunked vocoder (reference schedule
10 / 30 / 90 + 70-code tail per stream) = 12,800 codes = 170.67 s of 24 kHz audio per GPU.

  value : whole-job audio-s/s with the text ids already resident on the device (sessions opened and fed
          before the timed region), PCM left on the device.
  e2e   : the same work through the public host API (BatchSynthesizer): text ids start in host memory, are
          copied to the device inside the timed region, and every chunk's PCM is copied to pinned host memory.
  roofline : per-launch CUDA-event timing of the dominant kernel (lvx_profile_*), taken on a replay of the same
          step beside the timed region.
  other_precision : the same step in the other tensor-core precision (headline = exact: bf16 weights + fp32-class hi|lo
          activations, greedy tokens identical to the reference loop; other = bf16 activations, tolerance-level parity).
  streams256 : the north-star operating point (256 concurrent streams per GPU): value, e2e, p50 first-chunk latency at
          10 and 160 codes, roofline of its dominant kernel.
  vocoder_bulk : BASELINE config 3 at its stated size (1024 streams x 1280 codes) in launch groups.
  config4 : (under torchrun, N > 1) BASELINE config 4, 4096 sessions x 600 codes session-sharded over the ranks, e2e.
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
    streams), vocode the chunks that became ready as one ragged batch (control stream; overlaps the next iterations).
    No host synchronisation, except with `alone` (the per-kernel profile leg: every decode launch is timed without the
    vocoder running beside it)."""
    pos = 0
    for L in SCHEDULE:
        runner.decode(slots, L)
        if alone:
            torch.cuda.synchronize()
        codes = e.gather_codes(slots, pos, L)                       # (n, L) int32 on the device
        cu = list(range(0, (len(slots) + 1) * L, L))
        e.vocode(codes.view(-1), cu, 0, out=pcm_out[pos * len(slots) * 320:(pos + L) * len(slots) * 320])
        pos += L


def decode_bytes(streams, tokens, kv_bytes):
    """Algorithmic bytes of one step's decode per GPU: bf16 weights once per iteration + KV read / append."""
    return tokens * 62914560.0 + streams * 6144.0 * kv_bytes * (tokens * (tokens + 1) / 2)


def roofline_of(prof, peaks):
    """The launch class that took the most time in the profiled step.  When part of a batch runs on the kernel-per-op chain
    BESIDE a wave of the cluster-resident kernel (256 streams: 240 + 16), the chain's ~35 kernels per iteration are timed
    while they overlap the one cluster launch: summed, they can exceed it although the cluster kernel bounds the wall
    clock.  The dominant kernel is then the cluster kernel as long as its own time is at least half of the largest sum
    (`concurrent_classes` names what ran beside it)."""
    total_ms = sum(v["ms"] for v in prof.values())
    name, r = max(prof.items(), key=lambda kv: kv[1]["ms"])
    beside = None
    if name != "cluster_decode" and "cluster_decode" in prof and "tc_gemm_swap" in prof and prof["cluster_decode"]["ms"] >= 0.5 * r["ms"]:
        beside = {k: round(v["ms"], 3) for k, v in prof.items() if k in ("tc_gemm_swap", "layernorm", "decode_attention", "assemble_input", "sampler")}
        name, r = "cluster_decode", prof["cluster_decode"]
    if name.startswith("tc_gemm") and name != "tc_gemm_swap":
        ach = r["flops"] / (r["ms"] / 1e3) / 1e12
        roof = {"kernel": name, "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_tflops_sustained"]}
    else:
        ach = r["bytes"] / (r["ms"] / 1e3) / 1e9 if r["bytes"] else 0.0
        roof = {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                "algorithmic_bytes_per_launch": r["bytes"] / max(1, r["launches"])}
    # dram__bytes_read + write of the kernel is not measurable inside this run: see the committed ncu capture
    roof.update({"traffic": None, "traffic_source": "not measured in-run; ncu --set full capture under profiles/ (DRAM bytes per launch vs algorithmic)",
                 "launches_per_step": r["launches"], "avg_launch_us": 1e3 * r["ms"] / max(1, r["launches"]),
                 "share_of_step": r["ms"] / total_ms, "peak_source": peaks["source"] + " (MEASURED_PEAKS.json)"})
    if beside:
        wall = total_ms - sum(beside.values())
        roof.update({"share_of_step": r["ms"] / wall, "concurrent_classes": beside,
                     "note": "kernel-per-op tail (16 of 256 sessions) timed while it overlaps the cluster launches: share = cluster time / (profiled total - tail)"})
    return roof


def measure(args, precision, streams, K, Wm, rank, local, world, sd, barrier, first_chunk=True):
    """config-1-shaped workload at `streams` concurrent streams in one precision: device-resident value, e2e through
    BatchSynthesizer (host ids in, pinned-host PCM out), p50 first-chunk latency, per-kernel profile of one step."""
    from llmvox_b200.engine import Engine
    from llmvox_b200.streaming import BatchSynthesizer, LaneRunner
    n_groups = Wm + K
    e = Engine(sd, device=local, precision=precision, max_sessions=streams * (n_groups + 1), max_batch=streams,
               max_context=TOKENS + 8, max_vocode_frames=max(64, streams) * 96, decode_lanes=args.lanes)
    runner = LaneRunner(e, args.lanes)
    texts = synthetic_text(streams, 1000 + rank)
    groups = [list(range(g * streams, (g + 1) * streams)) for g in range(n_groups + 1)]
    pcm = torch.empty((streams * TOKENS * 320,), dtype=torch.float32, device=e.device)
    for g in groups[:n_groups]:                       # value: inputs resident before the clock starts
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

    # e2e: host text ids in, PCM out to pinned host memory, through the public batched API
    bs = BatchSynthesizer(e, streams, SCHEDULE[0], stop_on_eoa=False, slots=groups[n_groups], lanes=args.lanes)
    bs_b = BatchSynthesizer(e, streams, SCHEDULE[0], stop_on_eoa=False, slots=groups[0], lanes=args.lanes)
    e2e_steps = max(3, min(K, 10))

    def e2e_step():
        bs.start(texts)                                               # H2D of slots + text ids
        n = 0
        for chunks in bs.run(TOKENS, flush_tail=True, copy=False):     # D2H of every chunk into pinned memory
            n += sum(c.length for c in chunks)
        return n

    def e2e_steps_back_to_back(steps):
        """`steps` batches through the public API the way a server runs them: two synthesizers on alternating slot groups,
        batch i + 1 is started (H2D of its ids, its first decode rounds) as soon as all of batch i's GPU work is enqueued,
        THEN the host waits for batch i's last PCM.  Every batch's inputs still come from the host and every PCM sample
        still lands in pinned host memory inside the timed region."""
        def begin(b):
            b.start(texts)
            g = b.run(TOKENS, flush_tail=True, copy=False, yield_when_enqueued=True)
            n = 0
            for chunks in g:                                           # until everything of this batch is enqueued
                if not chunks:
                    break
                n += sum(c.length for c in chunks)
            return g, n
        total, prev = 0, None
        for i in range(steps):
            cur = begin(bs if i % 2 == 0 else bs_b)
            if prev is not None:
                total += prev[1] + sum(c.length for chunks in prev[0] for c in chunks)
            prev = cur
        total += prev[1] + sum(c.length for chunks in prev[0] for c in chunks)
        return total
    e2e_step()
    assert e2e_steps_back_to_back(4) == 4 * streams * TOKENS      # both synthesizers' pinned rings reach their final size
    fc = {}
    if first_chunk:   # p50 first-chunk latency: host text ids -> the first chunk's PCM of every stream in host memory
        for dump in (10, 160):
            if dump > TOKENS:
                continue
            b2 = BatchSynthesizer(e, streams, dump, stop_on_eoa=False, slots=groups[n_groups], lanes=args.lanes)
            lat = []
            for _ in range(9):
                torch.cuda.synchronize()
                tf = time.perf_counter()
                b2.start(texts)
                gen = b2.run(dump, flush_tail=False, copy=False)
                first = next(gen)
                lat.append(1e3 * (time.perf_counter() - tf))
                assert len(first) == streams
                for _ in gen:
                    pass
            fc[dump] = float(np.median(lat[2:]))
    barrier()
    t0 = time.perf_counter()
    assert e2e_steps_back_to_back(e2e_steps) == e2e_steps * streams * TOKENS
    torch.cuda.synchronize()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    barrier()
    # the same without overlapping consecutive batches (each batch drained before the next starts): reported beside it
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        assert e2e_step() == streams * TOKENS
    torch.cuda.synchronize()
    e2e_serial_ms = 1e3 * (time.perf_counter() - t0)
    barrier()

    # per-kernel profile of one step (events around every launch), beside the timed region
    e.open(groups[0])
    e.feed_text(groups[0], texts)
    runner.sync_from_control()
    torch.cuda.synchronize()
    e.profile(True)
    device_step(e, LaneRunner(e, args.lanes), groups[0], pcm, alone=True)
    prof = e.profile_report()
    e.profile(False)
    n_cluster, n_lanes = runner.plan(streams, None)
    path = []
    if n_cluster:
        w16, w8 = e.cluster_capacity()
        if w8 and n_cluster > w16:
            path.append("%d sessions on the cluster-resident kernel (8-CTA clusters, %d co-resident per wave)" % (n_cluster, w8 // 16))
        else:
            path.append("%d sessions on the cluster-resident kernel (%d clusters of 16 CTAs, waves of <= %d)" % (n_cluster, (n_cluster + 15) // 16, max(1, w16 // 16)))
    if n_lanes:
        path.append("%d sessions on the kernel-per-op chain (%d lanes)" % (n_lanes, args.lanes))
    e.close()
    return {"ms": ms, "e2e_ms": e2e_ms, "e2e_serial_ms": e2e_serial_ms, "e2e_steps": e2e_steps, "first_chunk": fc, "prof": prof, "launches": int(launches), "clocks": clk,
            "decode_path": " + ".join(path)}


def reduce_max(vals, world, device):
    import torch.distributed as dist
    t = torch.tensor(vals, dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def vocoder_bulk(args, local, sd, peaks):
    """BASELINE config 3 at its stated size: 1024 streams x 1280 codes, one chunk each, in launch groups of 48 chunks
    (61,440 frames of workspace): tensor-pipe roofline of the GEMM-bound half."""
    from llmvox_b200.engine import Engine
    n_streams, vb_len, group = (1024, 1280, 48) if not args.short else (96, 1280, 48)
    g = torch.Generator().manual_seed(3)
    vcodes = torch.randint(0, 4096, (n_streams * vb_len,), generator=g).to(torch.device("cuda", local), torch.int32)
    ve = Engine(sd, device=local, precision="bf16", max_sessions=2, max_context=32, max_vocode_frames=group * vb_len + 64)
    vcu = list(range(0, (n_streams + 1) * vb_len, vb_len))
    vout = torch.empty((n_streams * vb_len * 320,), dtype=torch.float32, device=vcodes.device)
    ve.vocode(vcodes[: 2 * group * vb_len], vcu[: 2 * group + 1], out=vout)
    torch.cuda.synchronize()
    v0, v1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 2
    v0.record()
    for _ in range(reps):
        ve.vocode(vcodes, vcu, out=vout)
    v1.record()
    torch.cuda.synchronize()
    vms = v0.elapsed_time(v1) / reps
    vfl = n_streams * vb_len * 132.78e6          # algorithmic FLOPs per frame at L = 1280
    ve.profile(2)
    ve.vocode(vcodes[: group * vb_len], vcu[: group + 1], out=vout)
    vrep = ve.profile_report()
    ve.profile(False)
    ve.close()
    return {"workload": f"config3: {n_streams} streams x {vb_len} codes, one chunk each, launch groups of {group} chunks", "ms": vms,
            "audio_s_per_s": n_streams * vb_len / CODES_PER_SEC / (vms / 1e3), "tflops": vfl / (vms / 1e3) / 1e12,
            "frac_of_sustained_peak": vfl / (vms / 1e3) / 1e12 / peaks["bf16_tflops_sustained"],
            "gemm_tflops_algorithmic_one_group": {k: round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1) for k, v in vrep.items() if v["flops"] > 0},
            "share_ms_one_group": {k: round(v["ms"], 3) for k, v in sorted(vrep.items(), key=lambda kv: -kv[1]["ms"])}}


def config4(args, rank, local, world, sd, barrier):
    """BASELINE config 4: 4096 concurrent sessions x 600 codes, session-sharded over the ranks (sharding.py; no collective
    on the data path), end to end: host text ids -> PCM in pinned host memory through BatchSynthesizer."""
    from llmvox_b200.engine import Engine
    from llmvox_b200.sharding import shard_indices
    from llmvox_b200.streaming import BatchSynthesizer
    sessions, tokens = (4096, 600) if not args.short else (512, 120)
    mine = shard_indices(sessions, world, rank)
    e = Engine(W_for(tokens + 8), device=local, precision="bf16", max_sessions=len(mine), max_batch=len(mine), max_context=tokens + 8,
               max_vocode_frames=65536, decode_lanes=1)
    rng = np.random.RandomState(rank)
    texts = [rng.randint(3, 259, size=tokens).tolist() for _ in mine]

    def once():
        bs = BatchSynthesizer(e, len(mine), 10, stop_on_eoa=False, lanes=1)
        bs.start(texts)
        return sum(c.length for chunks in bs.run(tokens, flush_tail=True, copy=False) for c in chunks)
    once()
    barrier()
    t0 = time.perf_counter()
    n = once()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    assert n == len(mine) * tokens
    e.profile(True)
    once()
    prof = e.profile_report()
    e.profile(False)
    mem = e.device_bytes
    e.close()
    dt = reduce_max([dt], world, torch.device("cuda", local))[0]
    audio = sessions * tokens / CODES_PER_SEC
    by = decode_bytes(len(mine), tokens, 2)
    top = sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:4]
    return {"workload": f"config4: {sessions} sessions x {tokens} codes over {world} GPU(s) ({len(mine)} per rank, all concurrent), text ids -> pinned-host PCM",
            "scaling": "strong", "wall_s": dt, "value": audio / dt, "unit": "audio-s/s", "per_gpu": audio / dt / world,
            "roofline": {"bound": "hbm", "algorithmic_bytes_per_rank": by, "achieved": by / dt / 1e9, "peak": measured_peaks()["hbm_gbs"], "unit": "GB/s",
                         "frac": by / dt / 1e9 / measured_peaks()["hbm_gbs"],
                         "note": "KV-bound decode bytes of one rank / e2e wall time (vocoder + host feeding inside the wall time)"},
            "limiter": {k: round(v["ms"], 1) for k, v in top}, "device_gib": mem / 2 ** 30}


_W_CACHE = {}


def W_for(rows):
    from llmvox_b200 import weights as W
    if rows not in _W_CACHE:
        _W_CACHE[rows] = W.make_random_weights(SEED, wpe_rows=rows)
    return _W_CACHE[rows]


def run_gpu(args):
    import torch.distributed as dist
    from llmvox_b200 import build as B

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B.build()
    K, Wm = args.steps, max(args.warmup, 3)
    sd = W_for(256)
    peaks = measured_peaks()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- headline: BASELINE config 1 in the precision whose greedy tokens are the reference's (exact), bf16 beside it
    main = measure(args, args.precision, STREAMS, K, Wm, rank, local, world, sd, barrier)
    other = "bf16" if args.precision == "exact" else "exact"
    alt = measure(args, other, STREAMS, K, Wm, rank, local, world, sd, barrier, first_chunk=False) if not args.no_alt else None
    # ---- north-star operating point: 256 concurrent streams per GPU
    K256 = max(2, min(K, 5))
    s256 = measure(args, "bf16", 256, K256, 3, rank, local, world, sd, barrier) if (STREAMS != 256 and not args.no_256) else None
    # the same point and one full wave of 8-CTA clusters (240 streams) in the exact precision / bf16: value + e2e only
    x256 = measure(args, "exact", 256, K256, 3, rank, local, world, sd, barrier, first_chunk=False) if s256 else None
    s240 = measure(args, "bf16", 240, K256, 3, rank, local, world, sd, barrier, first_chunk=False) if s256 else None
    voc = vocoder_bulk(args, local, sd, peaks) if (rank == 0 and not args.no_vocoder) else None
    barrier()
    c4 = config4(args, rank, local, world, sd, barrier) if (world > 1 or args.config4) else None

    def line_of(m, streams, k):
        ms, e2e_ms, e2e_serial_ms = reduce_max([m["ms"], m["e2e_ms"], m["e2e_serial_ms"]], world, dev)
        audio = world * streams * TOKENS / CODES_PER_SEC
        return {"value": audio * k / (ms / 1e3), "ms_per_step": ms / k, "e2e_value": audio * m["e2e_steps"] / (e2e_ms / 1e3),
                "e2e_drained": audio * m["e2e_steps"] / (e2e_serial_ms / 1e3)}
    lm = line_of(main, STREAMS, K)
    la = line_of(alt, STREAMS, K) if alt else None
    l256 = line_of(s256, 256, K256) if s256 else None
    lx256 = line_of(x256, 256, K256) if x256 else None
    l240 = line_of(s240, 240, K256) if s240 else None

    if rank == 0:
        threads = os.cpu_count() or 1
        cpu_streams = 6
        cpu_sample(1, threads)
        ca, cdt = cpu_sample(cpu_streams, threads)
        kvb = 4 if args.precision == "exact" else 2
        by = decode_bytes(STREAMS, TOKENS, kvb)
        ms_step = lm["ms_per_step"]
        precision_note = {"exact": "exact: bf16 LayerNorm-folded weights, bf16 hi|lo activation pairs (fp32-class), fp32 accumulate, fp32 KV cache; "
                                   "greedy tokens identical to the fp32 reference loop on the rounded weights (tests/test_gpu_exact.py)",
                          "bf16": "bf16 weights, activations and KV cache, fp32 accumulate (teacher-forced logits within 2e-2)",
                          "fp32": "fp32 FMA-pipe parity mode"}
        line = {"metric": "audio-sec generated/sec", "value": lm["value"], "unit": "audio-s/s", "n_gpus": world, "steps": K,
                "warmup": Wm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision != "fp32" else "f32", "data": "synthetic",
                "config": {"workload": f"config1: {STREAMS} streams x {TOKENS} codes per GPU, KV-cached greedy decode + chunked vocoder ({'/'.join(map(str, SCHEDULE))})",
                           "weights": "random-init english-tiny GPT + frame75 WavTokenizer decoder, seed 1234",
                           "precision": args.precision, "precision_note": precision_note[args.precision],
                           "decode_lanes": args.lanes, "decode_path": main["decode_path"],
                           "l2": "no flush: per-step working set (189 MB bf16 weights + KV + activations) exceeds the 126 MB L2"},
                "x_realtime_per_gpu": lm["value"] / world,
                # whole-step view: bf16 weights once per iteration + KV read/append, against measured HBM peak
                "roofline_step": {"bound": "hbm", "algorithmic_bytes": by, "achieved": by / (ms_step / 1e3) / 1e9, "peak": peaks["hbm_gbs"],
                                  "unit": "GB/s", "frac": by / (ms_step / 1e3) / 1e9 / peaks["hbm_gbs"],
                                  "note": "decode bytes of one step per GPU / step time; the step is latency-bound (DESIGN.md 5)"},
                "roofline": roofline_of(main["prof"], peaks),
                "kernels": {k: {"launches": v["launches"], "ms": round(v["ms"], 4)} for k, v in sorted(main["prof"].items(), key=lambda kv: -kv[1]["ms"])},
                "cpu_baseline": {"value": ca / cdt, "unit": "audio-s/s", "cores": threads, "kind": "port",
                                 "sample": f"{cpu_streams} of the {STREAMS} streams, sequential batch-1: {TOKENS} codes + chunks {SCHEDULE} each"},
                "e2e": {"value": lm["e2e_value"], "unit": "audio-s/s", "steps": main["e2e_steps"],
                        "h2d_bytes_per_step": STREAMS * TOKENS * 4 + (STREAMS + 1) * 4 + STREAMS * 4,
                        "d2h_bytes_per_step": STREAMS * TOKENS * 320 * 4,
                        "what": "BatchSynthesizer from host text ids to PCM in pinned host memory, consecutive batches back to back on two "
                                "alternating slot groups (batch i + 1 starts once batch i's GPU work is enqueued, as a server would); "
                                "value_drained = every batch drained on the host before the next starts",
                        "value_drained": lm["e2e_drained"]},
                "first_chunk_latency": {"p50_ms": main["first_chunk"].get(10), "p50_ms_160": main["first_chunk"].get(160), "streams": STREAMS,
                                        "codes": [10, 160],
                                        "what": "host text ids -> PCM of every stream's first chunk (replica 0: 10 codes, replica 1: 160) in pinned host memory"},
                "gpu_launches": main["launches"], "clocks": main["clocks"]}
        if alt:
            line["other_precision"] = {"precision": other, "precision_note": precision_note[other], "value": la["value"], "ms_per_step": la["ms_per_step"],
                                       "e2e": la["e2e_value"], "unit": "audio-s/s", "roofline": roofline_of(alt["prof"], peaks),
                                       "gpu_launches": alt["launches"]}
        if s256:
            by256 = decode_bytes(256, TOKENS, 2)
            line["streams256"] = {"workload": f"256 streams x {TOKENS} codes per GPU + chunked vocoder ({'/'.join(map(str, SCHEDULE))}), bf16", "steps": K256,
                                  "value": l256["value"], "e2e": l256["e2e_value"], "unit": "audio-s/s", "ms_per_step": l256["ms_per_step"],
                                  "x_realtime_per_gpu": l256["value"] / world,
                                  "first_chunk_p50_ms": {"10": s256["first_chunk"].get(10), "160": s256["first_chunk"].get(160)},
                                  "decode_path": s256["decode_path"], "roofline": roofline_of(s256["prof"], peaks),
                                  "roofline_step": {"algorithmic_bytes": by256, "achieved": by256 / (l256["ms_per_step"] / 1e3) / 1e9,
                                                    "frac": by256 / (l256["ms_per_step"] / 1e3) / 1e9 / peaks["hbm_gbs"], "unit": "GB/s"},
                                  "kernels": {k: {"launches": v["launches"], "ms": round(v["ms"], 3)}
                                              for k, v in sorted(s256["prof"].items(), key=lambda kv: -kv[1]["ms"])[:6]},
                                  "gpu_launches": s256["launches"],
                                  "exact": {"value": lx256["value"], "e2e": lx256["e2e_value"], "ms_per_step": lx256["ms_per_step"],
                                            "decode_path": x256["decode_path"], "roofline": roofline_of(x256["prof"], peaks)},
                                  "streams240_bf16": {"value": l240["value"], "e2e": l240["e2e_value"], "ms_per_step": l240["ms_per_step"],
                                                      "decode_path": s240["decode_path"], "roofline": roofline_of(s240["prof"], peaks),
                                                      "what": "one full wave of 8-CTA clusters (15 x 16 sessions), no kernel-per-op tail"}}
        line["vocoder_bulk"] = voc
        if c4:
            line["config4"] = c4
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="llmvox_b200", choices=["llmvox_b200", "reference"])
    ap.add_argument("--precision", default="exact", choices=["exact", "bf16", "fp32"],
                    help="headline precision: exact = token-identical tensor-core decode (default); the other of exact / bf16 is reported beside it")
    ap.add_argument("--no-alt", action="store_true", help="skip the other-precision leg")
    ap.add_argument("--no-256", action="store_true", help="skip the 256-stream leg")
    ap.add_argument("--no-vocoder", action="store_true", help="skip the config-3 vocoder bulk leg")
    ap.add_argument("--config4", action="store_true", help="run the config-4 leg on one GPU too (it always runs under torchrun)")
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
