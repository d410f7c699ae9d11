"""Batched generate / stream-chunk host loop on top of the engine.

The reference runs one Python thread per request per replica (streaming_server.py:250-426), decoding one code
per iteration with a device->host sync for every token and vocoding each ready chunk alone.  Here many sessions
advance together: the engine decodes `k` steps for all of them without touching the host (`k` = steps until the
next session can possibly emit), the per-session ChunkScheduler (scheduler.py) decides which code ranges became
chunks, and all ready chunks of a round are vocoded as ONE ragged batch -- each chunk still an independent decode,
exactly as the reference.  PCM leaves through a pinned host buffer with an asynchronous copy."""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .engine import Engine, Sampling
from .scheduler import ChunkScheduler, INITIAL_DUMP_SIZE_1, MAX_DUMP_SIZE


def os_environ_no8() -> bool:
    """LLMVOX_B200_CD_NO8=1: measurement knob of the engine (cluster_launch), mirrored here so that plan() agrees."""
    import os
    return os.environ.get("LLMVOX_B200_CD_NO8") == "1"


class LaneRunner:
    """Runs the decode iterations of a batch off the control stream, so that the control stream's work (gather / vocode /
    copies) overlaps with the next iterations.  The decode path is chosen PER CALL (lvx_decode_steps_ex), never through
    engine state:

    * cluster-resident decode kernel (bf16 / exact precision): ONE call on the side stream; the engine cuts it into
      balanced waves of co-resident clusters, which run back to back: 7 x 16 sessions per wave of 16-CTA clusters
      (~150 us per iteration however many of the clusters are used), and for greedy batches above that 15 x 16
      sessions per wave of 8-CTA clusters (each CTA streams twice the weights: a longer iteration, but 240 sessions
      abreast).
    * kernel-per-op chain (fp32 precision, sampled decoding, and the part of a batch the cluster waves do not take):
      `lanes` independent groups on their own streams.  A decode iteration there is a chain of ~35 dependent,
      latency-bound kernels that leaves most of the GPU idle; sessions never interact, so disjoint groups advance
      concurrently (engine decode lanes).
    * a call may also be split between both (`plan` returns the two counts and `launch` runs them at the same time on
      different streams): a greedy batch slightly above one wave of 8-CTA clusters (256 streams = 240 + 16) keeps the
      wave on the cluster kernel and runs the tail on ONE kernel-per-op lane (LVX_PATH_PER_OP_TAIL: grids sized for the 28
      SMs the wave leaves free) while decode rounds and vocoder batches take turns on the GPU.  Measured (bench.py --streams
      256, bf16, ms per step): all on the lanes 86.8, two waves 98.6, 240 + 16 on three tail lanes with the vocoder
      overlapped 92.4, one tail lane overlapped 72.4, one tail lane taking turns 69.3.  (With 16-CTA clusters only the split is off by default: 224 + 32 ran at 6858 audio-s/s against 8251 on the lanes.)

    `launch()` enqueues on the side streams and returns the completion events; `join()` makes the control stream wait
    for them.  Every launch first waits, on every stream it uses, for the previous round's events of the OTHER streams:
    a session may move between streams from round to round (the active set shrinks, the path changes with the batch
    size), and its context length / KV pages must be ordered across that move."""

    # 16-CTA clusters only (LLMVOX_B200_CD_NO8=1, or a device without room for 8-CTA clusters): two waves of 7; measured (bench.py --short,
    # audio-s/s, cluster vs kernel-per-op): 112: 8341 / 5418, 128: 5780 / 5992, 192: 8300 / 7624, 224: 9062 / 7457,
    # 256: 7647 / 8251.  Just above one full wave (113..139 sessions) the second wave is nearly empty and the kernel-per-op
    # lanes win.
    CLUSTER_DECODE_MAX_BATCH = 224
    CLUSTER_DECODE_GAP = (113, 139)
    HYBRID_ABOVE_MAX_BATCH = False
    # greedy decoding: above one wave of 16-CTA clusters the engine switches to 8-CTA clusters, 15 x 16 = 240 sessions per wave
    # (engine.cluster_capacity()).  A batch slightly above a wave (256 streams) runs the wave on the cluster kernel and the
    # rest on the kernel-per-op lanes AT THE SAME TIME (the 28 SMs fifteen 8-CTA clusters leave idle); larger batches take
    # further waves, and beyond CLUSTER8_MAX_WAVES waves the kernel-per-op chain's large-M GEMMs win (config 4).
    HYBRID_TAIL_MAX = 32
    CLUSTER8_MAX_WAVES = 2

    def __init__(self, engine: Engine, lanes: Optional[int] = None):
        import os
        self.e = engine
        self.G = max(1, min(engine.decode_lanes, lanes or engine.decode_lanes))
        # (default priority: with high-priority decode streams the control stream's vocoder only runs in the gaps, which
        # costs config 4 -- 2048 sessions per GPU, every decode kernel fills the GPU -- 18 % of its rate: measured)
        prio = int(os.environ.get("LLMVOX_B200_LANE_PRIORITY", "0"))
        self.streams = [torch.cuda.Stream(device=engine.device, priority=prio) for _ in range(self.G)]
        self.side = self.streams[0]                                   # cluster-kernel launches
        self.cluster_default = os.environ.get("LLMVOX_B200_CLUSTER", "1") != "0"
        if "LLMVOX_B200_CLUSTER_MAX_BATCH" in os.environ:      # measurement overrides of the thresholds above
            self.CLUSTER_DECODE_MAX_BATCH = int(os.environ["LLMVOX_B200_CLUSTER_MAX_BATCH"])
        if "LLMVOX_B200_HYBRID_TAIL" in os.environ:
            self.HYBRID_TAIL_MAX = int(os.environ["LLMVOX_B200_HYBRID_TAIL"])
        self._prev: List[Tuple[torch.cuda.Stream, torch.cuda.Event]] = []

    def split(self, slots: Sequence[int], groups: Optional[int] = None) -> List[List[int]]:
        n, G = len(slots), min(groups or self.G, len(slots))
        base, rem = divmod(n, G)
        out, pos = [], 0
        for g in range(G):
            k = base + (1 if g < rem else 0)
            out.append(list(slots[pos:pos + k]))
            pos += k
        return out

    def sync_from_control(self):
        """The side streams wait for everything enqueued so far on the current (control) stream."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.e.device))
        for st in self.streams:
            st.wait_event(ev)

    def wave8(self, sampling: Optional[Sampling]) -> int:
        """Sessions per wave of the 8-CTA cut for this sampler (0: the cut does not apply)."""
        if self.e.precision not in ("bf16", "exact"):
            return 0
        if not hasattr(self, "_caps"):
            self._caps = self.e.cluster_capacity() if hasattr(self.e, "cluster_capacity") else (112, 240)
        return 0 if os_environ_no8() else self._caps[1]

    def plan(self, n: int, sampling: Optional[Sampling]) -> Tuple[int, int]:
        """-> (sessions for the cluster-resident kernel, sessions for the kernel-per-op lanes)."""
        if not self.cluster_default or not self.e.cluster_decode_applicable(sampling):
            return 0, n
        w8 = self.wave8(sampling)
        if w8:
            if n <= w8:
                return n, 0
            if n - w8 <= self.HYBRID_TAIL_MAX and self.G > 1:
                return w8, n - w8
            return (n, 0) if n <= self.CLUSTER8_MAX_WAVES * w8 else (0, n)
        if n <= self.CLUSTER_DECODE_MAX_BATCH:
            gap = self.CLUSTER_DECODE_GAP[0] <= n <= self.CLUSTER_DECODE_GAP[1] and self.e.precision == "bf16"
            return (0, n) if gap else (n, 0)
        if self.HYBRID_ABOVE_MAX_BATCH:
            return self.CLUSTER_DECODE_MAX_BATCH, n - self.CLUSTER_DECODE_MAX_BATCH
        return 0, n

    def _cluster_call(self, n: int, sampling: Optional[Sampling]) -> bool:
        return self.plan(n, sampling)[0] > 0

    def launch(self, slots: Sequence[int], n_steps: int, sampling: Optional[Sampling] = None) -> List[torch.cuda.Event]:
        """Enqueues n_steps iterations for the batch on the side streams; returns their completion events."""
        from ._lib import PATH_CLUSTER, PATH_PER_OP, PATH_PER_OP_TAIL
        n_cluster, n_lanes = self.plan(len(slots), sampling)
        if n_cluster and n_lanes and os.environ.get("LLMVOX_B200_SERIAL", "1") == "1":
            # a full wave of 8-CTA clusters holds 120 of the 148 SMs: vocoder kernels squeezed into the rest run ~5x
            # slower and starve the kernel-per-op tail (measured at 256 streams: 72.9 ms per step overlapped, 69.3 taking
            # turns), so with a tail the decode rounds and the control stream's vocoder batches take turns on the whole
            # GPU; without one (<= 240 streams) overlapping stays better (57.7 vs 65.5 ms at 240)
            self.sync_from_control()
        calls = []                                                   # (stream, lane, slots, path)
        if n_cluster:
            calls.append((self.side, 0, list(slots[:n_cluster]), PATH_CLUSTER))
        if n_lanes:
            lane0 = 1 if (n_cluster and self.G > 1) else 0            # lane 0's workspace belongs to the cluster call
            tail_lanes = max(1, self.G - lane0)
            if n_cluster:
                tail_lanes = min(tail_lanes, int(os.environ.get("LLMVOX_B200_TAIL_LANES", "1")))
            groups = self.split(slots[n_cluster:], tail_lanes)
            for g, grp in enumerate(groups):
                calls.append((self.streams[lane0 + g], lane0 + g, grp, PATH_PER_OP_TAIL if n_cluster else PATH_PER_OP))
        used = {id(st): st for st, _, _, _ in calls}
        for st in used.values():
            for pst, pev in self._prev:
                if pst is not st:
                    st.wait_event(pev)
        out = []
        for st, lane, grp, path in calls:
            self.e.decode_steps(grp, n_steps, sampling, stream=st, lane=lane, path=path)
        for st in used.values():
            ev = torch.cuda.Event()
            ev.record(st)
            out.append((st, ev))
        self._prev = out
        return [ev for _, ev in out]

    def join(self, events: Sequence[torch.cuda.Event]):
        main = torch.cuda.current_stream(self.e.device)
        for ev in events:
            main.wait_event(ev)

    def decode(self, slots: Sequence[int], n_steps: int, sampling: Optional[Sampling] = None):
        """Enqueues n_steps iterations for the batch; the control stream then waits for them."""
        self.join(self.launch(slots, n_steps, sampling))


@dataclass
class Chunk:
    session: int          # index into the batch
    start: int            # first code of the chunk in the sentence
    length: int           # codes
    pcm: np.ndarray       # float32 mono 24 kHz, 320 * length samples (the reference's `.tobytes()` payload, :368)

    def tobytes(self) -> bytes:
        return self.pcm.astype("float32", copy=False).tobytes()


class ChunkEmitter:
    """Ready code ranges -> ONE ragged vocoder batch -> pinned host memory (streaming_server.py:359-368 for many chunks at
    once: gather the ranges from the sessions' device-side code histories, vocode every range as an independent chunk,
    copy the PCM out asynchronously).  `enqueue` only enqueues on the current (control) stream and returns a ticket;
    `finish` waits for that ticket's copy and returns one float32 array per range, in the order given."""

    def __init__(self, engine: Engine, bandwidth_id: int = 0, buffers: int = 3):
        self.e = engine
        self.bw = bandwidth_id
        self._ring: List[Optional[torch.Tensor]] = [None] * buffers
        self._next = 0
        self._cap = 1 << 20

    def _pinned(self, n: int) -> torch.Tensor:
        """Pinned buffers used round-robin: a ticket's PCM (and, with copy=False, the arrays handed out) stays valid until
        `buffers - 1` further tickets have been enqueued.  Every buffer is kept at the size of the LARGEST ticket seen so
        far (plus a quarter): allocating pinned memory synchronises the device and costs milliseconds, so the ring must stop
        growing after the first batch instead of reallocating a slot whenever a bigger round happens to land on it."""
        i = self._next
        self._next = (i + 1) % len(self._ring)
        if n > self._cap:
            self._cap = max(n + n // 4, 1 << 20)
        if self._ring[i] is None or self._ring[i].numel() < self._cap:
            self._ring[i] = torch.empty((self._cap,), dtype=torch.float32, pin_memory=True)
        return self._ring[i]

    def enqueue(self, ready: Sequence[Tuple[int, int, int]]):
        """ready: (slot, start, count) ranges."""
        if not ready:
            return None
        slots = [r[0] for r in ready]
        # a session may own several ready ranges in one round: the engine wants distinct slots per call
        parts, order = [], []
        remaining = list(range(len(ready)))
        while remaining:
            seen, take, rest = set(), [], []
            for j in remaining:
                (rest if slots[j] in seen else take).append(j)
                seen.add(slots[j])
            parts.append(self.e.gather_code_ranges([slots[j] for j in take], [ready[j][1] for j in take], [ready[j][2] for j in take]))
            order.extend(take)
            remaining = rest
        codes = torch.cat(parts) if len(parts) > 1 else parts[0]
        cu = [0]
        for j in order:
            cu.append(cu[-1] + ready[j][2])
        pcm = self.e.vocode(codes, cu, self.bw)
        host = self._pinned(pcm.numel())
        host[: pcm.numel()].copy_(pcm, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.e.device))
        return (order, cu, host, ev, pcm, len(ready))

    def done(self, ticket) -> bool:
        return ticket is None or ticket[3].query()

    def finish(self, ticket, copy: bool = True) -> List[np.ndarray]:
        if ticket is None:
            return []
        order, cu, host, ev, _pcm, n = ticket
        ev.synchronize()
        hop = self.e.cfg.hop
        arr = host.numpy()
        out: List[Optional[np.ndarray]] = [None] * n
        for k, j in enumerate(order):
            seg = arr[cu[k] * hop: cu[k + 1] * hop]
            out[j] = seg.copy() if copy else seg
        return out


class BatchSynthesizer:
    """One sentence per session, all sessions stepped together."""

    def __init__(self, engine: Engine, n_sessions: int, initial_dump_size: int = INITIAL_DUMP_SIZE_1,
                 max_dump_size: int = MAX_DUMP_SIZE, stop_on_eoa: bool = True, sampling: Optional[Sampling] = None,
                 slots: Optional[Sequence[int]] = None, bandwidth_id: int = 0, lanes: Optional[int] = None,
                 max_audio_length: Optional[int] = None):
        self.e = engine
        self.n = n_sessions
        self.slots = list(slots) if slots is not None else list(range(n_sessions))
        assert len(self.slots) == n_sessions
        kw = {} if max_audio_length is None else {"max_audio_len": max_audio_length}
        self.sched = [ChunkScheduler(dump_size=initial_dump_size, max_dump=max_dump_size, stop_on_eoa=stop_on_eoa,
                                     eoa=engine.cfg.eoa_token_id, **kw) for _ in range(n_sessions)]
        self.initial_dump_size = initial_dump_size
        self.stop_on_eoa = stop_on_eoa
        self.sampling = sampling or Sampling()
        self.bw = bandwidth_id
        self.steps_done = 0
        self.runner = LaneRunner(engine, lanes)
        self.emitter = ChunkEmitter(engine, bandwidth_id)
        self._prog = None

    def start(self, text_ids: Sequence[Union[Sequence[int], str]], keep_schedule: bool = False, clean: bool = False):
        """Opens the sessions (the per-sentence reset of :404-416) and hands them their text: lists of ids, or raw
        sentences (str) that the engine tokenises on the device (`clean=True` runs clean_text there first).  A new call is
        a new request (fresh generator threads in the reference, so the dump size starts at its initial value again);
        `keep_schedule=True` continues the same request, where the dump size only ever grows (:373-375)."""
        assert len(text_ids) == self.n
        self.e.open(self.slots)
        if self.n and all(isinstance(t, str) for t in text_ids):
            self.e.feed_sentences(self.slots, text_ids, clean=clean)
        else:
            self.e.feed_text(self.slots, text_ids)
        for s in self.sched:
            s.new_sentence()
            if not keep_schedule:
                s.dump_size = self.initial_dump_size
        self.steps_done = 0
        self.runner.sync_from_control()

    def _enqueue_emit(self, ready: List[Tuple[int, int, int]]):
        """ready: (session index, start, length).  Enqueues one ragged vocoder batch + one async D2H into pinned
        memory on the control stream; returns a ticket for `_finish_emit`."""
        if not ready:
            return None
        return (ready, self.emitter.enqueue([(self.slots[i], s, c) for i, s, c in ready]))

    def _finish_emit(self, ticket, copy: bool = True) -> List[Chunk]:
        if ticket is None:
            return []
        ready, t = ticket
        out = [Chunk(i, s, c, pcm) for (i, s, c), pcm in zip(ready, self.emitter.finish(t, copy))]
        out.sort(key=lambda ch: (ch.session, ch.start))
        return out

    def run(self, max_steps: int, flush_tail: bool = False, copy: bool = True, yield_when_enqueued: bool = False) -> Iterator[List[Chunk]]:
        """Decodes up to `max_steps` codes per session, yielding the chunks of each round as they are ready.

        `yield_when_enqueued`: yields one EMPTY list at the point where all of this batch's GPU work has been enqueued and
        only host waits remain, so that a caller serving several batches can start the next one (another BatchSynthesizer on
        other slots) before it blocks on this batch's last PCM -- the next batch's first decode rounds then overlap this
        batch's last vocoder batch and copy on the GPU.

        Software-pipelined: round r+1's decode iterations are enqueued on the lanes BEFORE the host waits for round
        r's PCM, so the vocoder and the copies of round r overlap the decode of round r+1.  With stop_on_eoa the
        codes of round r are read back first (they decide which sessions are still active); codes a session decodes
        past its EOA are discarded, as the reference resets there."""
        active = list(range(self.n))
        pending = None                      # ticket of the previous round
        while active and self.steps_done < max_steps:
            k = min(min(self.sched[i].steps_to_next_event() for i in active), max_steps - self.steps_done)
            slots = [self.slots[i] for i in active]
            self.runner.decode(slots, k, self.sampling)
            if pending is not None:         # round r-1's PCM is surely done by now or soon: hand it out
                chunks = self._finish_emit(pending, copy)
                pending = None
                if chunks:
                    yield chunks
            eoa_at = None
            if self.stop_on_eoa:   # the end of a sentence is detected on the device (SessionState.eoa_pos): no code value
                if self._prog is None or self._prog.numel() < 2 * self.n:      # visits the host, one small pinned copy does
                    self._prog = torch.empty((2 * self.n,), dtype=torch.int32, pin_memory=True)
                self.e.session_progress(slots, self._prog)
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.e.device))
                ev.synchronize()
                eoa_at = self._prog[: 2 * len(slots)].view(-1, 2)[:, 0].tolist()
            ready: List[Tuple[int, int, int]] = []
            for a, i in enumerate(active):
                sc = self.sched[i]
                if eoa_at is None or not (self.steps_done <= eoa_at[a] < self.steps_done + k):
                    ready.extend((i, s, c) for (s, c) in sc.push_many(k))      # no EOA in this round: O(emissions)
                    continue
                for t in range(k):
                    if sc.done:
                        break
                    code = sc.eoa if eoa_at[a] == self.steps_done + t else None
                    for (s, c) in sc.push(code):
                        ready.append((i, s, c))
            self.steps_done += k
            pending = self._enqueue_emit(ready)
            active = [i for i in active if not self.sched[i].done]
        tail = None
        if flush_tail:
            ready = []
            for i in range(self.n):
                if not self.sched[i].done:
                    ready.extend((i, s, c) for (s, c) in self.sched[i].flush())
            tail = self._enqueue_emit(ready)
        if yield_when_enqueued:
            yield []
        for ticket in (pending, tail):
            if ticket is not None:
                chunks = self._finish_emit(ticket, copy)
                if chunks:
                    yield chunks

    def codes(self, count: Optional[int] = None) -> np.ndarray:
        """(n_sessions, count) codes decoded so far (host copy)."""
        count = self.steps_done if count is None else count
        return self.e.gather_codes(self.slots, 0, count).cpu().numpy()


def synthesize(engine: Engine, text_ids: Sequence[Sequence[int]], max_steps: int, initial_dump_size: int = INITIAL_DUMP_SIZE_1,
               stop_on_eoa: bool = True, flush_tail: bool = True, sampling: Optional[Sampling] = None,
               bandwidth_id: int = 0, lanes: Optional[int] = None) -> Tuple[np.ndarray, List[List[Chunk]]]:
    """Text ids -> (codes, per-session chunk lists)."""
    bs = BatchSynthesizer(engine, len(text_ids), initial_dump_size, stop_on_eoa=stop_on_eoa, sampling=sampling,
                          bandwidth_id=bandwidth_id, lanes=lanes)
    bs.start(text_ids)
    per: List[List[Chunk]] = [[] for _ in text_ids]
    for chunks in bs.run(max_steps, flush_tail=flush_tail):
        for ch in chunks:
            per[ch.session].append(ch)
    return bs.codes(), per
