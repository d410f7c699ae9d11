"""Batched generate / stream-chunk host loop on top of the engine.

The reference runs one Python thread per request per replica (streaming_server.py:250-426), decoding one code
per iteration with a device->host sync for every token and vocoding each ready chunk alone.  Here many sessions
advance together: the engine decodes `k` steps for all of them without touching the host (`k` = steps until the
next session can possibly emit), the per-session ChunkScheduler (scheduler.py) decides which code ranges became
chunks, and all ready chunks of a round are vocoded as ONE ragged batch -- each chunk still an independent decode,
exactly as the reference.  PCM leaves through a pinned host buffer with an asynchronous copy."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .engine import Engine, Sampling
from .scheduler import ChunkScheduler, INITIAL_DUMP_SIZE_1, MAX_DUMP_SIZE


class LaneRunner:
    """Runs the decode iterations of a batch off the control stream, so that the control stream's work (gather / vocode /
    copies) overlaps with the next iterations.  Two shapes, chosen per call:

    * cluster-resident decode kernel (bf16, greedy, batches up to CLUSTER_DECODE_MAX_BATCH): ONE call for all sessions on
      one side stream; the engine cuts it into launches of at most 7 clusters (7 x 16 sessions are co-resident on a
      B200), which run back to back.  A wave of up to 112 sessions costs the same ~150 us per iteration however many of
      its 7 clusters are used, so batches up to 224 sessions (two waves) beat the kernel-per-op chain.
    * kernel-per-op chain (everything else): `lanes` independent groups on their own streams.  A decode iteration there
      is a chain of ~35 dependent, latency-bound kernels that leaves most of the GPU idle; sessions never interact, so
      disjoint groups advance concurrently (engine decode lanes).

    The side streams only ever wait for the event recorded by `sync_from_control()` (after open / feed), never for
    vocoder work."""

    # above this many sessions in one batch the kernel-per-op chain out-runs the cluster-resident kernel (three waves);
    # measured (bench.py --short, audio-s/s, cluster vs kernel-per-op): 112: 8341 / 5418, 128: 5780 / 5992, 192: 8300 / 7624,
    # 224: 9062 / 7457, 256: 7647 / 8251.  Just above one full wave (113..139 sessions) the second wave is nearly empty.
    CLUSTER_DECODE_MAX_BATCH = 224
    CLUSTER_DECODE_GAP = (113, 139)

    def __init__(self, engine: Engine, lanes: Optional[int] = None):
        import os
        self.e = engine
        self.G = max(1, min(engine.decode_lanes, lanes or engine.decode_lanes))
        self.streams = [torch.cuda.Stream(device=engine.device) for _ in range(self.G)] if self.G > 1 else [None]
        self.side = self.streams[0] if self.G > 1 else torch.cuda.Stream(device=engine.device)   # cluster-kernel launches
        self.cluster_default = os.environ.get("LLMVOX_B200_CLUSTER", "1") != "0"
        if "LLMVOX_B200_CLUSTER_MAX_BATCH" in os.environ:      # measurement override of the threshold above
            self.CLUSTER_DECODE_MAX_BATCH = int(os.environ["LLMVOX_B200_CLUSTER_MAX_BATCH"])

    def split(self, slots: Sequence[int]) -> List[List[int]]:
        n, G = len(slots), min(self.G, len(slots))
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
        for st in set(x for x in self.streams + [self.side] if x is not None):
            st.wait_event(ev)

    def _cluster_call(self, n: int, sampling: Optional[Sampling]) -> bool:
        if not self.cluster_default or n > self.CLUSTER_DECODE_MAX_BATCH or not self.e.cluster_decode_applicable(sampling):
            return False
        return not (self.CLUSTER_DECODE_GAP[0] <= n <= self.CLUSTER_DECODE_GAP[1]) or self.CLUSTER_DECODE_MAX_BATCH > 100000

    def decode(self, slots: Sequence[int], n_steps: int, sampling: Optional[Sampling] = None):
        """Enqueues n_steps iterations for the batch; the control stream then waits for them."""
        main = torch.cuda.current_stream(self.e.device)
        if self._cluster_call(len(slots), sampling):
            self.e.set_cluster_decode(True)
            self.e.decode_steps(slots, n_steps, sampling, stream=self.side, lane=0)
            ev = torch.cuda.Event()
            ev.record(self.side)
            main.wait_event(ev)
            return
        self.e.set_cluster_decode(False)
        if self.G == 1:
            self.e.decode_steps(slots, n_steps, sampling)
            return
        for g, grp in enumerate(self.split(slots)):
            self.e.decode_steps(grp, n_steps, sampling, stream=self.streams[g], lane=g)
            ev = torch.cuda.Event()
            ev.record(self.streams[g])
            main.wait_event(ev)


@dataclass
class Chunk:
    session: int          # index into the batch
    start: int            # first code of the chunk in the sentence
    length: int           # codes
    pcm: np.ndarray       # float32 mono 24 kHz, 320 * length samples (the reference's `.tobytes()` payload, :368)

    def tobytes(self) -> bytes:
        return self.pcm.astype("float32", copy=False).tobytes()


class BatchSynthesizer:
    """One sentence per session, all sessions stepped together."""

    def __init__(self, engine: Engine, n_sessions: int, initial_dump_size: int = INITIAL_DUMP_SIZE_1,
                 max_dump_size: int = MAX_DUMP_SIZE, stop_on_eoa: bool = True, sampling: Optional[Sampling] = None,
                 slots: Optional[Sequence[int]] = None, bandwidth_id: int = 0, lanes: Optional[int] = None):
        self.e = engine
        self.n = n_sessions
        self.slots = list(slots) if slots is not None else list(range(n_sessions))
        assert len(self.slots) == n_sessions
        self.sched = [ChunkScheduler(dump_size=initial_dump_size, max_dump=max_dump_size, stop_on_eoa=stop_on_eoa,
                                     eoa=engine.cfg.eoa_token_id) for _ in range(n_sessions)]
        self.initial_dump_size = initial_dump_size
        self.stop_on_eoa = stop_on_eoa
        self.sampling = sampling or Sampling()
        self.bw = bandwidth_id
        self.steps_done = 0
        self.runner = LaneRunner(engine, lanes)

    def start(self, text_ids: Sequence[Sequence[int]], keep_schedule: bool = False):
        """Opens the sessions (the per-sentence reset of :404-416) and hands them their text ids.  A new call is a new
        request (fresh generator threads in the reference, so the dump size starts at its initial value again);
        `keep_schedule=True` continues the same request, where the dump size only ever grows (:373-375)."""
        assert len(text_ids) == self.n
        self.e.open(self.slots)
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
        slots = [self.slots[i] for i, _, _ in ready]
        starts = [s for _, s, _ in ready]
        counts = [c for _, _, c in ready]
        # a session may own several ready ranges in one round: the engine wants distinct slots per call
        codes_parts, order = [], []
        remaining = list(range(len(ready)))
        while remaining:
            seen, take, rest = set(), [], []
            for j in remaining:
                (rest if slots[j] in seen else take).append(j)
                seen.add(slots[j])
            codes_parts.append(self.e.gather_code_ranges([slots[j] for j in take], [starts[j] for j in take],
                                                          [counts[j] for j in take]))
            order.extend(take)
            remaining = rest
        codes = torch.cat(codes_parts) if len(codes_parts) > 1 else codes_parts[0]
        cu = [0]
        for j in order:
            cu.append(cu[-1] + counts[j])
        pcm = self.e.vocode(codes, cu, self.bw)
        host = self._pinned_pair(pcm.numel())[: pcm.numel()]
        host.copy_(pcm, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.e.device))
        return (ready, order, cu, host, ev, pcm)

    def _finish_emit(self, ticket, copy: bool = True) -> List[Chunk]:
        if ticket is None:
            return []
        ready, order, cu, host, ev, _pcm = ticket
        ev.synchronize()
        hop = self.e.cfg.hop
        out = []
        arr = host.numpy()
        for k, j in enumerate(order):
            i, s, c = ready[j]
            seg = arr[cu[k] * hop: cu[k + 1] * hop]
            out.append(Chunk(i, s, c, seg.copy() if copy else seg))
        out.sort(key=lambda ch: (ch.session, ch.start))
        return out

    def _pinned_pair(self, n: int):
        """Two pinned buffers used alternately: round r's PCM is read by the host while round r+1's is written
        (with copy=False a yielded chunk aliases its buffer and is valid until two rounds later)."""
        if not hasattr(self, "_pp") or self._pp[0].numel() < n:
            self._pp = [torch.empty((max(n, 1 << 20),), dtype=torch.float32, pin_memory=True) for _ in range(2)]
            self._pp_i = 0
        self._pp_i ^= 1
        return self._pp[self._pp_i]

    def run(self, max_steps: int, flush_tail: bool = False, copy: bool = True) -> Iterator[List[Chunk]]:
        """Decodes up to `max_steps` codes per session, yielding the chunks of each round as they are ready.

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
            new_codes = None
            if self.stop_on_eoa:   # the EOA test is the only reason a code value visits the host
                new_codes = self.e.gather_codes(slots, self.steps_done, k).cpu().numpy()
            ready: List[Tuple[int, int, int]] = []
            for a, i in enumerate(active):
                sc = self.sched[i]
                for t in range(k):
                    if sc.done:
                        break
                    for (s, c) in sc.push(int(new_codes[a, t]) if new_codes is not None else None):
                        ready.append((i, s, c))
            self.steps_done += k
            pending = self._enqueue_emit(ready)
            active = [i for i in active if not self.sched[i].done]
        if pending is not None:
            chunks = self._finish_emit(pending, copy)
            if chunks:
                yield chunks
        if flush_tail:
            ready = []
            for i in range(self.n):
                if not self.sched[i].done:
                    ready.extend((i, s, c) for (s, c) in self.sched[i].flush())
            chunks = self._finish_emit(self._enqueue_emit(ready), copy)
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
