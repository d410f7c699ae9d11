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
                 slots: Optional[Sequence[int]] = None, bandwidth_id: int = 0):
        self.e = engine
        self.n = n_sessions
        self.slots = list(slots) if slots is not None else list(range(n_sessions))
        assert len(self.slots) == n_sessions
        self.sched = [ChunkScheduler(dump_size=initial_dump_size, max_dump=max_dump_size, stop_on_eoa=stop_on_eoa,
                                     eoa=engine.cfg.eoa_token_id) for _ in range(n_sessions)]
        self.stop_on_eoa = stop_on_eoa
        self.sampling = sampling or Sampling()
        self.bw = bandwidth_id
        self.steps_done = 0
        self._pinned: Optional[torch.Tensor] = None

    def start(self, text_ids: Sequence[Sequence[int]]):
        """Opens the sessions (the per-sentence reset of :404-416) and hands them their text ids."""
        assert len(text_ids) == self.n
        self.e.open(self.slots)
        self.e.feed_text(self.slots, text_ids)
        for s in self.sched:
            s.new_sentence()
        self.steps_done = 0

    def _pinned_buf(self, n: int) -> torch.Tensor:
        if self._pinned is None or self._pinned.numel() < n:
            self._pinned = torch.empty((max(n, 1 << 20),), dtype=torch.float32, pin_memory=True)
        return self._pinned

    def _emit(self, ready: List[Tuple[int, int, int]], copy: bool = True) -> List[Chunk]:
        """ready: (session index, start, length).  One ragged vocoder batch + one async D2H."""
        if not ready:
            return []
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
        hop = self.e.cfg.hop
        host = self._pinned_buf(pcm.numel())[: pcm.numel()]
        host.copy_(pcm, non_blocking=True)
        torch.cuda.current_stream(self.e.device).synchronize()
        out = []
        arr = host.numpy()
        for k, j in enumerate(order):
            i, s, c = ready[j]
            seg = arr[cu[k] * hop: cu[k + 1] * hop]
            out.append(Chunk(i, s, c, seg.copy() if copy else seg))
        out.sort(key=lambda ch: (ch.session, ch.start))
        return out

    def run(self, max_steps: int, flush_tail: bool = False, copy: bool = True) -> Iterator[List[Chunk]]:
        """Decodes up to `max_steps` codes per session, yielding the chunks of each round as they are ready."""
        active = list(range(self.n))
        while active and self.steps_done < max_steps:
            k = min(min(self.sched[i].steps_to_next_event() for i in active), max_steps - self.steps_done)
            slots = [self.slots[i] for i in active]
            self.e.decode_steps(slots, k, self.sampling)
            new_codes = None
            if self.stop_on_eoa:   # the EOA test is the only reason a code value visits the host
                new_codes = self.e.gather_codes(slots, self.steps_done, k).cpu().numpy()
            ready: List[Tuple[int, int, int]] = []
            for a, i in enumerate(active):
                sc = self.sched[i]
                for t in range(k):
                    if sc.done:
                        break          # codes decoded past the EOA are discarded, as the reference resets there
                    for (s, c) in sc.push(int(new_codes[a, t]) if new_codes is not None else None):
                        ready.append((i, s, c))
            self.steps_done += k
            chunks = self._emit(ready, copy)
            active = [i for i in active if not self.sched[i].done]
            if chunks:
                yield chunks
        if flush_tail:
            ready = []
            for i in range(self.n):
                if not self.sched[i].done:
                    ready.extend((i, s, c) for (s, c) in self.sched[i].flush())
            chunks = self._emit(ready, copy)
            if chunks:
                yield chunks

    def codes(self, count: Optional[int] = None) -> np.ndarray:
        """(n_sessions, count) codes decoded so far (host copy)."""
        count = self.steps_done if count is None else count
        return self.e.gather_codes(self.slots, 0, count).cpu().numpy()


def synthesize(engine: Engine, text_ids: Sequence[Sequence[int]], max_steps: int, initial_dump_size: int = INITIAL_DUMP_SIZE_1,
               stop_on_eoa: bool = True, flush_tail: bool = True, sampling: Optional[Sampling] = None,
               bandwidth_id: int = 0) -> Tuple[np.ndarray, List[List[Chunk]]]:
    """Text ids -> (codes, per-session chunk lists)."""
    bs = BatchSynthesizer(engine, len(text_ids), initial_dump_size, stop_on_eoa=stop_on_eoa, sampling=sampling,
                          bandwidth_id=bandwidth_id)
    bs.start(text_ids)
    per: List[List[Chunk]] = [[] for _ in text_ids]
    for chunks in bs.run(max_steps, flush_tail=flush_tail):
        for ch in chunks:
            per[ch.session].append(ch)
    return bs.codes(), per
