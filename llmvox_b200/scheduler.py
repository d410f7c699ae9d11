"""Chunk-emission schedule of the reference's audio_generator_sync (streaming_server.py:357-422) as a pure
host state machine: one instance per (session, replica).  Ranges refer to the sentence's code history held on
the device, so no code ever has to visit the host except for the EOA test."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

MAX_DUMP_SIZE = 1280       # configs/inference_config.py:32
MAX_AUDIO_LENGTH = 8000    # configs/inference_config.py:33
EOA_TOKEN_ID = 453         # configs/inference_config.py:41
INITIAL_DUMP_SIZE_1 = 10   # configs/inference_config.py:30
INITIAL_DUMP_SIZE_2 = 160  # configs/inference_config.py:31


@dataclass
class ChunkScheduler:
    dump_size: int = INITIAL_DUMP_SIZE_1
    max_dump: int = MAX_DUMP_SIZE
    eoa: int = EOA_TOKEN_ID
    max_audio_len: int = MAX_AUDIO_LENGTH
    stop_on_eoa: bool = True
    # state of the current sentence
    emitted: int = 0            # codes of this sentence already cut into chunks (or dropped)
    seen: int = 0               # codes of this sentence pushed so far
    eoa_pending: bool = False   # an EOA code sits in the pending tail
    done: bool = False          # sentence ended (EOA or length cap)
    chunks: List[Tuple[int, int]] = field(default_factory=list)

    def _grow(self):
        if self.dump_size < self.max_dump:                      # :373-375
            self.dump_size = min(self.dump_size * 3, self.max_dump)

    def push(self, code: Optional[int]) -> List[Tuple[int, int]]:
        """Accounts for one more decoded code (pass None when the value is irrelevant, i.e. EOA handling is
        off).  Returns the (start, length) ranges that become ready, in emission order."""
        assert not self.done, "sentence already ended; call new_sentence()"
        out: List[Tuple[int, int]] = []
        self.seen += 1
        if self.stop_on_eoa and code == self.eoa:
            self.eoa_pending = True
        pending = self.seen - self.emitted
        if pending >= self.dump_size:                           # :357-376
            out.append((self.emitted, self.dump_size))
            self.emitted += self.dump_size
            # an EOA inside the cut leaves the pending tail only if it was the last code and got cut too
            if self.eoa_pending and self.emitted == self.seen:
                self.eoa_pending = False
            self._grow()
        elif self.eoa_pending:                                  # :379-394 flush everything, EOA included
            out.append((self.emitted, pending))
            self.emitted = self.seen
            self.eoa_pending = False
            self._grow()
        if (self.stop_on_eoa and code == self.eoa) or (self.seen - self.emitted) > self.max_audio_len:   # :397-422
            self.emitted = self.seen                            # pending codes are dropped
            self.eoa_pending = False
            self.done = True
            self._grow()
        self.chunks.extend(out)
        return out

    def push_many(self, k: int) -> List[Tuple[int, int]]:
        """`k` pushes of codes that are not EOA (or whose value is irrelevant) in one call: the ranges that become ready, in
        emission order -- identical to k calls of push(None), in O(emissions) instead of O(codes) (a 256-stream batch
        costs the host 51,200 pushes per 200-code step otherwise)."""
        assert not self.done, "sentence already ended; call new_sentence()"
        out: List[Tuple[int, int]] = []
        while k > 0:
            if self.eoa_pending or self.dump_size > self.max_audio_len:  # rare shapes (the length cap can trigger before the
                n0 = len(self.chunks)                                    # next cut): the reference's per-code order
                for _ in range(k):
                    if self.done:
                        break
                    self.push(None)
                self.chunks[n0:n0] = out
                return out + self.chunks[n0 + len(out):]
            need = self.dump_size - (self.seen - self.emitted)           # codes until the next cut (>= 1)
            if k < need:
                self.seen += k
                break
            self.seen += need
            k -= need
            out.append((self.emitted, self.dump_size))
            self.emitted += self.dump_size
            self._grow()
        self.chunks.extend(out)
        return out

    def flush(self) -> List[Tuple[int, int]]:
        """Not in the reference (it only flushes on EOA): emits whatever is pending, for fixed-length
        benchmark utterances."""
        out = []
        if self.seen > self.emitted:
            out.append((self.emitted, self.seen - self.emitted))
            self.emitted = self.seen
            self.chunks.extend(out)
        return out

    def steps_to_next_event(self) -> int:
        """Decode steps that can run before this scheduler may emit (EOA aside)."""
        return max(1, self.dump_size - (self.seen - self.emitted))

    def project(self, n: int) -> int:
        """steps_to_next_event() after `n` more codes that are not EOA, without changing any state (the continuous
        batcher sizes a round while the previous one is still in flight)."""
        pend, d = self.seen - self.emitted + n, self.dump_size
        while pend >= d:
            pend -= d
            if d < self.max_dump:
                d = min(d * 3, self.max_dump)
        return max(1, d - pend)

    def end_sentence(self):
        """Ends the sentence without an EOA code (the engine's context is full): the reset of :404-421 incl. the growth of
        the dump size; pending codes must have been flushed by the caller."""
        self.emitted = self.seen
        self.eoa_pending = False
        self.done = True
        self._grow()

    def new_sentence(self):
        """State reset of :404-416; dump_size is NOT reset between sentences."""
        self.emitted = 0
        self.seen = 0
        self.eoa_pending = False
        self.done = False
        self.chunks = []
