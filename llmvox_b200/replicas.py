"""Two-replica pipeline of one answer on top of the continuous batcher (serving.py); the pure protocol pieces (text
cleaning, sentence routing, per-word ids, queue mux) live in protocol.py and are re-exported here."""
from __future__ import annotations

from typing import Iterable, Iterator, List, Optional, Sequence

from .protocol import (DEFAULT_EOS, QueueItem, Sentence, SentenceRouter, clean_text, mux_audio_queues,  # noqa: F401
                       split_into_sentences, word_to_ids)
from .scheduler import INITIAL_DUMP_SIZE_1, INITIAL_DUMP_SIZE_2


class ReplicaPipeline:
    """Answers through the two replicas of one engine, in the reference's queue protocol (streaming_server.py:184-248,
    :357-422, :428-469), STREAMING: `stream()` consumes the upstream word stream as it arrives and yields each chunk's
    PCM bytes as soon as it is playable -- the first byte leaves after the first 10-code chunk, not after the answer.

    What is preserved, event for event (tests/golden/replica_stream.npz, recorded from the reference's own loop): which
    replica speaks which sentence; each replica's chunk schedule carried across ITS sentences within the answer (the dump
    size triples at every dump and at every sentence end and is never reset); the per-chunk independent vocoder decode;
    the control tokens; the order in which audio reaches the client.  Every answer starts from the initial dump sizes, as
    the reference starts fresh generator threads per request (:521-531)."""

    def __init__(self, engine, initial_dump_sizes: Sequence[int] = (INITIAL_DUMP_SIZE_1, INITIAL_DUMP_SIZE_2), max_dump_size: int = 1280,
                 max_audio_length: int = 8000, slots: Optional[Sequence[int]] = None, batcher=None, **batcher_kw):
        from .serving import ContinuousBatcher
        self.e = engine
        self.batcher = batcher or ContinuousBatcher(engine, slots=slots, initial_dump_sizes=initial_dump_sizes,
                                                    max_dump_size=max_dump_size, max_audio_length=max_audio_length, **batcher_kw)
        self.last_request = None

    def stream(self, outputs: Iterable[str]) -> Iterator[bytes]:
        """Word stream in (items as text_streamer_producer sees them), PCM bytes out in playback order, as they complete.
        Drives the batcher itself (single-threaded use); under `server.create_app` a worker thread drives it instead."""
        b = self.batcher
        req = b.open_request()
        self.last_request = req
        it = iter(outputs)
        exhausted = False
        while not req.done:
            if not exhausted:                         # one word per round: text arrives while earlier words decode
                try:
                    b.push_word(req, next(it))
                except StopIteration:
                    exhausted = True
                    b.close_input(req)
            b.step()
            while not req.out.empty():
                item = req.out.get()
                if item is not None:
                    yield item
        while not req.out.empty():
            item = req.out.get()
            if item is not None:
                yield item

    def run(self, outputs: Iterable[str], eos_token: str = DEFAULT_EOS, sampling=None):
        """Whole answer at once -> (queue items of replica 0, queue items of replica 1) as the reference's generator
        threads would have put them; feed them to `mux_audio_queues`."""
        import numpy as np
        for _ in self.stream(list(outputs)):
            pass
        req = self.last_request
        queues: List[List[QueueItem]] = [[], []]
        for r in (0, 1):
            for item in req.items[r]:
                queues[r].append(np.ascontiguousarray(item.pcm, dtype="<f4").tobytes() if hasattr(item, "pcm") else item)
            queues[r].append(None)                    # "Audio generator finished" (:424)
        return queues
