"""Continuous-batching streaming synthesis: the reference's per-request generator threads as ONE engine loop.

The reference serves a request with a producer thread that routes the upstream LLM's words to two TTS replicas
(streaming_server.py:184-248) and one generator thread per replica that decodes one code per iteration, cuts chunks by
the growing dump size, vocodes each chunk alone and puts PCM + control tokens on its audio queue (:250-426);
`audio_generator_async` (:428-469) drains the two queues in the order the control tokens dictate.  Many requests = many
threads sharing one ModelHandler without a lock.

Here ONE loop owns the engine (`ContinuousBatcher.step`, driven by `serve_forever` on a worker thread or by the caller):

* admission / retirement between rounds: a sentence becomes an engine session the moment its first word arrives (words
  are fed as they arrive, exactly like the reference's `text_token_queue.get()`), and its slot returns to the pool when its
  last chunk has been cut -- requests join and leave a running batch (continuous batching);
* every round decodes `k` steps for ALL live sessions in one call (`k` = steps until the next session can possibly emit)
  with no host sync per token; the end of a sentence (EOA, streaming_server.py:379, 397) is detected ON THE DEVICE
  (SessionState.eoa_pos) and read back as one small pinned copy per round (lvx_session_progress);
* the loop is software-pipelined with one round of speculation: round r+1 is enqueued (assuming no EOA in round r) BEFORE
  the host waits for round r's progress report, and round r's chunks are vocoded on the control stream while round r+1
  decodes -- so the EOA test never idles the GPU.  Codes a session decodes past its EOA are discarded, as the reference
  resets there (:404-416);
* chunk boundaries, control tokens and playback order are the reference's, event for event: each (request, replica) owns
  ONE ChunkScheduler whose dump size is never reset (it triples at every dump and at every sentence end, :373-375,
  :418-421), sentences of a replica are scheduled strictly in order (codes of a later sentence are decoded concurrently
  but cut into chunks only once its predecessor's final dump size is known), and the two replica queues are muxed like
  `audio_generator_async`.

PAD-until-EOA: a sentence is decoded until its EOA code or until pending codes exceed max_audio_length (:397), never
by a text-length heuristic; the only extra bound is the engine's max_context, and hitting it is reported
(`StreamRequest.truncated`)."""
from __future__ import annotations

import queue
import threading
import time
from collections import deque
from dataclasses import dataclass, field
from typing import Deque, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .engine import Sampling
from .protocol import DEFAULT_EOS, SentenceRouter, word_to_ids
from .scheduler import ChunkScheduler, INITIAL_DUMP_SIZE_1, INITIAL_DUMP_SIZE_2, MAX_AUDIO_LENGTH, MAX_DUMP_SIZE
from .tokenizer import ByT5Tokenizer


class GpuBackend:
    """The engine calls the batcher makes, on the streams it makes them on (tests substitute a scripted host object to
    pin the scheduling logic to reference fixtures without a GPU; the product path has only this one)."""

    def __init__(self, engine, bandwidth_id: int = 0, lanes: Optional[int] = None):
        from .streaming import ChunkEmitter, LaneRunner
        self.e = engine
        self.max_context = engine.cfg.max_context
        self.max_batch = engine.cfg.max_batch
        self.eoa_token_id = engine.cfg.eoa_token_id
        self.runner = LaneRunner(engine, lanes)
        self.emitter = ChunkEmitter(engine, bandwidth_id)
        self._progress = [torch.empty((max(2, 2 * engine.cfg.max_batch),), dtype=torch.int32, pin_memory=True) for _ in range(2)]
        self._flip = 0

    def open(self, slots):
        self.e.open(slots)

    def feed(self, slots, ids):
        self.e.feed_text(slots, ids)

    def release(self, slots):
        self.e.release(slots)

    def sync_decode_streams(self):
        self.runner.sync_from_control()

    def launch(self, slots, k, sampling):
        return self.runner.launch(slots, k, sampling)

    def report(self, slots, events):
        """Control stream: wait for the round, then one pinned copy of (eoa_pos, ctx_len) per session."""
        self.runner.join(events)
        self._flip ^= 1
        buf = self._progress[self._flip]
        self.e.session_progress(slots, buf)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.e.device))
        return (buf, ev, len(slots))

    def wait_report(self, handle):
        buf, ev, n = handle
        ev.synchronize()
        return buf[: 2 * n].view(-1, 2).tolist()

    def emit(self, ready):
        return self.emitter.enqueue(ready)

    def emit_done(self, ticket):
        return self.emitter.done(ticket)

    def emit_finish(self, ticket):
        return self.emitter.finish(ticket)


@dataclass
class _Sentence:
    req: "StreamRequest"
    replica: int
    ids: List[int] = field(default_factory=list)
    slot: int = -1                 # engine session slot (-1: waiting for a free one)
    fed: int = 0                   # ids already handed to the engine
    complete: bool = False         # the sentence's last word has arrived (ids beyond it are PAD, :316-320)
    end_generation: bool = False
    issued: int = 0                # decode steps enqueued so far (host projection of the session's context length)
    known: int = 0                 # context length confirmed by the last consumed progress report
    eoa: Optional[int] = None      # position of the EOA code once the device has reported it
    pushed: int = 0                # codes accounted for by the replica's scheduler
    finished: bool = False


class _PendingChunk:
    """Placeholder of a chunk on a replica's audio queue until its PCM has left the vocoder."""
    __slots__ = ("pcm",)

    def __init__(self):
        self.pcm: Optional[np.ndarray] = None


QueueItem = Union[_PendingChunk, int, str, None]


class StreamRequest:
    """One answer: two replicas (chunk schedules 10 / 160 -> x3 -> 1280), their sentence FIFOs and audio queues."""

    def __init__(self, rid: int, dump_sizes: Sequence[int], max_dump: int, max_audio_len: int, eoa: int, eos_token: str):
        self.rid = rid
        self.eos = eos_token
        self.router = SentenceRouter(eos_token)
        self.sched = [ChunkScheduler(dump_size=d, max_dump=max_dump, eoa=eoa, max_audio_len=max_audio_len) for d in dump_sizes]
        self.fifo: List[Deque[_Sentence]] = [deque(), deque()]
        self.open_sentence: List[Optional[_Sentence]] = [None, None]
        self.last_sentence: Optional[_Sentence] = None
        self.saw_end = False
        self.queues: List[Deque[QueueItem]] = [deque(), deque()]
        self.mux_cur = 0
        self.input_closed = False
        self.done = False
        self.truncated = False            # a sentence hit the engine's max_context before its EOA
        self.out: "queue.Queue[Optional[bytes]]" = queue.Queue()      # PCM chunks in playback order; None = end of answer
        self.t_submit = time.perf_counter()
        self.t_first: Optional[float] = None
        self.items: List[List[QueueItem]] = [[], []]    # everything ever put on each replica's audio queue, in order
        self.events: List[Tuple[int, str, object]] = [] # (replica, "chunk" | "ctrl", chunk length | control token), same order

    # ---- consumer side
    def chunks(self) -> Iterator[bytes]:
        """Blocks for each chunk (float32 PCM bytes, streaming_server.py:368) until the end of the answer."""
        while True:
            item = self.out.get()
            if item is None:
                return
            yield item


class ContinuousBatcher:
    def __init__(self, engine, slots: Optional[Sequence[int]] = None,
                 initial_dump_sizes: Sequence[int] = (INITIAL_DUMP_SIZE_1, INITIAL_DUMP_SIZE_2), max_dump_size: int = MAX_DUMP_SIZE,
                 max_audio_length: int = MAX_AUDIO_LENGTH, sampling: Optional[Sampling] = None, bandwidth_id: int = 0,
                 lanes: Optional[int] = None, eos_token: str = DEFAULT_EOS, max_round_steps: int = 160, backend=None):
        self.b = backend if backend is not None else GpuBackend(engine, bandwidth_id, lanes)
        self.free: List[int] = list(slots) if slots is not None else list(range(self.b.max_batch))
        self.free.reverse()
        self.dump_sizes = tuple(initial_dump_sizes)
        self.max_dump = max_dump_size
        self.max_audio_len = max_audio_length
        self.sampling = sampling or Sampling()
        self.eos = eos_token
        # a round never runs longer than this many steps, so a request that arrives while another is deep in a 1280-code
        # chunk joins the batch within ~25 ms (the reference starts fresh threads per request instead)
        self.max_round_steps = max_round_steps
        self.tok = ByT5Tokenizer()
        self.requests: List[StreamRequest] = []
        self.waiting: Deque[_Sentence] = deque()          # sentences without a slot yet
        self.live: List[_Sentence] = []                   # sentences holding a slot
        self._rid = 0
        self._round_id = 0
        self._cur = None                                  # round whose progress report has not been consumed yet
        self._tickets: Deque = deque()                    # emit tickets in flight: (ticket, [pending chunks])
        self._zombies: List[Tuple[int, int]] = []         # (round id after which the slot is free, slot)
        self._round_id_consumed = 0
        self._lock = threading.Lock()
        self._wake = threading.Condition(self._lock)
        self._inbox: Deque[Tuple[StreamRequest, Optional[str]]] = deque()
        self._stop = False
        self.rounds = 0

    # ------------------------------------------------------------------ producer side (any thread)
    def open_request(self) -> StreamRequest:
        with self._lock:
            self._rid += 1
            req = StreamRequest(self._rid, self.dump_sizes, self.max_dump, self.max_audio_len, self.b.eoa_token_id, self.eos)
        return req

    def push_word(self, req: StreamRequest, output: str):
        """One item of the upstream word stream (what text_streamer_producer sees, :226-244).  Thread-safe."""
        with self._wake:
            self._inbox.append((req, output))
            self._wake.notify()

    def close_input(self, req: StreamRequest):
        """No more words will come (an answer normally ends with the EOS token; this also ends one that does not)."""
        with self._wake:
            self._inbox.append((req, None))
            self._wake.notify()

    def submit(self, outputs: Iterable[str]) -> StreamRequest:
        req = self.open_request()
        for o in outputs:
            self.push_word(req, o)
        self.close_input(req)
        return req

    # ------------------------------------------------------------------ engine loop (one thread)
    def _drain_inbox(self):
        with self._lock:
            items = list(self._inbox)
            self._inbox.clear()
        for req, output in items:
            if req not in self.requests:
                self.requests.append(req)
            if output is None:
                self._close_input(req)
                continue
            routed = req.router.route(output)
            if routed is None:
                continue
            dest, word = routed
            ids, eos_flag, end_gen = word_to_ids(word, req.eos, self.tok)
            s = req.open_sentence[dest]
            if s is None:
                s = _Sentence(req, dest)
                req.open_sentence[dest] = s
                req.fifo[dest].append(s)
                req.last_sentence = s
                self.waiting.append(s)
            s.ids.extend(ids)
            if eos_flag:
                s.complete = True
                s.end_generation = end_gen
                req.saw_end = req.saw_end or end_gen
                req.open_sentence[dest] = None

    def _close_input(self, req: StreamRequest):
        """The word stream is over.  The reference's answers end with the EOS token (-> "end" on the audio queue, :398); an
        answer that stops without it still has to end: its last sentence becomes the end of generation."""
        if req.input_closed:
            return
        req.input_closed = True
        for r in (0, 1):
            s = req.open_sentence[r]
            if s is not None:
                s.complete = True
                req.open_sentence[r] = None
        last = req.last_sentence
        if last is None:
            self._finish_request(req)
        elif not req.saw_end:
            if not last.finished:
                last.end_generation = True
            else:                                         # its switch token is already queued: the consumer is now on the
                req.queues[1 - last.replica].append("end")    # other replica's queue
                req.items[1 - last.replica].append("end")
                req.events.append((1 - last.replica, "ctrl", "end"))
            req.saw_end = True

    def _admit(self) -> bool:
        """Gives waiting sentences a slot and hands newly arrived text ids to the engine.  -> anything enqueued."""
        self._reclaim_zombies()
        opened = []
        while self.waiting and self.free:
            s = self.waiting.popleft()
            s.slot = self.free.pop()
            opened.append(s.slot)
            self.live.append(s)
        if opened:
            self.b.open(opened)
        feed = [s for s in self.live if s.fed < len(s.ids)]
        if feed:
            self.b.feed([s.slot for s in feed], [s.ids[s.fed:] for s in feed])
            for s in feed:
                s.fed = len(s.ids)
        if opened or feed:
            self.b.sync_decode_streams()
        return bool(opened or feed)

    def _reclaim_zombies(self):
        self._zombies = [(rid, slot) for rid, slot in self._zombies if not self._reclaim(rid, slot)]

    def _reclaim(self, round_id: int, slot: int) -> bool:
        if self._round_id_consumed >= round_id:
            self.b.release([slot])
            self.free.append(slot)
            return True
        return False

    def _budget(self, s: _Sentence) -> int:
        """Decode steps sentence s may take now."""
        if s.eoa is not None or s.finished:
            return 0
        lim = self.b.max_context - s.issued
        if not s.complete:
            lim = min(lim, s.fed - s.issued)      # the reference blocks on the next word here (:291)
        return max(0, lim)

    def _launch(self):
        """Enqueues the next round on the decode streams (speculating that the round in flight meets no EOA)."""
        part, k = [], None
        for s in self.live:
            b = self._budget(s)
            if b <= 0:
                continue
            part.append(s)
            k = b if k is None else min(k, b)
            fifo = s.req.fifo[s.replica]
            if fifo and fifo[0] is s:             # head of its replica: the next emission bounds the round
                k = min(k, s.req.sched[s.replica].project(s.issued - s.pushed))
        if not part:
            return None
        k = min(k, self.max_round_steps)
        events = self.b.launch([s.slot for s in part], k, self.sampling)
        for s in part:
            s.issued += k
        self._round_id += 1
        return {"id": self._round_id, "sentences": part, "k": k, "events": events, "end": [s.issued for s in part], "handle": None}

    def _report(self, rnd):
        rnd["handle"] = self.b.report([s.slot for s in rnd["sentences"]], rnd["events"])

    def _consume(self, rnd):
        """Host: progress of round `rnd` -> schedulers -> chunk ranges -> one ragged vocoder batch."""
        rep = self.b.wait_report(rnd["handle"])
        for s, end, (eoa, _ctx) in zip(rnd["sentences"], rnd["end"], rep):
            if s.finished:
                continue
            # codes [0, end) are final: `end` is the context length this round was launched to reach (the device's own
            # counter may already be ahead, the next round is running), and eoa_pos only ever goes from -1 to its value
            s.known = max(s.known, end)
            if eoa >= 0 and s.eoa is None:
                s.eoa = eoa
        self._round_id_consumed = rnd["id"]
        ready: List[Tuple[int, int, int]] = []
        pend: List[_PendingChunk] = []
        touched = {id(s.req): s.req for s in rnd["sentences"]}
        for req in touched.values():
            for r in (0, 1):
                self._schedule(req, r, ready, pend)
        if ready:
            self._tickets.append((self.b.emit(ready), pend))

    def _schedule(self, req: StreamRequest, r: int, ready, pend):
        fifo, sc = req.fifo[r], req.sched[r]
        while fifo:
            s = fifo[0]
            if s.slot < 0:
                break
            avail = s.known if s.eoa is None else min(s.known, s.eoa + 1)
            ended = False
            plain = (avail - s.pushed) if s.eoa is None else (min(avail, s.eoa) - s.pushed)   # codes before the EOA: in one go
            if plain > 1:
                for (st, ln) in sc.push_many(plain):
                    self._queue_chunk(req, r, s, st, ln, ready, pend)
                s.pushed += plain
                ended = sc.done
            while s.pushed < avail and not ended:
                code = sc.eoa if (s.eoa is not None and s.pushed == s.eoa) else None
                for (st, ln) in sc.push(code):
                    self._queue_chunk(req, r, s, st, ln, ready, pend)
                s.pushed += 1
                ended = sc.done
            if not ended and s.complete and s.eoa is None and s.known >= self.b.max_context:
                # the engine's context is full before the EOA code: flush what is pending like the EOA branch (:379-394)
                for (st, ln) in sc.flush():
                    self._queue_chunk(req, r, s, st, ln, ready, pend)
                sc.end_sentence()
                req.truncated = True
                ended = True
            if not ended:
                break
            ctrl = "end" if s.end_generation else 1 - r          # :398-403
            req.queues[r].append(ctrl)
            req.items[r].append(ctrl)
            req.events.append((r, "ctrl", ctrl))
            s.finished = True
            sc.new_sentence()
            fifo.popleft()
            self.live.remove(s)
            self._zombies.append((self._round_id, s.slot))       # a speculative round may still be running on the slot

    def _queue_chunk(self, req, r, s, start, length, ready, pend):
        p = _PendingChunk()
        req.queues[r].append(p)
        req.items[r].append(p)
        req.events.append((r, "chunk", length))
        ready.append((s.slot, start, length))
        pend.append(p)

    def _mux(self, req: StreamRequest):
        """audio_generator_async (:440-465) over the live queues: hand out what is playable now."""
        while not req.done:
            q = req.queues[req.mux_cur]
            if not q:
                return
            item = q[0]
            if isinstance(item, _PendingChunk):
                if item.pcm is None:
                    return
                q.popleft()
                if req.t_first is None:
                    req.t_first = time.perf_counter()
                req.out.put(np.ascontiguousarray(item.pcm, dtype="<f4").tobytes())
            elif isinstance(item, str):                           # "end"
                q.popleft()
                self._finish_request(req)
            elif item is None:
                q.popleft()
            else:
                q.popleft()
                req.mux_cur = int(item)

    def _finish_request(self, req: StreamRequest):
        req.done = True
        req.out.put(None)
        if req in self.requests:
            self.requests.remove(req)

    def _collect(self, block: bool):
        """Tickets whose PCM has reached the host -> resolve placeholders -> mux."""
        while self._tickets and (block or self.b.emit_done(self._tickets[0][0])):
            ticket, pend = self._tickets.popleft()
            for p, pcm in zip(pend, self.b.emit_finish(ticket)):
                p.pcm = pcm
            block = False
        for req in list(self.requests):
            self._mux(req)

    def idle(self) -> bool:
        return not (self.live or self.waiting or self._cur is not None or self._tickets or self._inbox or self._zombies)

    def step(self):
        """One pipelined round (see the module docstring)."""
        self._drain_inbox()
        self._admit()
        nxt = self._launch()                    # speculative: round r+1 before round r's report is read
        self._collect(block=False)
        if self._cur is not None:
            self._consume(self._cur)            # waits for round r only; enqueues its vocoder batch
        if nxt is not None:
            self._report(nxt)
        self._cur = nxt
        self._collect(block=nxt is None)        # nothing left to overlap with: wait for the PCM
        for req in list(self.requests):
            self._mux(req)
        self._reclaim_zombies()
        self.rounds += 1

    def run_until_idle(self):
        while not self.idle():
            self.step()

    def serve_forever(self):
        """Worker-thread body: steps while there is work, sleeps on the inbox otherwise."""
        eng = getattr(self.b, "e", None)
        if eng is not None:
            torch.cuda.set_device(eng.device)
        while True:
            with self._wake:
                while not self._stop and self.idle():
                    self._wake.wait(timeout=0.5)
                if self._stop:
                    return
            self.step()

    def shutdown(self):
        with self._wake:
            self._stop = True
            self._wake.notify_all()
