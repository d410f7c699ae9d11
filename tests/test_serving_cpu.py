"""Continuous batcher (llmvox_b200/serving.py) against the reference's own producer / generator loops / consumer
(tests/golden/replica_stream.npz, recorded by oracle/make_golden.py from streaming_server.py:184-248, :250-426, :428-469
with scripted code streams): chunk boundaries, control tokens and playback order, event for event.  The engine is replaced
by a scripted host backend here (the scheduling logic is host code); tests/test_serving_gpu.py runs the same loop on the GPU."""
import os

import numpy as np
import pytest

from llmvox_b200.serving import ContinuousBatcher

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "replica_stream.npz")


class ScriptedBackend:
    """Stand-in for GpuBackend: session k (in admission order) decodes scripts[k], then filler codes."""

    def __init__(self, scripts, eoa, max_context=1024, max_batch=8):
        self.scripts, self.eoa_token_id = scripts, eoa
        self.max_context, self.max_batch = max_context, max_batch
        self.slot_script, self.ctx, self.fed = {}, {}, {}
        self.opened = 0
        self.launches = []
        self.released = []

    def open(self, slots):
        for s in slots:
            self.slot_script[s] = self.scripts[self.opened]
            self.opened += 1
            self.ctx[s], self.fed[s] = 0, 0

    def feed(self, slots, ids):
        for s, x in zip(slots, ids):
            self.fed[s] += len(x)

    def release(self, slots):
        self.released.extend(slots)

    def sync_decode_streams(self):
        pass

    def launch(self, slots, k, sampling):
        assert k >= 1 and len(set(slots)) == len(slots)
        for s in slots:
            self.ctx[s] += k
            assert self.ctx[s] <= self.max_context
        self.launches.append((list(slots), k))
        return []

    def report(self, slots, events):
        out = []
        for s in slots:
            sc = self.slot_script[s]
            pos = sc.index(self.eoa_token_id) if self.eoa_token_id in sc else -1
            out.append((pos if 0 <= pos < self.ctx[s] else -1, self.ctx[s]))
        return out

    def wait_report(self, handle):
        return handle

    def emit(self, ready):
        return list(ready)

    def emit_done(self, ticket):
        return True

    def emit_finish(self, ticket):
        return [np.full((320 * c,), float(st), dtype=np.float32) for (_slot, st, c) in ticket]


def _events(req, r):
    out = []
    for rr, tag, val in req.events:
        if rr != r:
            continue
        out.append(val if tag == "chunk" else (-3 if val == "end" else -1 - int(val)))
    return out


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD, allow_pickle=True)


@pytest.mark.parametrize("feed", ["all_at_once", "word_per_round"])
@pytest.mark.parametrize("max_round_steps", [160, 7, 100000])
def test_batcher_matches_the_reference_loops_event_for_event(gold, feed, max_round_steps):
    scripts = [x.tolist() + [7] * 2000 for x in gold["scripts"]]
    be = ScriptedBackend(scripts, int(gold["eoa_id"]))
    b = ContinuousBatcher(None, backend=be, eos_token=str(gold["eos"]), max_round_steps=max_round_steps)
    outputs = gold["outputs"].tolist()
    played = []
    if feed == "all_at_once":
        req = b.submit(outputs)
        b.run_until_idle()
    else:
        req = b.open_request()
        for w in outputs:
            b.push_word(req, w)
            b.step()
        b.close_input(req)
        b.run_until_idle()
    while not req.out.empty():
        item = req.out.get()
        played.append(-3 if item is None else len(item) // (4 * 320))
    assert _events(req, 0) == gold["events_r0"].tolist()
    assert _events(req, 1) == gold["events_r1"].tolist()
    assert played == gold["playback"].tolist()
    assert req.done and not req.truncated
    assert len(be.released) == be.opened == 7                     # every session retired, every slot returned
    assert b.idle()


def test_batcher_admits_requests_into_a_running_batch(gold):
    """Continuous batching: a second answer submitted while the first is mid-decode shares its rounds, and both match the
    reference's events for their own word streams."""
    scripts = [x.tolist() + [7] * 2000 for x in gold["scripts"]]
    n = len(scripts)
    # the second request's sentences are admitted after the first's: their scripts follow in admission order
    be = ScriptedBackend(scripts + scripts, int(gold["eoa_id"]), max_batch=16)
    b = ContinuousBatcher(None, backend=be, eos_token=str(gold["eos"]))
    outputs = gold["outputs"].tolist()
    r1 = b.submit(outputs)
    for _ in range(3):
        b.step()
    assert not r1.done
    r2 = b.submit(outputs)
    b.run_until_idle()
    shared = [l for l in be.launches if len(l[0]) > n]
    assert shared, "the two answers never shared a decode round"
    for req in (r1, r2):
        assert _events(req, 0) == gold["events_r0"].tolist()
        assert _events(req, 1) == gold["events_r1"].tolist()
        assert req.done


def test_batcher_waits_for_slots_and_ends_answers_without_eos(gold):
    """More sentences than slots: later sentences wait for a retired session's slot (FIFO), events unchanged.  An answer
    whose word stream just stops (no EOS token) still ends: close_input makes its last sentence the end of generation."""
    scripts = [x.tolist() + [7] * 2000 for x in gold["scripts"]]
    be = ScriptedBackend(scripts, int(gold["eoa_id"]), max_batch=2)
    b = ContinuousBatcher(None, backend=be, slots=[0, 1], eos_token=str(gold["eos"]))
    outputs = gold["outputs"].tolist()
    outputs[-1] = outputs[-1].replace(str(gold["eos"]), "")
    req = b.submit(outputs)
    b.run_until_idle()
    assert _events(req, 0) == gold["events_r0"].tolist()          # incl. the final "end"
    assert _events(req, 1) == gold["events_r1"].tolist()
    assert req.done and max(len(l[0]) for l in be.launches) <= 2


def test_batcher_reports_truncation_at_max_context(gold):
    """No EOA within the engine's context: the sentence is flushed like the EOA branch and the request says so."""
    be = ScriptedBackend([[7] * 5000], int(gold["eoa_id"]), max_context=100)
    b = ContinuousBatcher(None, backend=be, eos_token=str(gold["eos"]))
    req = b.submit(["hello." + str(gold["eos"])])
    b.run_until_idle()
    assert req.done and req.truncated
    assert _events(req, 0) == [10, 30, 60, -3]
