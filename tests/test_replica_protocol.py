"""Two-replica queue protocol against fixtures produced by the reference's own
clean_text / text_streamer_producer / audio_generator_async (oracle/make_golden.py protocol).  CPU only."""
import json
import os

import numpy as np
import pytest

from llmvox_b200.replicas import SentenceRouter, clean_text, mux_audio_queues, split_into_sentences, word_to_ids

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "protocol.json")


@pytest.fixture(scope="module")
def proto():
    return json.load(open(GOLD))


def _dec(x):
    return x["b"].encode() if isinstance(x, dict) else x


def test_clean_text_matches_reference(proto):
    for text, want in proto["clean_text"]:
        assert clean_text(text, proto["eos"]) == want, text


def test_router_matches_reference_producer(proto):
    for case in proto["router"]:
        r = SentenceRouter(proto["eos"])
        q = [[], []]
        for out in case["outputs"]:
            routed = r.route(out)
            if routed is not None:
                q[routed[0]].append(routed[1])
        assert q[0] == case["q0"] and q[1] == case["q1"], case["outputs"]


def test_mux_matches_reference_consumer(proto):
    for case in proto["mux"]:
        got = list(mux_audio_queues([_dec(x) for x in case["q0"]], [_dec(x) for x in case["q1"]]))
        assert got == [_dec(x) for x in case["yield"]]


def test_word_ids_sentence_end_and_eos(proto):
    eos = proto["eos"]
    ids, eos_flag, end = word_to_ids("there.", eos)
    assert ids[-2:] == [1, 385] and eos_flag and not end          # streaming_server.py:299-310
    ids, eos_flag, end = word_to_ids("thanks." + eos, eos)
    assert eos_flag and end and ids[-1] == 385
    ids, eos_flag, end = word_to_ids("plain", eos)
    assert ids == [115, 111, 100, 108, 113, 1] and not eos_flag and not end


def test_sentences_alternate_replicas(proto):
    case = proto["router"][0]
    sents = split_into_sentences(case["outputs"], proto["eos"])
    assert [s.replica for s in sents] == [0, 1, 0]
    assert [s.words for s in sents] == [["Hello", "there."], ["How", "are", "you."], ["Fine", "thanks." + proto["eos"]]]
    assert [s.end_generation for s in sents] == [False, False, True]
    assert all(s.ids[-1] == 385 for s in sents)


@pytest.mark.gpu
def test_replica_pipeline_end_to_end(weights):
    """Two sentences through two replicas of one engine (random-init weights never emit EOA 453, so both sentences run to
    the engine's context and are flushed there, reported as truncated): replica 0 speaks sentence 0 with chunks
    10/30/90 + the flushed rest, replica 1 speaks sentence 1 starting at 160; the muxed stream is sentence 0's audio, then
    sentence 1's, then the end marker."""
    from llmvox_b200.engine import Engine
    from llmvox_b200.replicas import ReplicaPipeline
    from oracle import llmvox_oracle as O
    ctx = 176
    e = Engine(weights, device=0, precision="fp32", max_sessions=4, max_context=ctx, max_vocode_frames=1024)
    eos = "<|eot_id|>"
    words = ["the", " quick", " brown", " fox", " jumps.", " over", " the", " lazy", " dog." + eos]
    pipe = ReplicaPipeline(e)
    q0, q1 = pipe.run(words, eos)
    sents = split_into_sentences(words, eos)
    stream = list(mux_audio_queues(q0, q1))
    assert stream[-1] is None and all(isinstance(x, bytes) for x in stream[:-1])
    lens0 = [len(x) // 1280 for x in q0 if isinstance(x, bytes)]
    lens1 = [len(x) // 1280 for x in q1 if isinstance(x, bytes)]
    assert lens0 == [10, 30, 90, ctx - 130] and lens1 == [160, ctx - 160]
    assert q0[-2:] == [1, None] and q1[-2:] == ["end", None]
    assert pipe.last_request.truncated
    # first chunk of sentence 0 == oracle decode of the oracle's first 10 codes
    codes = O.decode_steps(weights, O.GPTArch(), sents[0].ids, 10)
    ref = O.vocoder_decode(weights, codes).numpy()
    got = np.frombuffer(stream[0], dtype=np.float32)
    err = got.astype(np.float64) - ref
    assert 10 * np.log10((ref.astype(np.float64) ** 2).sum() / (err ** 2).sum()) > 80
    # a new answer starts from the initial dump sizes again (fresh generator threads per request, streaming_server.py:521-531)
    q0b, _ = pipe.run(words[:5], eos)
    assert [len(x) // 1280 for x in q0b if isinstance(x, bytes)][:3] == [10, 30, 90]
    e.close()


def test_wire_format_helpers():
    from llmvox_b200.server import pcm_to_wire, text_to_word_stream, wire_to_pcm
    x = np.array([0.5, -1.0, 3.25], dtype=np.float32)
    b = pcm_to_wire(x)
    assert b == x.astype("<f4").tobytes() and len(b) == 12
    assert (wire_to_pcm(b) == x).all()
    ws = text_to_word_stream("Hi there. Second one.")
    assert ws == ["Hi", " there.", " Second", " one.<|eot_id|>"]
    sents = split_into_sentences(ws)
    assert [s.replica for s in sents] == [0, 1] and sents[1].end_generation


def test_tts_endpoint_wiring_on_cpu(monkeypatch):
    """The ASGI layer alone (stub chunk source): POST /tts -> 200, octet-stream, chunks concatenated untouched."""
    from fastapi.testclient import TestClient
    import llmvox_b200.server as S

    class StubHandler:
        config = {"initial_dump_size_1": 10, "initial_dump_size_2": 160, "max_dump_size": 1280}
        engine = None
    chunks = [np.arange(320, dtype=np.float32).tobytes(), np.zeros(640, dtype=np.float32).tobytes()]
    monkeypatch.setattr(S, "tts_stream", lambda batcher, text, eos=S.DEFAULT_EOS: iter(chunks))
    r = TestClient(S.create_app(StubHandler())).post("/tts", json={"text": "hi."})
    assert r.status_code == 200 and r.headers["content-type"].startswith("application/octet-stream")
    assert r.content == b"".join(chunks)
    assert TestClient(S.create_app(StubHandler())).post("/tts", json={}).status_code == 422
