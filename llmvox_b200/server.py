"""Wire-format shim: a `/tts` endpoint with the reference's request / response format
(streaming_server.py:494-540; consumer client/endpoints.py:30-55) on top of the engine.

Response = `application/octet-stream`, chunked: the raw float32 little-endian mono 24 kHz PCM of each vocoder chunk, in
the order the two-replica protocol (replicas.py) plays them; nothing else is framed on the wire (the reference yields
`None` at end of answer, which the HTTP layer drops).  The upstream LLM / ASR / VLM of the reference's other endpoints
stay out of scope: the text to speak is the request's `text`, split into words like an LLM token stream would be."""
from typing import Iterator, List, Optional

import numpy as np

from .protocol import DEFAULT_EOS

SAMPLE_RATE = 24000


def pcm_to_wire(pcm: np.ndarray) -> bytes:
    """`audio.cpu().numpy().astype('float32').tobytes()` (streaming_server.py:368): float32, native = little endian."""
    return np.ascontiguousarray(pcm, dtype="<f4").tobytes()


def wire_to_pcm(data: bytes) -> np.ndarray:
    return np.frombuffer(data, dtype="<f4")


def text_to_word_stream(text: str, eos_token: str = DEFAULT_EOS) -> List[str]:
    """A whole text as the word stream an LLM streamer would have produced: words keep their leading space and the
    last one carries the EOS token (text_streamer_producer sees exactly such items, streaming_server.py:226-244)."""
    words = [w for w in text.strip().split(" ") if w]
    out = [(" " if i else "") + w for i, w in enumerate(words)]
    if out:
        out[-1] = out[-1] + eos_token
    return out


def tts_stream(batcher, text: str, eos_token: str = DEFAULT_EOS) -> Iterator[bytes]:
    """One request against a batcher that a worker thread is driving: submits the words, then blocks on the request's
    own output queue -- every chunk is yielded the moment it is playable (first byte after the first 10-code chunk)."""
    if callable(batcher):            # lazily built on the first request (create_app)
        batcher = batcher()
    req = batcher.submit(text_to_word_stream(text, eos_token))
    yield from req.chunks()


def create_app(model_handler, eos_token: str = DEFAULT_EOS):
    """FastAPI app with POST /tts {"text": ...} -> StreamingResponse, as the reference's endpoint.

    The reference shares one ModelHandler between all request threads without a lock, each with private KV caches.  Here
    ONE worker thread owns the engine and runs the continuous batcher (serving.py); request handlers only talk to it
    through thread-safe queues, so overlapping requests share decode rounds instead of racing on the engine."""
    import threading
    from fastapi import FastAPI
    from fastapi.responses import StreamingResponse
    from pydantic import BaseModel

    class TTSRequest(BaseModel):
        text: str

    app = FastAPI()
    cfg = model_handler.config
    state = {"batcher": None, "thread": None}
    lock = threading.Lock()

    def batcher():
        with lock:
            if state["batcher"] is None:
                from .serving import ContinuousBatcher
                b = ContinuousBatcher(model_handler.engine, slots=getattr(model_handler, "_batch_slots", None),
                                      initial_dump_sizes=(cfg["initial_dump_size_1"], cfg["initial_dump_size_2"]),
                                      max_dump_size=cfg["max_dump_size"], max_audio_length=cfg.get("max_audio_length", 8000),
                                      eos_token=eos_token)
                th = threading.Thread(target=b.serve_forever, name="llmvox-batcher", daemon=True)
                th.start()
                state["batcher"], state["thread"] = b, th
            return state["batcher"]

    @app.post("/tts")
    def tts(request: TTSRequest):
        return StreamingResponse(tts_stream(batcher, request.text, eos_token), media_type="application/octet-stream")

    app.state.llmvox = state
    return app
