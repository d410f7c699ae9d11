"""Two-replica ("multi-queue") streaming protocol of the reference server, pure host side.

The reference runs two TTS replicas per request (streaming_server.py:521-531): a producer thread routes the upstream
LLM's words to replica 0 or 1, switching at every sentence end (:184-248); each replica's generator thread puts PCM
chunks and control tokens on its audio queue (:357-422) -- `bytes` = a chunk, `1` / `0` = "now listen to replica 1 / 0",
`"end"` = the whole answer is done, `None` = generator finished -- and `audio_generator_async` (:428-469) drains the two
queues in that order, so replica 1's (larger, 160-code) first chunk plays right after replica 0's sentence.

Here the same protocol drives sessions of ONE engine (or of two engines on two GPUs): a replica is a chunk-schedule
state (`ChunkScheduler`, whose dump size is never reset between sentences) plus a stream of sentences."""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple, Union

from .scheduler import ChunkScheduler, INITIAL_DUMP_SIZE_1, INITIAL_DUMP_SIZE_2
from .tokenizer import ByT5Tokenizer, EOS_TEXT_ID

DEFAULT_EOS = "<|eot_id|>"           # configs/inference_config.py:39

# clean_text (streaming_server.py:106-149) as an ordered rule table: (pattern, replacement, is_regex)
_CLEAN_RULES: Tuple[Tuple[str, str, bool], ...] = (
    ("**", "", False),
    ("-", " ", False),
    (r"(\d)\.(?=\s|$)", r"\1", True),      # "5." -> "5"
    (r"\*", "", True),
    (r"#", " number ", True),
    (r"&", " and ", True),
    (r"@", " at ", True),
    (r"\s+", " ", True),
    (r"\.{3,}", " pause ", True),
    (r"(\d),(\d)", r"\1\2", True),          # thousands separators
    (r"\/+", " slash ", True),
    (r"\\+", " backslash ", True),
)


def clean_text(text: str, eos_token: str = DEFAULT_EOS) -> str:
    text = text.strip()
    for pat, rep, is_re in _CLEAN_RULES:
        text = re.sub(pat, rep, text) if is_re else text.replace(pat, rep)
    return text


class SentenceRouter:
    """text_streamer_producer's routing (streaming_server.py:226-244): skip '' and '-', strip, clean unless the word is
    the EOS token itself, drop if empty, send to the active replica, switch replicas after a word ending in '.'."""

    def __init__(self, eos_token: str = DEFAULT_EOS):
        self.eos = eos_token
        self.active = 0

    def route(self, output: str) -> Optional[Tuple[int, str]]:
        if output in ("", "-"):
            return None
        output = output.strip()
        if output != self.eos:
            output = clean_text(output, self.eos)
        if not output:
            return None
        dest = self.active
        if output.endswith("."):
            self.active = 1 - self.active
        return dest, output


def word_to_ids(text_token: str, eos_token: str = DEFAULT_EOS, tokenizer: Optional[ByT5Tokenizer] = None):
    """One queue item -> (text ids, end_of_speech, end_generation)  (streaming_server.py:297-310).  Note the reference's
    `rstrip(eos)` strips any trailing CHARACTERS that occur in the eos string, not the suffix; kept as is."""
    tok = tokenizer or ByT5Tokenizer()
    end_generation = False
    end_of_speech = False
    if (eos_token in text_token) or (text_token[-1:] == "."):
        if eos_token in text_token:
            end_generation = True
        text_token = text_token.rstrip(eos_token)
        end_of_speech = True
    ids = tok(text_token.strip())["input_ids"]
    if end_of_speech:
        ids = ids + [EOS_TEXT_ID]
    return ids, end_of_speech, end_generation


QueueItem = Union[bytes, int, str, None]


def mux_audio_queues(items_0: Iterable[QueueItem], items_1: Iterable[QueueItem]) -> Iterator[Optional[bytes]]:
    """audio_generator_async's ordering (streaming_server.py:440-465) over two finite item streams: start on queue 0;
    `bytes` are yielded; "end" yields None (end-of-answer marker for the HTTP layer); 0 / 1 switch the queue being
    drained; None is ignored.  Stops when the queue it is draining runs dry (the reference then blocks on it)."""
    its = [iter(items_0), iter(items_1)]
    cur = 0
    while True:
        try:
            item = next(its[cur])
        except StopIteration:
            return
        if isinstance(item, str) and item == "end":
            yield None
            continue
        if isinstance(item, int) and not isinstance(item, bool) and item in (0, 1):
            cur = item
            continue
        if item is None:
            continue
        yield item


@dataclass
class Sentence:
    replica: int
    words: List[str] = field(default_factory=list)
    ids: List[int] = field(default_factory=list)
    end_generation: bool = False


def split_into_sentences(outputs: Iterable[str], eos_token: str = DEFAULT_EOS) -> List[Sentence]:
    """Runs the router over the upstream word stream and groups the routed words into per-replica sentences (a
    sentence = the words up to and including the one that ends in '.' or carries the EOS token)."""
    router = SentenceRouter(eos_token)
    tok = ByT5Tokenizer()
    sentences: List[Sentence] = []
    open_by_replica = {0: None, 1: None}
    for out in outputs:
        r = router.route(out)
        if r is None:
            continue
        dest, word = r
        cur = open_by_replica[dest]
        if cur is None:
            cur = Sentence(dest)
            sentences.append(cur)
            open_by_replica[dest] = cur
        ids, eos_flag, end_gen = word_to_ids(word, eos_token, tok)
        cur.words.append(word)
        cur.ids.extend(ids)
        if eos_flag:
            cur.end_generation = end_gen
            open_by_replica[dest] = None
    return sentences
