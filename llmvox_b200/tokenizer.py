"""ByT5 byte tokenizer of the reference's `model_handler.tokenizer` (inference/model_handler.py:89-102):
google/byt5-small plus the two added special tokens "[PAD]" -> 384 and "EOS" -> 385.  Pure host code."""
from __future__ import annotations

from typing import Dict, List

PAD_TOKEN_ID = 384   # configs/inference_config.py:40
EOS_TEXT_ID = 385    # streaming_server.py:310
EOS_ID = 1           # </s> appended by the ByT5 tokenizer

_SPECIALS = (("[PAD]", PAD_TOKEN_ID), ("EOS", EOS_TEXT_ID))


class ByT5Tokenizer:
    """`tokenizer(word)["input_ids"]` as called at streaming_server.py:306: utf-8 byte + 3, trailing </s> = 1;
    the added special tokens are matched as literal substrings before byte encoding."""

    vocab_size = 386

    def encode(self, text: str, add_eos: bool = True) -> List[int]:
        out: List[int] = []
        i = 0
        while i < len(text):
            for lit, tid in _SPECIALS:
                if text.startswith(lit, i):
                    out.append(tid)
                    i += len(lit)
                    break
            else:
                out.extend(b + 3 for b in text[i].encode("utf-8"))
                i += 1
        if add_eos:
            out.append(EOS_ID)
        return out

    def __call__(self, text: str) -> Dict[str, List[int]]:
        ids = self.encode(text)
        return {"input_ids": ids, "attention_mask": [1] * len(ids)}

    def __len__(self) -> int:
        return self.vocab_size


def word_ids(word: str, sentence_end: bool) -> List[int]:
    """streaming_server.py:305-310: strip, tokenise, `+ [385]` at sentence end."""
    ids = ByT5Tokenizer().encode(word.strip())
    if sentence_end:
        ids = ids + [EOS_TEXT_ID]
    return ids


def sentence_ids(sentence: str) -> List[int]:
    """Text ids of one whole sentence as the reference's producer + generator threads feed it: the text is
    split at spaces (text_streamer_producer, streaming_server.py:184-248), every word gets its own </s>, and
    the last word of the sentence gets the 385 marker (:309-310)."""
    words = [w for w in sentence.strip().split(" ") if w != ""]
    ids: List[int] = []
    for i, w in enumerate(words):
        ids.extend(word_ids(w, i == len(words) - 1))
    return ids
