"""Device text front-end (lvx_feed_utf8): ids identical to the host pipeline protocol.clean_text + tokenizer.sentence_ids,
which test_host_cpu.py pins to the reference's behaviour (streaming_server.py:106-149, 184-248, 297-310).  Integer work:
the bar is exact equality."""
import os
import random
import re
import sys
import unicodedata

import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _nd_runs():
    runs, start, prev = [], None, None
    for c in range(0x110000):
        if unicodedata.category(chr(c)) == "Nd":
            if start is None:
                start = c
            prev = c
        elif start is not None:
            runs.append((start, prev))
            start = None
    return runs


def test_digit_and_space_tables_are_pythons():
    """The kernel's \\d table = this interpreter's Unicode category Nd; its \\s predicate = str.isspace."""
    src = open(os.path.join(ROOT, "llmvox_b200", "csrc", "text_kernels.cuh")).read()
    body = src[src.index("#define TX_DIGIT_RUNS"):]
    body = body[: body.index("\n")]
    table = [(int(a, 16), int(b, 16)) for a, b in re.findall(r"\{0x([0-9A-Fa-f]+), 0x([0-9A-Fa-f]+)\}", body)]
    assert table == _nd_runs()
    spaces = [c for c in range(0x110000) if chr(c).isspace()]
    mine = [c for c in range(0x110000)
            if (0x9 <= c <= 0xD) or (0x1C <= c <= 0x20) or c in (0x85, 0xA0, 0x1680, 0x2028, 0x2029, 0x202F, 0x205F, 0x3000) or (0x2000 <= c <= 0x200A)]
    assert mine == spaces
    assert all(re.fullmatch(r"\s", chr(c)) for c in spaces)


CASES = [
    "hello there.", "Hello,   world -- this is **bold** text.", "  leading and trailing \t\n", "", " ", "a", ".", "...", "....", "..",
    "wait... what", "5.", "5. ", "it costs 5. and 6.5 or 7.", "1,2,3", "1,000,000 dollars", "12,345.", "a,b", "1, 2", ",1", "1,",
    "#1 & #2 @ home", "#", "&&", "@@@", "a/b", "a//b///c", "\\", "\\\\\\", "path\\to\\file/or/this", "*", "**", "***", "****", "* * *", "*-*",
    "co-operate - - -", "-", "--", "tab\tseparated\nlines\r\nhere", "nbsp between", "ideographic　space", "thin space and line",
    "٥.", "١,٢", "price ５. ok", "\U0001d7ce,\U0001d7cf", "café naïve 日本語 \U0001f600", "\x1c\x1d\x1e\x1f x \x85",
    "[PAD]", "EOS", "xEOSy [PAD][PAD] EO S", "the EOS token", "[PAD", "PAD]", "E OS", "3.14...2,5//\\\\#&@**-", "end with dash-", "dot.dot.dot.", "5.\n", "5. x",
    "a" * 180, "word " * 30, "\\" * 12, "5." * 20,
]


def _expect(s, clean):
    from llmvox_b200.protocol import clean_text
    from llmvox_b200.tokenizer import sentence_ids
    return sentence_ids(clean_text(s) if clean else s)


@pytest.fixture(scope="module")
def engine():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from llmvox_b200 import weights as W
    from llmvox_b200.engine import Engine
    e = Engine(W.make_random_weights(7, wpe_rows=512), device=0, precision="bf16", max_sessions=128, max_context=512, max_vocode_frames=256)
    yield e
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("clean", [True, False])
def test_device_ids_equal_host_pipeline(engine, clean):
    e = engine
    slots = list(range(len(CASES)))
    e.open(slots)
    counts = e.feed_sentences(slots, CASES, clean=clean)
    for s, slot, cnt in zip(CASES, slots, counts):
        want = _expect(s, clean)
        assert cnt == len(want), (s, cnt, len(want))
        assert e.session_text(slot) == want, repr(s)
    e.release(slots)


@pytest.mark.gpu
def test_device_ids_random_sentences(engine):
    """Seeded random sentences over an alphabet dense in the characters clean_text rewrites."""
    e = engine
    rng = random.Random(5)
    alphabet = list("abcXYZ  .,.,--**##&@//\\\\019 \t\n") + [" ", "٣", "　", "é", " ", "EOS", "[PAD]", "\U0001d7d8"]
    slots = list(range(128))
    for _ in range(6):
        sents = ["".join(rng.choice(alphabet) for _ in range(rng.randint(0, 60))) for _ in slots]
        e.open(slots)
        counts = e.feed_sentences(slots, sents)
        for s, slot, cnt in zip(sents, slots, counts):
            want = _expect(s, True)
            assert cnt == len(want) and e.session_text(slot) == want, repr(s)
    e.release(slots)


@pytest.mark.gpu
def test_feed_sentences_appends_and_decodes_like_feed_text(engine):
    """Two sentences fed one after the other append; decoding from device-tokenised text gives the same codes as from host ids."""
    from llmvox_b200.protocol import clean_text
    from llmvox_b200.tokenizer import sentence_ids
    e = engine
    a, b = "The year 1,999 was #1.", "Dr. Who & co... went home."
    e.open([0, 1])
    e.feed_sentences([0], [a])
    e.feed_sentences([0], [b])
    e.feed_text([1], [sentence_ids(clean_text(a)) + sentence_ids(clean_text(b))])
    assert e.session_text(0) == e.session_text(1)
    e.decode_steps([0, 1], 24)
    codes = e.gather_codes([0, 1], 0, 24).cpu().numpy()
    assert codes[0].tolist() == codes[1].tolist()
    e.release([0, 1])


@pytest.mark.gpu
def test_feed_sentences_capacity_errors(engine):
    from llmvox_b200._lib import LvxError
    e = engine
    e.open([0])
    with pytest.raises(LvxError):
        e.feed_sentences([0], ["x" * 600])            # longer than max_context bytes
    with pytest.raises(LvxError):
        e.feed_sentences([0], ["a " * 256])           # 512 bytes -> 513 ids: does not fit 512
    assert e.session_text(0) == []                    # nothing was appended
    e.feed_sentences([0], ["ok."])
    assert len(e.session_text(0)) == 5
    e.release([0])


@pytest.mark.gpu
def test_long_rewrites_fall_back_to_the_global_scratch():
    """Sentences whose rewritten form outgrows the per-thread shared-memory scratch (1024 bytes) are redone in global memory."""
    import torch
    from llmvox_b200 import weights as W
    from llmvox_b200.engine import Engine
    e = Engine(W.make_random_weights(7, wpe_rows=4096), device=0, precision="bf16", max_sessions=8, max_context=4096, max_vocode_frames=256)
    sents = ["&" * 300, "\\a" * 250, "x" * 1023 + ".", "y" * 1025, "#7, " * 150, "short one.", "", "a-b " * 400]
    slots = list(range(len(sents)))
    e.open(slots)
    counts = e.feed_sentences(slots, sents)
    for s_, slot, cnt in zip(sents, slots, counts):
        want = _expect(s_, True)
        assert cnt == len(want) and e.session_text(slot) == want, repr(s_[:20])
    e.close()
