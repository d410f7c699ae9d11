"""CPU fuzz of the text front-end's source: tests/tx_host_harness.cu compiles llmvox_b200/csrc/text_kernels.cuh for the host (nvcc is
needed, no GPU) and the ids it gives are compared with protocol.clean_text + tokenizer.sentence_ids -- the pipeline the reference's
producer / generator threads run (streaming_server.py:106-149, 184-248, 297-310) -- over the edge cases of test_gpu_text.py and
20,000 seeded random strings.  tests/test_gpu_text.py runs the same source on the device."""
import ctypes as C
import os
import random
import shutil
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def harness():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not found")
    out = os.path.join(ROOT, "tests", "_build", "libtx_host.so")
    src = os.path.join(ROOT, "tests", "tx_host_harness.cu")
    hdr = os.path.join(ROOT, "llmvox_b200", "csrc", "text_kernels.cuh")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run([nvcc, "-std=c++17", "-O1", "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-o", out, src], check=True, capture_output=True)
    lib = C.CDLL(out)
    lib.tx_host_ids.restype = C.c_int
    lib.tx_host_ids.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int32), C.c_int]

    def ids(text: str, clean: bool):
        raw = text.encode("utf-8")
        buf = np.zeros((13 * len(raw) + 80,), dtype=np.int32)
        n = lib.tx_host_ids(raw, len(raw), int(clean), buf.ctypes.data_as(C.POINTER(C.c_int32)), buf.size)
        assert 0 <= n <= buf.size
        return buf[:n].tolist()
    return ids


def _expect(s, clean):
    from llmvox_b200.protocol import clean_text
    from llmvox_b200.tokenizer import sentence_ids
    return sentence_ids(clean_text(s) if clean else s)


@pytest.mark.parametrize("clean", [True, False])
def test_edge_cases(harness, clean):
    from test_gpu_text import CASES
    for s in CASES:
        assert harness(s, clean) == _expect(s, clean), repr(s)


def test_random_strings(harness):
    rng = random.Random(11)
    dense = list("abcXYZ  .,.,--**##&@//\\\\019 \t\n") + [" ", "٣", "　", "é", " ", "EOS", "[PAD]", "\U0001d7d8",
                                                           "\x1c", "\x85", " ", "５", "\U0001f600", "..."]
    for it in range(20000):
        if it % 4 == 3:     # any code points: whitespace and digit classes from all over Unicode
            s = "".join(chr(rng.choice([rng.randint(1, 0x7f), rng.randint(0x80, 0x7ff), rng.randint(0x800, 0xd7ff), rng.randint(0xe000, 0xffff),
                                        rng.randint(0x10000, 0x10ffff)])) for _ in range(rng.randint(0, 24)))
        else:
            s = "".join(rng.choice(dense) for _ in range(rng.randint(0, 48)))
        clean = it % 5 != 0
        assert harness(s, clean) == _expect(s, clean), repr(s)


def test_every_digit_and_space_code_point(harness):
    """Each Unicode decimal digit and each whitespace character in the positions the rules test: '<d>.', '<d>,<d>', runs of spaces."""
    import unicodedata
    special = [chr(c) for c in range(0x110000) if unicodedata.category(chr(c)) == "Nd" or chr(c).isspace()]
    for ch in special:
        for s in (f"a{ch}. b", f"{ch},{ch}", f"x{ch}{ch}y", f"{ch}lead", f"trail{ch}", f"5.{ch}z", f"w {ch} v"):
            assert harness(s, True) == _expect(s, True), repr(s)
            assert harness(s, False) == _expect(s, False), repr(s)
