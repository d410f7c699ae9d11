// Test harness (tests/test_text_host.py): the text front-end's scanning routines (llmvox_b200/csrc/text_kernels.cuh, the same source
// the device kernel runs) compiled for the HOST, so that the CPU suite can compare them with Python's regexes over thousands of
// random strings.  Not part of the product library.
#include <cstdint>
#include <vector>

#include "../llmvox_b200/csrc/text_kernels.cuh"

extern "C" int tx_host_ids(const uint8_t* bytes, int len, int clean, int32_t* out, int room) {
  const int cap = 12 * len + 64;
  std::vector<uint8_t> a(cap), b(cap);
  const uint8_t* s = bytes;
  int n = len;
  if (clean) {
    n = lvx::tx_clean(bytes, len, a.data(), b.data(), cap, &s);
    if (n < 0) return -1;
  }
  lvx::TxIdSink sink{out, room, 0};
  lvx::tx_tokenize(s, n, sink);
  return sink.m;
}
