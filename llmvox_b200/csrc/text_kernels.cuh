// Text front-end on the device: the reference's clean_text (streaming_server.py:106-149) and the
// ByT5 tokenisation of a whole sentence as its producer / generator threads feed it (text split at spaces, every word
// stripped, utf-8 byte + 3, </s> = 1 after every word, 385 after the last: streaming_server.py:184-248, 305-310), from raw
// UTF-8 bytes straight into the sessions' device-side text ids -- one H2D copy and one launch for a batch of sentences
// instead of host regexes and a Python list of ids per sentence.
//
// clean_text is a chain of 13 rewrite passes; every pass is a left-to-right scan, so ONE thread (lane 0 of the sentence's warp)
// runs a sentence through them between two scratch buffers (sentences are a few hundred bytes; all sentences of a batch run in
// parallel).  The
// character classes are Python's: `\s` / str.strip = str.isspace (29 code points), `\d` = Unicode category Nd (64 runs of
// ten digits, Unicode 15); bytes are decoded as UTF-8 where a class is tested.  tests/test_gpu_text.py compares the ids
// with the host pipeline (protocol.clean_text + tokenizer.sentence_ids, themselves pinned to the reference's fixtures).
#pragma once
#include "decode_kernels.cuh"

namespace lvx {

struct TxRange { uint32_t lo, hi; };
#define TX_DIGIT_RUNS {{0x30, 0x39}, {0x660, 0x669}, {0x6F0, 0x6F9}, {0x7C0, 0x7C9}, {0x966, 0x96F}, {0x9E6, 0x9EF}, {0xA66, 0xA6F}, {0xAE6, 0xAEF}, {0xB66, 0xB6F}, {0xBE6, 0xBEF}, {0xC66, 0xC6F}, {0xCE6, 0xCEF}, {0xD66, 0xD6F}, {0xDE6, 0xDEF}, {0xE50, 0xE59}, {0xED0, 0xED9}, {0xF20, 0xF29}, {0x1040, 0x1049}, {0x1090, 0x1099}, {0x17E0, 0x17E9}, {0x1810, 0x1819}, {0x1946, 0x194F}, {0x19D0, 0x19D9}, {0x1A80, 0x1A89}, {0x1A90, 0x1A99}, {0x1B50, 0x1B59}, {0x1BB0, 0x1BB9}, {0x1C40, 0x1C49}, {0x1C50, 0x1C59}, {0xA620, 0xA629}, {0xA8D0, 0xA8D9}, {0xA900, 0xA909}, {0xA9D0, 0xA9D9}, {0xA9F0, 0xA9F9}, {0xAA50, 0xAA59}, {0xABF0, 0xABF9}, {0xFF10, 0xFF19}, {0x104A0, 0x104A9}, {0x10D30, 0x10D39}, {0x11066, 0x1106F}, {0x110F0, 0x110F9}, {0x11136, 0x1113F}, {0x111D0, 0x111D9}, {0x112F0, 0x112F9}, {0x11450, 0x11459}, {0x114D0, 0x114D9}, {0x11650, 0x11659}, {0x116C0, 0x116C9}, {0x11730, 0x11739}, {0x118E0, 0x118E9}, {0x11950, 0x11959}, {0x11C50, 0x11C59}, {0x11D50, 0x11D59}, {0x11DA0, 0x11DA9}, {0x11F50, 0x11F59}, {0x16A60, 0x16A69}, {0x16AC0, 0x16AC9}, {0x16B50, 0x16B59}, {0x1D7CE, 0x1D7FF}, {0x1E140, 0x1E149}, {0x1E2F0, 0x1E2F9}, {0x1E4F0, 0x1E4F9}, {0x1E950, 0x1E959}, {0x1FBF0, 0x1FBF9}}
__device__ const TxRange tx_digit_runs[64] = TX_DIGIT_RUNS;
// The scanning routines below are __host__ __device__ so that tests/tx_host_harness.cu can run the very same source on the CPU
// against Python's regexes over thousands of random strings (a test harness: the product library exports no host path).
static const TxRange tx_digit_runs_host[64] = TX_DIGIT_RUNS;
#define TX_HD __host__ __device__ __forceinline__

TX_HD bool tx_is_space(uint32_t c) {
  return (c >= 0x9 && c <= 0xD) || (c >= 0x1C && c <= 0x20) || c == 0x85 || c == 0xA0 || c == 0x1680 || (c >= 0x2000 && c <= 0x200A) ||
         c == 0x2028 || c == 0x2029 || c == 0x202F || c == 0x205F || c == 0x3000;
}
TX_HD bool tx_is_digit(uint32_t c) {
  if (c < 0x80) return c >= '0' && c <= '9';
#ifdef __CUDA_ARCH__
  const TxRange* runs = tx_digit_runs;
#else
  const TxRange* runs = tx_digit_runs_host;
#endif
  int lo = 1, hi = 63;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    if (c < runs[mid].lo) hi = mid - 1;
    else if (c > runs[mid].hi) lo = mid + 1;
    else return true;
  }
  return false;
}
// code point at byte i of s[0, n) and its length in bytes (invalid sequences: the byte itself, length 1)
TX_HD uint32_t tx_cp(const uint8_t* s, int n, int i, int& len) {
  const uint32_t b = s[i];
  len = 1;
  if (b < 0x80) return b;
  if ((b & 0xE0) == 0xC0 && i + 1 < n) { len = 2; return ((b & 0x1F) << 6) | (s[i + 1] & 0x3F); }
  if ((b & 0xF0) == 0xE0 && i + 2 < n) { len = 3; return ((b & 0x0F) << 12) | ((s[i + 1] & 0x3F) << 6) | (s[i + 2] & 0x3F); }
  if ((b & 0xF8) == 0xF0 && i + 3 < n) { len = 4; return ((b & 0x07) << 18) | ((s[i + 1] & 0x3F) << 12) | ((s[i + 2] & 0x3F) << 6) | (s[i + 3] & 0x3F); }
  return 0xFFFD0000u | b;   // not a class member
}
// code point that ENDS at byte i (exclusive): start index of the last code point of s[0, i)
TX_HD int tx_prev(const uint8_t* s, int i) {
  int j = i - 1;
  while (j > 0 && (s[j] & 0xC0) == 0x80 && i - j < 4) --j;
  return j;
}
struct TxOut {
  uint8_t* p;
  int n, cap;
  TX_HD void put(uint8_t b) { if (n < cap) p[n] = b; ++n; }
  TX_HD void put(const char* lit) { for (; *lit; ++lit) put((uint8_t)*lit); }
  TX_HD void copy(const uint8_t* s, int i, int len) { for (int k = 0; k < len; ++k) put(s[i + k]); }
};

// clean_text (streaming_server.py:120-149) of s[0, n) between the two scratch buffers a / b (capacity cap each); returns the
// result's buffer and length.  Overflow of cap (an input made of backslashes grows 11x) is reported as a negative length.
__host__ __device__ inline int tx_clean(const uint8_t* s, int n, uint8_t* a, uint8_t* b, int cap, const uint8_t** res) {
  int len;
  // text.strip()
  int lo = 0, hi = n;
  while (lo < hi) { const uint32_t c = tx_cp(s, hi, lo, len); if (!tx_is_space(c)) break; lo += len; }
  while (hi > lo) { const int j = tx_prev(s + lo, hi - lo) + lo; const uint32_t c = tx_cp(s, hi, j, len); if (!tx_is_space(c) || j + len != hi) break; hi = j; }
  TxOut o{a, 0, cap};
  // .replace("**", "").replace("-", " ")
  for (int i = lo; i < hi;) {
    if (s[i] == '*' && i + 1 < hi && s[i + 1] == '*') { i += 2; continue; }
    o.put(s[i] == '-' ? (uint8_t)' ' : s[i]);
    ++i;
  }
  if (o.n > cap) return -1;
  const uint8_t* src = a;
  n = o.n;
  uint8_t* dst = b;
  auto flip = [&](TxOut& w) { src = w.p; n = w.n; dst = (w.p == a) ? b : a; };
  {  // re.sub(r'(\d)\.(?=\s|$)', r'\1'): a '.' after a digit and before whitespace / the end goes
    TxOut w{dst, 0, cap};
    bool prev_digit = false;
    for (int i = 0; i < n;) {
      const uint32_t c = tx_cp(src, n, i, len);
      if (c == '.' && prev_digit) {
        int l2;
        if (i + 1 >= n || tx_is_space(tx_cp(src, n, i + 1, l2))) { prev_digit = false; i += 1; continue; }
      }
      prev_digit = tx_is_digit(c);
      w.copy(src, i, len);
      i += len;
    }
    if (w.n > cap) return -1;
    flip(w);
  }
  {  // '*' removed; '#', '&', '@' spelled out
    TxOut w{dst, 0, cap};
    for (int i = 0; i < n; ++i) {
      const uint8_t c = src[i];
      if (c == '*') continue;
      if (c == '#') w.put(" number ");
      else if (c == '&') w.put(" and ");
      else if (c == '@') w.put(" at ");
      else w.put(c);
    }
    if (w.n > cap) return -1;
    flip(w);
  }
  {  // re.sub(r'\s+', ' ')
    TxOut w{dst, 0, cap};
    bool in_space = false;
    for (int i = 0; i < n;) {
      const uint32_t c = tx_cp(src, n, i, len);
      if (tx_is_space(c)) {
        if (!in_space) w.put((uint8_t)' ');
        in_space = true;
      } else {
        in_space = false;
        w.copy(src, i, len);
      }
      i += len;
    }
    if (w.n > cap) return -1;
    flip(w);
  }
  {  // re.sub(r'\.{3,}', ' pause ')
    TxOut w{dst, 0, cap};
    for (int i = 0; i < n;) {
      if (src[i] == '.') {
        int j = i;
        while (j < n && src[j] == '.') ++j;
        if (j - i >= 3) w.put(" pause ");
        else w.copy(src, i, j - i);
        i = j;
      } else {
        w.put(src[i]);
        ++i;
      }
    }
    if (w.n > cap) return -1;
    flip(w);
  }
  {  // re.sub(r'(\d),(\d)', r'\1\2'): non-overlapping matches, left to right ("1,2,3" -> "12,3")
    TxOut w{dst, 0, cap};
    for (int i = 0; i < n;) {
      const uint32_t c = tx_cp(src, n, i, len);
      if (tx_is_digit(c) && i + len < n && src[i + len] == ',' && i + len + 1 < n) {
        int l2;
        const uint32_t d = tx_cp(src, n, i + len + 1, l2);
        if (tx_is_digit(d)) {
          w.copy(src, i, len);
          w.copy(src, i + len + 1, l2);
          i += len + 1 + l2;
          continue;
        }
      }
      w.copy(src, i, len);
      i += len;
    }
    if (w.n > cap) return -1;
    flip(w);
  }
  {  // re.sub(r'\/+', ' slash '), re.sub(r'\\+', ' backslash ')
    TxOut w{dst, 0, cap};
    for (int i = 0; i < n;) {
      const uint8_t c = src[i];
      if (c == '/' || c == '\\') {
        int j = i;
        while (j < n && src[j] == c) ++j;
        w.put(c == '/' ? " slash " : " backslash ");
        i = j;
      } else {
        w.put(c);
        ++i;
      }
    }
    if (w.n > cap) return -1;
    flip(w);
  }
  *res = src;
  return n;
}

// Sentence tokenisation (streaming_server.py:184-248, 305-310) of s[0, len): sentence.strip().split(" ") with empty pieces dropped,
// every word stripped and byte-tokenised ("[PAD]" -> 384 and "EOS" -> 385 as literal substrings, else utf-8 byte + 3), </s> = 1
// after every word, 385 after the last.  sink(id) receives the ids in order.
template <typename Sink>
__host__ __device__ inline void tx_tokenize(const uint8_t* s, int len, Sink& sink) {
  int l2;
  int lo = 0, hi = len;
  while (lo < hi) { const uint32_t c = tx_cp(s, hi, lo, l2); if (!tx_is_space(c)) break; lo += l2; }
  while (hi > lo) { const int j = tx_prev(s + lo, hi - lo) + lo; const uint32_t c = tx_cp(s, hi, j, l2); if (!tx_is_space(c) || j + l2 != hi) break; hi = j; }
  bool any = false;
  for (int p = lo; p < hi;) {
    int e = p;
    while (e < hi && s[e] != ' ') ++e;
    // word = s[p, e).strip()
    int a = p, b = e;
    while (a < b) { const uint32_t c = tx_cp(s, b, a, l2); if (!tx_is_space(c)) break; a += l2; }
    while (b > a) { const int j = tx_prev(s + a, b - a) + a; const uint32_t c = tx_cp(s, b, j, l2); if (!tx_is_space(c) || j + l2 != b) break; b = j; }
    if (e > p) {   // a non-empty piece between spaces is a word, even if it strips to nothing (then it is just </s>)
      for (int k = a; k < b;) {
        if (b - k >= 5 && s[k] == '[' && s[k + 1] == 'P' && s[k + 2] == 'A' && s[k + 3] == 'D' && s[k + 4] == ']') { sink(384); k += 5; }
        else if (b - k >= 3 && s[k] == 'E' && s[k + 1] == 'O' && s[k + 2] == 'S') { sink(385); k += 3; }
        else { sink((int)s[k] + 3); ++k; }
      }
      sink(1);
      any = true;
    }
    p = e + 1;
  }
  if (any) sink(385);
}
// ids into a session's text buffer; counts past the capacity without writing
struct TxIdSink {
  int* ids;
  int room, m;
  __host__ __device__ void operator()(int id) { if (m < room) ids[m] = id; ++m; }
};

// One sentence per warp: [clean_text,] split at ' ', strip every word, tokenise ("[PAD]" -> 384 and "EOS" -> 385 as literal
// substrings, else utf-8 byte + 3), </s> after every word, 385 after the last; ids to the slot's text buffer.
// status[i]: number of ids written, or -1 (scratch overflow) / -2 (more ids than max_context).
// The rewrite passes read what the previous pass wrote, byte by byte: between two GLOBAL scratch buffers every such read is an
// L1 miss, and threads of a warp working on different sentences diverge at every byte.  So a sentence gets a WARP: its lanes copy
// the raw bytes into shared memory together, then lane 0 runs the passes between the warp's two shared-memory buffers (2 x TX_SCAP
// bytes) and tokenises; only a sentence that outgrows them is redone in the global buffers.
constexpr int TX_WARPS = 4;
constexpr int TX_SCAP = 1024;
__global__ void __launch_bounds__(32 * TX_WARPS) text_frontend_kernel(const uint8_t* __restrict__ bytes, const int* __restrict__ offs,
                                                                      const int* __restrict__ slots, int n, int clean,
                                                                      uint8_t* __restrict__ scratch, int cap, SessionState st,
                                                                      int* __restrict__ status) {
  __shared__ __align__(16) uint8_t tx_sm[TX_WARPS * 2 * TX_SCAP];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * TX_WARPS + warp;
  if (i >= n) return;
  const uint8_t* s = bytes + offs[i];
  int len = offs[i + 1] - offs[i];
  uint8_t* sa = tx_sm + warp * (2 * TX_SCAP);
  uint8_t* sb = sa + TX_SCAP;
  const uint8_t* raw = s;
  if (len <= TX_SCAP) {
    for (int k = lane; k < len; k += 32) sb[k] = raw[k];
    __syncwarp();
    s = sb;
  }
  if (lane != 0) return;
  if (clean) {
    const int raw_len = len;
    len = raw_len <= TX_SCAP ? tx_clean(sb, raw_len, sa, sb, TX_SCAP, &s) : -1;   // pass 1 reads sb and writes sa; sb is free from pass 2 on
    if (len < 0) len = tx_clean(raw, raw_len, scratch + (size_t)(2 * i) * cap, scratch + (size_t)(2 * i + 1) * cap, cap, &s);
    if (len < 0) { status[i] = -1; return; }
  }
  const int slot = slots[i];
  const int base = st.text_len[slot];   // appended after the text the session already holds (like lvx_feed_text)
  TxIdSink sink{st.text_ids + (size_t)slot * st.max_context + base, st.max_context - base, 0};
  tx_tokenize(s, len, sink);
  if (sink.m > sink.room) { status[i] = -2; return; }
  st.text_len[slot] = base + sink.m;
  status[i] = sink.m;
}

}  // namespace lvx
