// sbimg.cpp — native decoders for the image files either side of the path.
//
// The reference reads every frame and mask with cv::imread (CStereoMatching.cpp:147-151 colour, mask with
// CV_LOAD_IMAGE_GRAYSCALE; CManageData.cpp:68 for the original size); its data sets are JPEG files
// ("%.4d_Cam%d.jpg", BatchProcess/main.cpp:66).  OpenCV's codecs are third-party code that is not under /root/reference
// (OpenCV 2.4.5 wraps libjpeg / libpng), so this restates the published algorithms and is pinned against the OpenCV
// 4.13 build of this image (libjpeg-turbo, libpng): tests/test_host_decode.py compares bit for bit.
//   JPEG  baseline / extended sequential and progressive Huffman, 8 bit, 1 or 3 components, any scan layout, restart intervals.
//         Inverse DCT = the "islow" integer LL&M transform, chroma by "fancy" (triangle) upsampling, YCbCr -> RGB by the
//         16-bit fixed-point tables: libjpeg's default decompression path.  Grey output of a colour file = the Y plane
//         (what cv::imread(..., 0) asks libjpeg for).  Arithmetic-coded, lossless and hierarchical files are refused.
//   PNG   all colour types, bit depths 1-16, Adam7 interlace; 16 -> 8 bit by dropping the low byte, alpha dropped, grey from
//         RGB by libpng's 15-bit coefficients for the 0.299 / 0.587 OpenCV passes.
//   BMP   uncompressed 24 / 32 bit.       PNM   P5 / P6 (sbcv.cpp).
// EXIF orientation is ignored (as OpenCV 2.4.5 did).

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "sbcv.h"

namespace sbcv {
namespace {

// =================================================================================================== JPEG
const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                             41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                             30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct Huff {
  bool present = false;
  uint8_t bits[17] = {0};
  uint8_t vals[256] = {0};
  // canonical decoding tables (ITU T.81 F.2.2.3)
  int mincode[17], maxcode[18], valptr[17];
  uint16_t fast[1 << 11];  // 11-bit look-ahead: (length << 8) | symbol, 0 = longer code
  // AC tables only: the same 11 bits resolved down to the coefficient when code + value bits fit in them:
  // (value << 16) | (kind << 12) | (run << 8) | (code length + value bits); kind 1 = coefficient, 2 = end of block, 3 = ZRL
  // (sixteen zeros), 0 = not resolved here
  int32_t fast_ac[1 << 11];
  // false: the code-length counts over-subscribe the code space (no prefix code has them; libjpeg's jdhuff.c
  // refuses such a table too) - building the look-ahead table from it would index past its end
  bool build() {
    int code = 0, k = 0;
    for (int l = 1; l <= 16; l++) {
      if (code + bits[l] > (1 << l)) return false;
      valptr[l] = k;
      mincode[l] = code;
      code += bits[l];
      k += bits[l];
      maxcode[l] = bits[l] ? code - 1 : -1;
      code <<= 1;
    }
    maxcode[17] = 0x7fffffff;
    memset(fast, 0, sizeof fast);
    code = 0; k = 0;
    for (int l = 1; l <= 11; l++) {
      for (int i = 0; i < bits[l]; i++, k++, code++) {
        const int lo = code << (11 - l);
        for (int j = 0; j < (1 << (11 - l)); j++) fast[lo + j] = (uint16_t)((l << 8) | vals[k]);
      }
      code <<= 1;
    }
    for (int i = 0; i < (1 << 11); i++) {
      fast_ac[i] = 0;
      const int f = fast[i];
      if (!f) continue;
      const int len = f >> 8, run = (f >> 4) & 15, size = f & 15;
      if (size == 0) {
        if (run == 0) fast_ac[i] = (2 << 12) | len;
        else if (run == 15) fast_ac[i] = (3 << 12) | len;
        else fast_ac[i] = (2 << 12) | len;  // an undefined symbol in a sequential scan: ends the block like EOB (decode_block)
        continue;
      }
      if (len + size > 11) continue;
      const int bits = (i >> (11 - len - size)) & ((1 << size) - 1);
      const int v = bits + ((((bits >> (size - 1)) & 1) - 1) & (1 - (1 << size)));  // EXTEND
      fast_ac[i] = (int32_t)((uint32_t)v << 16) | (1 << 12) | (run << 8) | (len + size);
    }
    return true;
  }
};

// The entropy-coded data of one scan (T.81 B.1.1.5), handed to the bit reader one restart interval at a time with the stuffed
// zero bytes removed and eight zero bytes appended - so the reader never looks for 0xFF and may always load 8 bytes.  A piece
// ends at the first marker of any kind; what a decoder that runs into it sees is zeros (jdhuff.c's "insert_fake_zeros").
struct EntropySegment {
  const uint8_t* end = nullptr;   // end of the file
  const uint8_t* stop = nullptr;  // the 0xFF of the marker that ended the current piece (or `end`)
  std::vector<uint8_t> buf;
  size_t len = 0;                 // bytes of the current piece in buf (the padding follows)
  void load(const uint8_t* p) {
    buf.clear();
    for (;;) {
      const uint8_t* f = p < end ? (const uint8_t*)memchr(p, 0xFF, (size_t)(end - p)) : nullptr;
      if (!f || f + 1 >= end) {  // ran off the file; a lone trailing 0xFF is not data
        buf.insert(buf.end(), p, f ? f : end);
        stop = end;
        break;
      }
      buf.insert(buf.end(), p, f);
      if (f[1] != 0x00) { stop = f; break; }
      buf.push_back(0xFF);
      p = f + 2;
    }
    len = buf.size();
    buf.insert(buf.end(), 8, (uint8_t)0);
  }
  void start(const uint8_t* p, const uint8_t* file_end) { end = file_end; load(p); }
  // the piece after the next RSTn marker, wherever that is (anything in between is skipped, as a resynchronising decoder does)
  bool next_restart() {
    const uint8_t* q = stop;
    while (q + 1 < end && !(q[0] == 0xFF && q[1] >= 0xD0 && q[1] <= 0xD7)) q++;
    if (q + 1 >= end) return false;
    load(q + 2);
    return true;
  }
  // where the marker parser goes on after the scan: the first marker that is neither a stuffed zero nor RSTn
  const uint8_t* after() const {
    const uint8_t* q = stop;
    while (q + 1 < end && !(q[0] == 0xFF && q[1] != 0x00 && !(q[1] >= 0xD0 && q[1] <= 0xD7))) q++;
    return q;
  }
};

// MSB-first reader over one piece of an EntropySegment.  `n` counts the valid bits at the top of `acc` (56..63 after a refill);
// bits below them are a preview of the bytes at `p` - or zeros past the end of the piece, which is what a decoder that ran into
// a marker is fed (jdhuff.c's "insert_fake_zeros").
struct BitReader {
  const uint8_t* p = nullptr;
  const uint8_t* end = nullptr;
  uint64_t acc = 0;
  int n = 0;
  void open(const EntropySegment& e) {
    p = e.buf.data();
    end = p + e.len;
    acc = 0;
    n = 0;
  }
  void fill() {
    if (p < end) {
      uint64_t w;
      memcpy(&w, p, 8);
      acc |= __builtin_bswap64(w) >> n;
      p += (63 - n) >> 3;
    }
    n |= 56;
  }
  int peek(int k) { if (n < k) fill(); return (int)(acc >> (64 - k)); }
  void skip(int k) { acc <<= k; n -= k; }
  int get(int k) { if (k == 0) return 0; const int v = peek(k); skip(k); return v; }
};

inline int huff_decode(BitReader& br, const Huff& h) {
  const int look = br.peek(16);
  const uint16_t f = h.fast[look >> 5];
  if (f) { br.skip(f >> 8); return f & 255; }
  for (int l = 12; l <= 16; l++) {
    const int code = look >> (16 - l);
    if (h.maxcode[l] >= 0 && code <= h.maxcode[l] && code >= h.mincode[l]) {
      br.skip(l);
      return h.vals[h.valptr[l] + code - h.mincode[l]];
    }
  }
  br.skip(16);
  return -1;
}
// T.81 F.2.2.1 EXTEND, branch-free: values whose top bit is clear are negative (v - 2^s + 1)
inline int extend(int v, int s) { return v + ((((v >> (s - 1)) & 1) - 1) & (1 - (1 << s))); }

struct Comp {
  int id = 0, h = 1, v = 1, tq = 0, td = 0, ta = 0;
  int bw = 0, bh = 0;     // blocks allocated (padded to whole MCUs)
  int dw = 0, dh = 0;     // downsampled_width / height (real samples)
  int pitch = 0;
  std::vector<uint8_t> plane;
  int pred = 0;
  std::vector<int16_t> coefs;  // progressive files: every block's 64 coefficients (natural order), filled scan by scan
  bool needed = true;  // false: entropy-decoded only (chroma of a colour file read as grey: jpeg_component_info.component_needed)
};

// jidctint.c (jpeg_idct_islow): CONST_BITS 13, PASS1_BITS 2
inline int64_t descale(int64_t x, int n) { return (x + ((int64_t)1 << (n - 1))) >> n; }
inline uint8_t range_limit(int64_t x) {  // range_limit[(x) & RANGE_MASK] with the table centred on 128
  const int i = (int)(x & 1023);
  if (i < 128) return (uint8_t)(i + 128);
  if (i < 512) return 255;
  if (i < 896) return 0;
  return (uint8_t)(i - 896);
}
void idct_islow(const int16_t* coef, const uint16_t* q, uint8_t* out, int pitch) {
  enum { F0_298 = 2446, F0_390 = 3196, F0_541 = 4433, F0_765 = 6270, F0_899 = 7373, F1_175 = 9633, F1_501 = 12299, F1_847 = 15137,
         F1_961 = 16069, F2_053 = 16819, F2_562 = 20995, F3_072 = 25172 };
  int64_t ws[64];
  for (int c = 0; c < 8; c++) {
    int64_t in[8];
    for (int r = 0; r < 8; r++) in[r] = (int64_t)coef[r * 8 + c] * q[r * 8 + c];
    if ((in[1] | in[2] | in[3] | in[4] | in[5] | in[6] | in[7]) == 0) {
      const int64_t dc = in[0] * 4;
      for (int r = 0; r < 8; r++) ws[r * 8 + c] = dc;
      continue;
    }
    int64_t z2 = in[2], z3 = in[6];
    int64_t z1 = (z2 + z3) * F0_541;
    int64_t tmp2 = z1 + z3 * -(int64_t)F1_847, tmp3 = z1 + z2 * F0_765;
    z2 = in[0]; z3 = in[4];
    int64_t tmp0 = (z2 + z3) * 8192, tmp1 = (z2 - z3) * 8192;
    const int64_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = in[7]; tmp1 = in[5]; tmp2 = in[3]; tmp3 = in[1];
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int64_t z4 = tmp1 + tmp3;
    const int64_t z5 = (z3 + z4) * F1_175;
    tmp0 *= F0_298; tmp1 *= F2_053; tmp2 *= F3_072; tmp3 *= F1_501;
    z1 *= -(int64_t)F0_899; z2 *= -(int64_t)F2_562; z3 *= -(int64_t)F1_961; z4 *= -(int64_t)F0_390;
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    ws[0 * 8 + c] = descale(tmp10 + tmp3, 11); ws[7 * 8 + c] = descale(tmp10 - tmp3, 11);
    ws[1 * 8 + c] = descale(tmp11 + tmp2, 11); ws[6 * 8 + c] = descale(tmp11 - tmp2, 11);
    ws[2 * 8 + c] = descale(tmp12 + tmp1, 11); ws[5 * 8 + c] = descale(tmp12 - tmp1, 11);
    ws[3 * 8 + c] = descale(tmp13 + tmp0, 11); ws[4 * 8 + c] = descale(tmp13 - tmp0, 11);
  }
  for (int r = 0; r < 8; r++) {
    const int64_t* w = ws + r * 8;
    uint8_t* o = out + (size_t)r * pitch;
    if ((w[1] | w[2] | w[3] | w[4] | w[5] | w[6] | w[7]) == 0) {  // jidctint.c's zero-AC row: the same value as the full pass
      memset(o, range_limit(descale(w[0], 5)), 8);
      continue;
    }
    int64_t z2 = w[2], z3 = w[6];
    int64_t z1 = (z2 + z3) * F0_541;
    int64_t tmp2 = z1 + z3 * -(int64_t)F1_847, tmp3 = z1 + z2 * F0_765;
    int64_t tmp0 = (w[0] + w[4]) * 8192, tmp1 = (w[0] - w[4]) * 8192;
    const int64_t tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
    tmp0 = w[7]; tmp1 = w[5]; tmp2 = w[3]; tmp3 = w[1];
    z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
    int64_t z4 = tmp1 + tmp3;
    const int64_t z5 = (z3 + z4) * F1_175;
    tmp0 *= F0_298; tmp1 *= F2_053; tmp2 *= F3_072; tmp3 *= F1_501;
    z1 *= -(int64_t)F0_899; z2 *= -(int64_t)F2_562; z3 *= -(int64_t)F1_961; z4 *= -(int64_t)F0_390;
    z3 += z5; z4 += z5;
    tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
    o[0] = range_limit(descale(tmp10 + tmp3, 18)); o[7] = range_limit(descale(tmp10 - tmp3, 18));
    o[1] = range_limit(descale(tmp11 + tmp2, 18)); o[6] = range_limit(descale(tmp11 - tmp2, 18));
    o[2] = range_limit(descale(tmp12 + tmp1, 18)); o[5] = range_limit(descale(tmp12 - tmp1, 18));
    o[3] = range_limit(descale(tmp13 + tmp0, 18)); o[4] = range_limit(descale(tmp13 - tmp0, 18));
  }
}


#if defined(__SSE2__)
// SSE2 twin of idct_islow: the same integer algebra regrouped into pmaddwd pairs (z1 = (a + b) * K; t = z1 + a * L  ==
// a * (K + L) + b * K, exact in integers), eight columns per register, 16-bit lanes between the passes.  It is exact only while
// every dequantised value and every pass-1 output fits 16 bits and no 32-bit sum wraps; the guard is the block's
// L1 = sum |coef * q| <= 5600:  |pass-1 output| <= 11363 * L1(column) / 2048 + 1 (11363 = the largest weight of any input in any
// output of one pass), so a row of pass-1 outputs sums to at most 5.55 * L1 + 8 = 31088 < 2^15, and every 32-bit partial sum of
// pass 2 stays below 32768 * 31088 < 2^31.  Blocks outside the guard (and tables with entries above 255) take idct_islow.
inline __m128i pair16(int a, int b) { return _mm_set1_epi32((int)(((uint32_t)(uint16_t)(int16_t)b << 16) | (uint16_t)(int16_t)a)); }

template <int SHIFT, bool FINAL>
inline void pass8(__m128i v[8]) {
  const __m128i c26a = pair16(10703, 4433), c26b = pair16(4433, -10704);
  const __m128i c04a = pair16(8192, 8192), c04b = pair16(8192, -8192);
  const __m128i c34a = pair16(-6436, 9633), c34b = pair16(9633, 6437);
  const __m128i c71a = pair16(-4927, -7373), c71b = pair16(-7373, 4926);
  const __m128i c53a = pair16(-4176, -20995), c53b = pair16(-20995, 4177);
  const __m128i rnd = _mm_set1_epi32(1 << (SHIFT - 1));
  const __m128i z3 = _mm_add_epi16(v[7], v[3]), z4 = _mm_add_epi16(v[5], v[1]);
  __m128i o[8][2];
  for (int h = 0; h < 2; h++) {
#define UNPK(a, b) (h ? _mm_unpackhi_epi16(a, b) : _mm_unpacklo_epi16(a, b))
    const __m128i a26 = UNPK(v[2], v[6]), a04 = UNPK(v[0], v[4]), a34 = UNPK(z3, z4), a71 = UNPK(v[7], v[1]), a53 = UNPK(v[5], v[3]);
#undef UNPK
    const __m128i e3 = _mm_madd_epi16(a26, c26a), e2 = _mm_madd_epi16(a26, c26b);
    const __m128i e0 = _mm_add_epi32(_mm_madd_epi16(a04, c04a), rnd), e1 = _mm_add_epi32(_mm_madd_epi16(a04, c04b), rnd);
    const __m128i t10 = _mm_add_epi32(e0, e3), t13 = _mm_sub_epi32(e0, e3), t11 = _mm_add_epi32(e1, e2), t12 = _mm_sub_epi32(e1, e2);
    const __m128i y3 = _mm_madd_epi16(a34, c34a), y4 = _mm_madd_epi16(a34, c34b);
    const __m128i o0 = _mm_add_epi32(_mm_madd_epi16(a71, c71a), y3), o3 = _mm_add_epi32(_mm_madd_epi16(a71, c71b), y4);
    const __m128i o1 = _mm_add_epi32(_mm_madd_epi16(a53, c53a), y4), o2 = _mm_add_epi32(_mm_madd_epi16(a53, c53b), y3);
    o[0][h] = _mm_srai_epi32(_mm_add_epi32(t10, o3), SHIFT); o[7][h] = _mm_srai_epi32(_mm_sub_epi32(t10, o3), SHIFT);
    o[1][h] = _mm_srai_epi32(_mm_add_epi32(t11, o2), SHIFT); o[6][h] = _mm_srai_epi32(_mm_sub_epi32(t11, o2), SHIFT);
    o[2][h] = _mm_srai_epi32(_mm_add_epi32(t12, o1), SHIFT); o[5][h] = _mm_srai_epi32(_mm_sub_epi32(t12, o1), SHIFT);
    o[3][h] = _mm_srai_epi32(_mm_add_epi32(t13, o0), SHIFT); o[4][h] = _mm_srai_epi32(_mm_sub_epi32(t13, o0), SHIFT);
  }
  for (int k = 0; k < 8; k++) {
    if (FINAL) {  // range_limit[x & 1023]: the 10-bit wrap first, then the clamp (done by the caller's packuswb after + 128)
      const __m128i m = _mm_set1_epi32(1023), b = _mm_set1_epi32(512);
      o[k][0] = _mm_sub_epi32(_mm_xor_si128(_mm_and_si128(o[k][0], m), b), b);
      o[k][1] = _mm_sub_epi32(_mm_xor_si128(_mm_and_si128(o[k][1], m), b), b);
    }
    v[k] = _mm_packs_epi32(o[k][0], o[k][1]);
  }
}

inline void transpose8(__m128i v[8]) {
  const __m128i a0 = _mm_unpacklo_epi16(v[0], v[1]), a1 = _mm_unpackhi_epi16(v[0], v[1]);
  const __m128i a2 = _mm_unpacklo_epi16(v[2], v[3]), a3 = _mm_unpackhi_epi16(v[2], v[3]);
  const __m128i a4 = _mm_unpacklo_epi16(v[4], v[5]), a5 = _mm_unpackhi_epi16(v[4], v[5]);
  const __m128i a6 = _mm_unpacklo_epi16(v[6], v[7]), a7 = _mm_unpackhi_epi16(v[6], v[7]);
  const __m128i b0 = _mm_unpacklo_epi32(a0, a2), b1 = _mm_unpackhi_epi32(a0, a2);
  const __m128i b2 = _mm_unpacklo_epi32(a1, a3), b3 = _mm_unpackhi_epi32(a1, a3);
  const __m128i b4 = _mm_unpacklo_epi32(a4, a6), b5 = _mm_unpackhi_epi32(a4, a6);
  const __m128i b6 = _mm_unpacklo_epi32(a5, a7), b7 = _mm_unpackhi_epi32(a5, a7);
  v[0] = _mm_unpacklo_epi64(b0, b4); v[1] = _mm_unpackhi_epi64(b0, b4);
  v[2] = _mm_unpacklo_epi64(b1, b5); v[3] = _mm_unpackhi_epi64(b1, b5);
  v[4] = _mm_unpacklo_epi64(b2, b6); v[5] = _mm_unpackhi_epi64(b2, b6);
  v[6] = _mm_unpacklo_epi64(b3, b7); v[7] = _mm_unpackhi_epi64(b3, b7);
}

// false: the block is outside the range in which 16-bit lanes / 32-bit sums are exact -> caller takes the int64 path.
// q16: the quantisation table as int16 (the caller guarantees every entry <= 255).
inline bool idct_islow_sse2(const int16_t* coef, const int16_t* q16, uint8_t* out, int pitch) {
  __m128i v[8], l1 = _mm_setzero_si128();
  for (int r = 0; r < 8; r++) {
    const __m128i c = _mm_loadu_si128((const __m128i*)(coef + 8 * r)), q = _mm_loadu_si128((const __m128i*)(q16 + 8 * r));
    const __m128i a = _mm_max_epi16(c, _mm_subs_epi16(_mm_setzero_si128(), c));
    l1 = _mm_add_epi32(l1, _mm_madd_epi16(a, q));
    v[r] = _mm_mullo_epi16(c, q);
  }
  l1 = _mm_add_epi32(l1, _mm_shuffle_epi32(l1, 0x4E));
  l1 = _mm_add_epi32(l1, _mm_shuffle_epi32(l1, 0xB1));
  if (_mm_cvtsi128_si32(l1) > 5600) return false;
  pass8<11, false>(v);
  transpose8(v);
  pass8<18, true>(v);
  transpose8(v);
  const __m128i c128 = _mm_set1_epi16(128);
  for (int r = 0; r < 8; r += 2) {
    const __m128i p = _mm_packus_epi16(_mm_add_epi16(v[r], c128), _mm_add_epi16(v[r + 1], c128));
    _mm_storel_epi64((__m128i*)(out + (size_t)r * pitch), p);
    _mm_storel_epi64((__m128i*)(out + (size_t)(r + 1) * pitch), _mm_unpackhi_epi64(p, p));
  }
  return true;
}
#endif

// Per-row loops of the output stage (jdsample.c / jdcolor.c arithmetic, see Jpeg::upsample_row / to_mat).  On x86-64 GCC builds
// an AVX2 clone next to the baseline one and picks at load time (ifunc): the 32-bit multiplies of the colour conversion only
// vectorise from SSE4.1 on, and the baseline build is plain SSE2.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define SB_ROW_CLONES __attribute__((target_clones("avx2", "default")))
#else
#define SB_ROW_CLONES
#endif
SB_ROW_CLONES void ycc_to_bgr_row(const uint8_t* __restrict__ p0, const uint8_t* __restrict__ p1, const uint8_t* __restrict__ p2,
                                  uint8_t* __restrict__ o, int W) {
  // jdcolor.c ycc_rgb_convert, SCALEBITS 16: the table entries written out as arithmetic (FIX(1.40200) = 91881,
  // FIX(1.77200) = 116130, FIX(0.71414) = 46802, FIX(0.34414) = 22554, ONE_HALF = 32768; arithmetic right shifts)
  for (int x = 0; x < W; x++) {
    const int yy = p0[x], cb = p1[x] - 128, cr = p2[x] - 128;
    int r = yy + ((91881 * cr + 32768) >> 16);
    int g = yy + ((-22554 * cb + 32768 - 46802 * cr) >> 16);
    int b = yy + ((116130 * cb + 32768) >> 16);
    r = r < 0 ? 0 : (r > 255 ? 255 : r);
    g = g < 0 ? 0 : (g > 255 ? 255 : g);
    b = b < 0 ? 0 : (b > 255 ? 255 : b);
    o[3 * x] = (uint8_t)b; o[3 * x + 1] = (uint8_t)g; o[3 * x + 2] = (uint8_t)r;
  }
}
SB_ROW_CLONES void h2v1_fancy_row(const uint8_t* __restrict__ in, uint8_t* __restrict__ dst, int n) {
  dst[0] = in[0];
  dst[1] = (uint8_t)((in[0] * 3 + in[1] + 2) >> 2);
  for (int i = 1; i < n - 1; i++) {
    const int v = in[i] * 3;
    dst[2 * i] = (uint8_t)((v + in[i - 1] + 1) >> 2);
    dst[2 * i + 1] = (uint8_t)((v + in[i + 1] + 2) >> 2);
  }
  dst[2 * n - 2] = (uint8_t)((in[n - 1] * 3 + in[n - 2] + 1) >> 2);
  dst[2 * n - 1] = in[n - 1];
}
SB_ROW_CLONES void h2v2_fancy_row(const uint8_t* __restrict__ in0, const uint8_t* __restrict__ in1, uint8_t* __restrict__ dst, int n) {
  dst[0] = (uint8_t)(((in0[0] * 3 + in1[0]) * 4 + 8) >> 4);
  dst[1] = (uint8_t)(((in0[0] * 3 + in1[0]) * 3 + (in0[1] * 3 + in1[1]) + 7) >> 4);
  for (int i = 1; i < n - 1; i++) {  // column sums recomputed per column: independent iterations (vectorisable)
    const int last = in0[i - 1] * 3 + in1[i - 1], cur = in0[i] * 3 + in1[i], next = in0[i + 1] * 3 + in1[i + 1];
    dst[2 * i] = (uint8_t)((cur * 3 + last + 8) >> 4);
    dst[2 * i + 1] = (uint8_t)((cur * 3 + next + 7) >> 4);
  }
  const int cur = in0[n - 1] * 3 + in1[n - 1], last = in0[n - 2] * 3 + in1[n - 2];
  dst[2 * n - 2] = (uint8_t)((cur * 3 + last + 8) >> 4);
  dst[2 * n - 1] = (uint8_t)((cur * 4 + 7) >> 4);
}
SB_ROW_CLONES void h1v2_fancy_row(const uint8_t* __restrict__ in0, const uint8_t* __restrict__ in1, uint8_t* __restrict__ dst, int W, int bias) {
  for (int x = 0; x < W; x++) dst[x] = (uint8_t)((in0[x] * 3 + in1[x] + bias) >> 2);
}
SB_ROW_CLONES void grey_to_bgr_row(const uint8_t* __restrict__ p0, uint8_t* __restrict__ o, int W) {
  for (int x = 0; x < W; x++) o[3 * x] = o[3 * x + 1] = o[3 * x + 2] = p0[x];
}

struct Jpeg {
  const uint8_t* d;
  size_t n;
  bool want_gray = false;  // only component 0 is reconstructed
  bool progressive = false;
  std::string err;
  int W = 0, H = 0, nc = 0, hmax = 1, vmax = 1, restart = 0;
  bool adobe = false;
  int adobe_transform = -1;
  Comp comp[3];
  uint16_t qt[4][64];
  int16_t qt16[4][64];  // the same table for the 16-bit-lane transform; usable when every entry is <= 255 (qt_small)
  bool qt_set[4] = {false, false, false, false}, qt_small[4] = {false, false, false, false};
  Huff hdc[4], hac[4];
  EntropySegment ent;  // the scan being decoded

  static int be16(const uint8_t* p) { return (p[0] << 8) | p[1]; }

  void idct(const int16_t* coef, int tq, uint8_t* out, int pitch) const {
#if defined(__SSE2__)
    static const bool scalar_only = getenv("SB200_JPEG_SCALAR") && atoi(getenv("SB200_JPEG_SCALAR")) != 0;  // tests: both transforms agree
    if (!scalar_only && qt_small[tq] && idct_islow_sse2(coef, qt16[tq], out, pitch)) return;
#endif
    idct_islow(coef, qt[tq], out, pitch);
  }

  // Returns the zigzag position of the last non-zero coefficient (0 = DC only), or -1 on corrupt data.  A code and the value
  // bits that follow it (at most 16 + 15 bits) are taken from one look at the 64-bit accumulator.
  int decode_block(BitReader& br, Comp& c, int16_t* coef) {
    memset(coef, 0, 64 * sizeof(int16_t));
    const Huff& dc = hdc[c.td];
    const Huff& ac = hac[c.ta];
    int s = huff_decode(br, dc);
    if (s < 0 || s > 15) return -1;
    const int diff = s ? extend(br.get(s), s) : 0;
    c.pred += diff;
    coef[0] = (int16_t)c.pred;
    int last = 0;
    // the reader lives in registers across the AC loop and is refilled before every symbol (no data-dependent branch: a
    // symbol and its value bits take at most 31 of the >= 56 bits a refill leaves)
    uint64_t acc = br.acc;
    int n = br.n;
    const uint8_t* bp = br.p;
    const uint8_t* const bend = br.end;
    for (int k = 1; k < 64;) {
      if (bp < bend) {
        uint64_t w;
        memcpy(&w, bp, 8);
        acc |= __builtin_bswap64(w) >> n;
        bp += (63 - n) >> 3;
      }
      n |= 56;
      const int look = (int)(acc >> 48);
      const int32_t fa = ac.fast_ac[look >> 5];
      if ((fa >> 12) & 3) {  // code and value bits resolved by one look-up
        const int len = fa & 255, kind = (fa >> 12) & 3;
        acc <<= len; n -= len;
        if (kind == 1) {
          k += (fa >> 8) & 15;
          if (k > 63) { br.acc = acc; br.n = n; br.p = bp; return -1; }
          coef[kZigzag[k]] = (int16_t)(fa >> 16);
          last = k++;
          continue;
        }
        if (kind == 2) break;
        k += 16;
        continue;
      }
      int len, rs;
      const uint16_t f = ac.fast[look >> 5];
      if (f) { len = f >> 8; rs = f & 255; }
      else {
        len = 0; rs = -1;
        for (int l = 12; l <= 16; l++) {
          const int code = look >> (16 - l);
          if (ac.maxcode[l] >= 0 && code <= ac.maxcode[l] && code >= ac.mincode[l]) { len = l; rs = ac.vals[ac.valptr[l] + code - ac.mincode[l]]; break; }
        }
        if (rs < 0) { br.acc = acc; br.n = n; br.p = bp; return -1; }
      }
      const int r = rs >> 4;
      s = rs & 15;
      if (s == 0) {
        acc <<= len; n -= len;
        if (r != 15) break;
        k += 16;
        continue;
      }
      k += r;
      if (k > 63) { br.acc = acc; br.n = n; br.p = bp; return -1; }
      const int bits = (int)((acc << len) >> (64 - s));
      acc <<= len + s; n -= len + s;
      coef[kZigzag[k]] = (int16_t)extend(bits, s);
      last = k;
      k++;
    }
    br.acc = acc; br.n = n; br.p = bp;
    return last;
  }

  // one scan: `sc` lists component indices.  Interleaved scans walk MCUs; a single-component scan walks that component's own
  // blocks (ceil(dw / 8) x ceil(dh / 8), T.81 A.2.2)
  bool scan(const uint8_t*& p, const std::vector<int>& sc) {
    ent.start(p, d + n);
    BitReader br;
    br.open(ent);
    for (int ci : sc) comp[ci].pred = 0;
    const int mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
    const bool inter = sc.size() > 1;
    Comp& c0 = comp[sc[0]];
    const int ux = inter ? mcux : (c0.dw + 7) / 8, uy = inter ? mcuy : (c0.dh + 7) / 8;
    int16_t coef[64];
    int until_restart = restart;
    for (int my = 0; my < uy; my++)
      for (int mx = 0; mx < ux; mx++) {
        if (restart && until_restart == 0) {  // expect RSTn
          if (!ent.next_restart()) { err = "missing restart marker"; return false; }
          br.open(ent);
          for (int ci : sc) comp[ci].pred = 0;
          until_restart = restart;
        }
        for (int ci : sc) {
          Comp& c = comp[ci];
          const int nh = inter ? c.h : 1, nv = inter ? c.v : 1;
          for (int by = 0; by < nv; by++)
            for (int bx = 0; bx < nh; bx++) {
              const int last = decode_block(br, c, coef);
              if (last < 0) { err = "corrupt entropy-coded data"; return false; }
              const int X = mx * nh + bx, Y = my * nv + by;
              if (X < c.bw && Y < c.bh && c.needed) {
                uint8_t* dst = c.plane.data() + ((size_t)Y * 8) * c.pitch + (size_t)X * 8;
                if (last == 0) {  // DC only: both passes of jidctint.c reduce to one value for the whole block
                  const uint8_t v = range_limit(descale((int64_t)coef[0] * qt[c.tq][0] * 4, 5));
                  for (int r = 0; r < 8; r++) memset(dst + (size_t)r * c.pitch, v, 8);
                } else {
                  idct(coef, c.tq, dst, c.pitch);
                }
              }
            }
        }
        if (restart) until_restart--;
      }
    p = ent.after();  // position after the entropy-coded segment: the next marker
    return true;
  }

  // One scan of a progressive file (ITU T.81 G.1.2; the procedure of libjpeg's jdphuff.c): DC scans may interleave components,
  // AC scans carry one component and walk its own blocks; first passes (Ah == 0) write coefficient << Al, refinement passes add one
  // bit of precision to what is already there.
  bool scan_progressive(const uint8_t*& p, const std::vector<int>& sc, int Ss, int Se, int Ah, int Al) {
    ent.start(p, d + n);
    BitReader br;
    br.open(ent);
    for (int ci : sc) comp[ci].pred = 0;
    const int mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
    const bool inter = sc.size() > 1;
    Comp& c0 = comp[sc[0]];
    const int ux = inter ? mcux : (c0.dw + 7) / 8, uy = inter ? mcuy : (c0.dh + 7) / 8;
    int until_restart = restart;
    unsigned eobrun = 0;
    const int p1 = 1 << Al, m1 = -(1 << Al);
    for (int my = 0; my < uy; my++)
      for (int mx = 0; mx < ux; mx++) {
        if (restart && until_restart == 0) {
          if (!ent.next_restart()) { err = "missing restart marker"; return false; }
          br.open(ent);
          for (int ci : sc) comp[ci].pred = 0;
          eobrun = 0;
          until_restart = restart;
        }
        for (int ci : sc) {
          Comp& c = comp[ci];
          const int nh = inter ? c.h : 1, nv = inter ? c.v : 1;
          for (int by = 0; by < nv; by++)
            for (int bx = 0; bx < nh; bx++) {
              const int X = mx * nh + bx, Y = my * nv + by;
              int16_t scratch[64];
              int16_t* blk = (X < c.bw && Y < c.bh) ? c.coefs.data() + ((size_t)Y * c.bw + X) * 64 : scratch;
              if (blk == scratch) memset(scratch, 0, sizeof scratch);
              if (Ss == 0) {
                if (Ah == 0) {  // DC first
                  const int t = huff_decode(br, hdc[c.td]);
                  if (t < 0 || t > 15) { err = "corrupt entropy-coded data"; return false; }
                  c.pred += t ? extend(br.get(t), t) : 0;
                  blk[0] = (int16_t)(c.pred * (1 << Al));
                } else if (br.get(1)) {  // DC refinement
                  blk[0] = (int16_t)(blk[0] | p1);
                }
                continue;
              }
              const Huff& ac = hac[c.ta];
              if (Ah == 0) {  // AC first
                if (eobrun > 0) { eobrun--; continue; }
                for (int k = Ss; k <= Se; k++) {
                  const int rs = huff_decode(br, ac);
                  if (rs < 0) { err = "corrupt entropy-coded data"; return false; }
                  const int r = rs >> 4, t = rs & 15;
                  if (t) {
                    k += r;
                    if (k > 63) { err = "corrupt entropy-coded data"; return false; }
                    blk[kZigzag[k]] = (int16_t)(extend(br.get(t), t) * (1 << Al));
                  } else if (r == 15) {
                    k += 15;
                  } else {
                    eobrun = 1u << r;
                    if (r) eobrun += (unsigned)br.get(r);
                    eobrun--;
                    break;
                  }
                }
                continue;
              }
              // AC refinement
              int k = Ss;
              if (eobrun == 0) {
                for (; k <= Se; k++) {
                  const int rs = huff_decode(br, ac);
                  if (rs < 0) { err = "corrupt entropy-coded data"; return false; }
                  int r = rs >> 4, t = rs & 15;
                  if (t) {
                    t = br.get(1) ? p1 : m1;  // a newly non-zero coefficient (size is always 1)
                  } else if (r != 15) {
                    eobrun = 1u << r;
                    if (r) eobrun += (unsigned)br.get(r);
                    break;  // end of band: the rest of the block is only refined
                  }
                  // skip r still-zero coefficients, refining the non-zero ones on the way
                  do {
                    int16_t& v = blk[kZigzag[k]];
                    if (v != 0) {
                      if (br.get(1) && (v & p1) == 0) v = (int16_t)(v + (v >= 0 ? p1 : m1));
                    } else if (--r < 0) {
                      break;
                    }
                    k++;
                  } while (k <= Se);
                  if (t && k <= 63) blk[kZigzag[k]] = (int16_t)t;
                }
              }
              if (eobrun > 0) {
                for (; k <= Se; k++) {
                  int16_t& v = blk[kZigzag[k]];
                  if (v != 0 && br.get(1) && (v & p1) == 0) v = (int16_t)(v + (v >= 0 ? p1 : m1));
                }
                eobrun--;
              }
            }
        }
        if (restart) until_restart--;
      }
    p = ent.after();
    return true;
  }

  bool parse() {
    if (n < 4 || d[0] != 0xFF || d[1] != 0xD8) { err = "not a JPEG file"; return false; }
    const uint8_t* p = d + 2;
    const uint8_t* end = d + n;
    bool have_sof = false;
    int scans = 0;
    while (p + 4 <= end) {
      if (p[0] != 0xFF) { p++; continue; }
      const int m = p[1];
      if (m == 0xFF) { p++; continue; }
      p += 2;
      if (m == 0xD9) break;  // EOI
      if (m == 0x01 || (m >= 0xD0 && m <= 0xD7)) continue;
      if (p + 2 > end) break;
      const int len = be16(p);
      if (len < 2 || p + len > end) { err = "truncated segment"; return false; }
      const uint8_t* s = p + 2;
      const uint8_t* se = p + len;
      if (m == 0xDB) {  // DQT
        while (s < se) {
          const int pq = s[0] >> 4, tq = s[0] & 15;
          s++;
          if (tq > 3 || s + (pq ? 128 : 64) > se) { err = "bad DQT"; return false; }
          for (int i = 0; i < 64; i++) {
            qt[tq][kZigzag[i]] = (uint16_t)(pq ? be16(s) : s[0]);
            s += pq ? 2 : 1;
          }
          qt_set[tq] = true;
          qt_small[tq] = true;
          for (int i = 0; i < 64; i++) {
            qt16[tq][i] = (int16_t)(qt[tq][i] <= 255 ? qt[tq][i] : 0);
            if (qt[tq][i] > 255) qt_small[tq] = false;
          }
        }
      } else if (m == 0xC4) {  // DHT
        while (s < se) {
          const int tc = s[0] >> 4, th = s[0] & 15;
          s++;
          if (tc > 1 || th > 3 || s + 16 > se) { err = "bad DHT"; return false; }
          Huff& h = tc ? hac[th] : hdc[th];
          int cnt = 0;
          h.bits[0] = 0;
          for (int i = 1; i <= 16; i++) { h.bits[i] = s[i - 1]; cnt += s[i - 1]; }
          s += 16;
          if (cnt > 256 || s + cnt > se) { err = "bad DHT"; return false; }
          memcpy(h.vals, s, cnt);
          s += cnt;
          if (!h.build()) { h.present = false; err = "bad DHT (over-subscribed code lengths)"; return false; }
          h.present = true;
        }
      } else if (m == 0xC0 || m == 0xC1 || m == 0xC2) {  // SOF0 / SOF1 / SOF2 (progressive)
        progressive = m == 0xC2;
        if (len < 8 || s[0] != 8) { err = "only 8-bit JPEG is supported"; return false; }
        H = be16(s + 1); W = be16(s + 3); nc = s[5];
        if (W <= 0 || H <= 0 || (nc != 1 && nc != 3) || len < 8 + 3 * nc) { err = "unsupported JPEG frame (components)"; return false; }
        for (int i = 0; i < nc; i++) {
          comp[i].id = s[6 + 3 * i];
          comp[i].h = s[7 + 3 * i] >> 4;
          comp[i].v = s[7 + 3 * i] & 15;
          comp[i].tq = s[8 + 3 * i];
          if (comp[i].h < 1 || comp[i].h > 4 || comp[i].v < 1 || comp[i].v > 4 || comp[i].tq > 3) { err = "bad SOF"; return false; }
          hmax = comp[i].h > hmax ? comp[i].h : hmax;
          vmax = comp[i].v > vmax ? comp[i].v : vmax;
        }
        const int mcux = (W + 8 * hmax - 1) / (8 * hmax), mcuy = (H + 8 * vmax - 1) / (8 * vmax);
        for (int i = 0; i < nc; i++) {
          Comp& c = comp[i];
          c.bw = mcux * c.h; c.bh = mcuy * c.v;
          c.dw = (W * c.h + hmax - 1) / hmax; c.dh = (H * c.v + vmax - 1) / vmax;
          c.pitch = c.bw * 8;
          c.needed = !(want_gray && nc == 3 && i > 0);
          if (c.needed) c.plane.assign((size_t)c.pitch * c.bh * 8, 0);
          if (progressive) c.coefs.assign((size_t)c.bw * c.bh * 64, 0);
        }
        have_sof = true;
      } else if (m >= 0xC3 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
        err = "unsupported JPEG coding process (lossless / hierarchical / arithmetic)";
        return false;
      } else if (m == 0xDD) {
        if (len >= 4) restart = be16(s);
      } else if (m == 0xEE) {  // Adobe
        if (len >= 14 && memcmp(s, "Adobe", 5) == 0) { adobe = true; adobe_transform = s[11]; }
      } else if (m == 0xDA) {  // SOS
        if (!have_sof) { err = "SOS before SOF"; return false; }
        const int ns = s[0];
        if (ns < 1 || ns > nc || len < 6 + 2 * ns) { err = "bad SOS"; return false; }
        std::vector<int> sc;
        for (int i = 0; i < ns; i++) {
          int ci = -1;
          for (int k = 0; k < nc; k++) if (comp[k].id == s[1 + 2 * i]) ci = k;
          if (ci < 0) { err = "bad SOS component"; return false; }
          comp[ci].td = s[2 + 2 * i] >> 4;
          comp[ci].ta = s[2 + 2 * i] & 15;
          const int Ss_ = s[1 + 2 * ns], Ah_ = s[3 + 2 * ns] >> 4;
          const bool need_dc = !progressive || (Ss_ == 0 && Ah_ == 0), need_ac = !progressive || Ss_ > 0;
          if (comp[ci].td > 3 || comp[ci].ta > 3 || (need_dc && !hdc[comp[ci].td].present) || (need_ac && !hac[comp[ci].ta].present) ||
              !qt_set[comp[ci].tq]) {
            err = "scan refers to a missing table";
            return false;
          }
          sc.push_back(ci);
        }
        const int Ss = s[1 + 2 * ns], Se = s[2 + 2 * ns], Ah = s[3 + 2 * ns] >> 4, Al = s[3 + 2 * ns] & 15;
        p += len;
        if (progressive) {
          if (Ss > Se || Se > 63 || (Ss == 0 && Se != 0) || (Ss > 0 && ns != 1) || Al > 13) { err = "bad progressive scan parameters"; return false; }
          if (!scan_progressive(p, sc, Ss, Se, Ah, Al)) return false;
        } else if (!scan(p, sc)) return false;
        scans++;
        continue;
      }
      p += len;
    }
    if (!have_sof || scans == 0) { err = "no image data"; return false; }
    if (progressive)  // all scans are in: reconstruct every block of the components that are wanted
      for (int i = 0; i < nc; i++) {
        Comp& c = comp[i];
        if (!c.needed) continue;
        for (int Y = 0; Y < c.bh; Y++)
          for (int X = 0; X < c.bw; X++)
            idct(c.coefs.data() + ((size_t)Y * c.bw + X) * 64, c.tq, c.plane.data() + ((size_t)Y * 8) * c.pitch + (size_t)X * 8, c.pitch);
      }
    return true;
  }

  // jdsample.c: fancy (triangle) upsampling where libjpeg uses it (downsampled_width > 2), replication otherwise.
  // Rows above the first / below the last real sample row replicate it (jdmainct.c context rows).  One output row at a time
  // (`dst` holds at least 2 * downsampled_width + 2 bytes), so no full-resolution plane is ever materialised.
  void upsample_row(const Comp& c, int y, uint8_t* __restrict__ dst) const {
    const int hx = hmax / c.h, vx = vmax / c.v;
    const bool exact = hmax % c.h == 0 && vmax % c.v == 0;
    const int n = c.dw;
    auto row = [&](int r) { r = r < 0 ? 0 : (r >= c.dh ? c.dh - 1 : r); return c.plane.data() + (size_t)r * c.pitch; };
    if (exact && hx == 1 && vx == 1) { memcpy(dst, row(y), W); return; }
    const bool fancy = n > 2;
    if (exact && fancy && hx == 2 && vx == 1) {  // h2v1_fancy_upsample
      h2v1_fancy_row(row(y), dst, n);
      return;
    }
    if (exact && fancy && hx == 2 && vx == 2) {  // h2v2_fancy_upsample
      const int r = y >> 1;
      h2v2_fancy_row(row(r), row((y & 1) ? r + 1 : r - 1), dst, n);
      return;
    }
    if (exact && hx == 1 && vx == 2) {  // h1v2_fancy_upsample (libjpeg-turbo; no width condition)
      const int r = y >> 1;
      h1v2_fancy_row(row(r), row((y & 1) ? r + 1 : r - 1), dst, W, (y & 1) ? 2 : 1);
      return;
    }
    // replication (int_upsample / h2v1_upsample / h2v2_upsample); non-integral ratios are approximated the same way
    const uint8_t* in = row(exact ? y / vx : y * c.v / vmax);
    for (int x = 0; x < W; x++) {
      int sx = exact ? x / hx : x * c.h / hmax;
      if (sx >= n) sx = n - 1;
      dst[x] = in[sx];
    }
  }

  // full-resolution row y of a component: the plane's own row when it is not subsampled, else upsampled into `tmp`
  const uint8_t* full_row(const Comp& c, int y, uint8_t* tmp) const {
    if (c.h == hmax && c.v == vmax) return c.plane.data() + (size_t)(y < c.dh ? y : c.dh - 1) * c.pitch;
    upsample_row(c, y, tmp);
    return tmp;
  }

  bool to_mat(Mat& out, bool gray) {
    const bool rgb_coded = nc == 3 && ((adobe && adobe_transform == 0) || (!adobe && comp[0].id == 'R' && comp[1].id == 'G' && comp[2].id == 'B'));
    size_t line_len = (size_t)W + 8;
    for (int i = 0; i < nc; i++) line_len = std::max(line_len, (size_t)2 * comp[i].dw + 8);
    std::vector<uint8_t> lines(3 * line_len);
    uint8_t* ln[3] = {lines.data(), lines.data() + line_len, lines.data() + 2 * line_len};
    if (gray) {
      if (rgb_coded) { err = "grey output of an RGB-coded JPEG is not supported"; return false; }
      out.create(H, W, SB_8UC1);
      for (int y = 0; y < H; y++)  // luma is never subsampled in practice; handled anyway
        memcpy(out.ptr<uint8_t>(y), full_row(comp[0], y, ln[0]), W);
      return true;
    }
    out.create(H, W, SB_8UC3);
    for (int y = 0; y < H; y++) {
      uint8_t* __restrict__ o = out.ptr<uint8_t>(y);
      const uint8_t* src[3] = {nullptr, nullptr, nullptr};
      for (int i = 0; i < nc; i++) src[i] = full_row(comp[i], y, ln[i]);
      const uint8_t* __restrict__ p0 = src[0];
      if (nc == 1) {
        grey_to_bgr_row(p0, o, W);
        continue;
      }
      const uint8_t* __restrict__ p1 = src[1];
      const uint8_t* __restrict__ p2 = src[2];
      if (rgb_coded) {
        for (int x = 0; x < W; x++) { o[3 * x] = p2[x]; o[3 * x + 1] = p1[x]; o[3 * x + 2] = p0[x]; }
        continue;
      }
      ycc_to_bgr_row(p0, p1, p2, o, W);
    }
    return true;
  }
};

// =================================================================================================== PNG
uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; }

bool png_decode(const uint8_t* d, size_t n, Mat& out, bool gray, std::string& err) {
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (n < 8 || memcmp(d, sig, 8) != 0) { err = "not a PNG file"; return false; }
  size_t p = 8;
  uint32_t W = 0, H = 0;
  int depth = 0, ctype = 0, interlace = 0;
  std::vector<uint8_t> idat, plte;
  while (p + 12 <= n) {
    const uint32_t len = be32(d + p);
    const uint8_t* type = d + p + 4;
    if (p + 12 + (size_t)len > n) { err = "truncated PNG chunk"; return false; }
    const uint8_t* body = d + p + 8;
    if (memcmp(type, "IHDR", 4) == 0 && len >= 13) {
      W = be32(body); H = be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
    } else if (memcmp(type, "PLTE", 4) == 0) {
      plte.assign(body, body + len);
    } else if (memcmp(type, "IDAT", 4) == 0) {
      idat.insert(idat.end(), body, body + len);
    } else if (memcmp(type, "IEND", 4) == 0) {
      break;
    }
    p += 12 + (size_t)len;
  }
  if (W == 0 || H == 0 || W > 65535 || H > 65535) { err = "bad PNG header"; return false; }
  const int ch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
  if (!ch || (depth != 1 && depth != 2 && depth != 4 && depth != 8 && depth != 16) || (ctype == 3 && depth == 16) ||
      ((ctype == 2 || ctype == 4 || ctype == 6) && depth < 8) || interlace > 1) { err = "unsupported PNG format"; return false; }
  const int bpp_bits = ch * depth, bpp = bpp_bits >= 8 ? bpp_bits / 8 : 1;
  auto row_bytes = [&](uint32_t w) { return ((size_t)w * bpp_bits + 7) / 8; };
  // passes: (x0, y0, dx, dy)
  static const int adam7[7][4] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
  size_t raw_size = 0;
  const int npass = interlace ? 7 : 1;
  uint32_t pw[7], ph[7];
  for (int i = 0; i < npass; i++) {
    pw[i] = interlace ? (W - adam7[i][0] + adam7[i][2] - 1) / adam7[i][2] : W;
    ph[i] = interlace ? (H - adam7[i][1] + adam7[i][3] - 1) / adam7[i][3] : H;
    if (interlace && ((int)W <= adam7[i][0] || (int)H <= adam7[i][1])) pw[i] = ph[i] = 0;
    if (pw[i] && ph[i]) raw_size += (row_bytes(pw[i]) + 1) * ph[i];
  }
  std::vector<uint8_t> raw(raw_size);
  {
    z_stream zs;
    memset(&zs, 0, sizeof zs);
    if (inflateInit(&zs) != Z_OK) { err = "zlib init failed"; return false; }
    zs.next_in = idat.data(); zs.avail_in = (uInt)idat.size();
    zs.next_out = raw.data(); zs.avail_out = (uInt)raw.size();
    const int rc = inflate(&zs, Z_FINISH);
    const size_t got = zs.total_out;
    inflateEnd(&zs);
    if ((rc != Z_STREAM_END && rc != Z_OK && rc != Z_BUF_ERROR) || got != raw.size()) { err = "PNG data does not inflate to the image size"; return false; }
  }
  const bool colour_src = ctype == 2 || ctype == 6 || ctype == 3;
  // 8-bit grey / grey+alpha / RGB / RGBA without interlace (what cameras and OpenCV write): rows go from the unfiltered scanline
  // straight to the output.  Everything else is expanded to 16-bit RGBA samples first (px) and converted below.
  const bool direct = !interlace && depth == 8 && ctype != 3;
  if (direct) out.create((int)H, (int)W, gray ? SB_8UC1 : SB_8UC3);
  // sample (x, y, channel) as 8-bit after the conversions OpenCV asks libpng for
  std::vector<uint16_t> px;
  if (!direct) px.assign((size_t)W * H * 4, 255);  // RGBA, full sample width (alpha unused)
  size_t off = 0;
  for (int ps = 0; ps < npass; ps++) {
    if (!pw[ps] || !ph[ps]) continue;
    const size_t rb = row_bytes(pw[ps]);
    const std::vector<uint8_t> zero_row(rb, 0);
    for (uint32_t y = 0; y < ph[ps]; y++) {
      const int ft = raw[off];
      uint8_t* __restrict__ cur = raw.data() + off + 1;
      const uint8_t* __restrict__ prev = y ? cur - (rb + 1) : zero_row.data();  // the row above, already unfiltered in place
      const size_t B = (size_t)bpp;
      switch (ft) {
        case 0: break;
        case 1:
          for (size_t i = B; i < rb; i++) cur[i] = (uint8_t)(cur[i] + cur[i - B]);
          break;
        case 2:
          for (size_t i = 0; i < rb; i++) cur[i] = (uint8_t)(cur[i] + prev[i]);
          break;
        case 3:
          for (size_t i = 0; i < B && i < rb; i++) cur[i] = (uint8_t)(cur[i] + (prev[i] >> 1));
          for (size_t i = B; i < rb; i++) cur[i] = (uint8_t)(cur[i] + ((cur[i - B] + prev[i]) >> 1));
          break;
        case 4:
          for (size_t i = 0; i < B && i < rb; i++) cur[i] = (uint8_t)(cur[i] + prev[i]);  // a = c = 0: the predictor is b
          for (size_t i = B; i < rb; i++) {
            const int a = cur[i - B], b = prev[i], c = prev[i - B];
            const int pp = a + b - c, pa = abs(pp - a), pb = abs(pp - b), pc = abs(pp - c);
            cur[i] = (uint8_t)(cur[i] + ((pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c)));
          }
          break;
        default: err = "bad PNG filter"; return false;
      }
      if (direct) {
        uint8_t* __restrict__ o = out.ptr<uint8_t>((int)y);
        if (gray) {
          if (ch == 1) memcpy(o, cur, W);
          else if (ch == 2) { for (uint32_t x = 0; x < W; x++) o[x] = cur[2 * (size_t)x]; }
          else
            for (uint32_t x = 0; x < W; x++) {  // png_set_rgb_to_gray, see below
              const long r = cur[(size_t)x * ch], g = cur[(size_t)x * ch + 1], b = cur[(size_t)x * ch + 2];
              o[x] = (uint8_t)((r == g && g == b) ? r : (9797 * r + 19234 * g + 3737 * b) >> 15);
            }
        } else {
          if (ch <= 2) { for (uint32_t x = 0; x < W; x++) o[3 * (size_t)x] = o[3 * (size_t)x + 1] = o[3 * (size_t)x + 2] = cur[(size_t)x * ch]; }
          else
            for (uint32_t x = 0; x < W; x++) {
              o[3 * (size_t)x] = cur[(size_t)x * ch + 2]; o[3 * (size_t)x + 1] = cur[(size_t)x * ch + 1]; o[3 * (size_t)x + 2] = cur[(size_t)x * ch];
            }
        }
        off += rb + 1;
        continue;
      }
      const uint32_t Y = interlace ? adam7[ps][1] + y * adam7[ps][3] : y;
      for (uint32_t x = 0; x < pw[ps]; x++) {
        const uint32_t X = interlace ? adam7[ps][0] + x * adam7[ps][2] : x;
        uint16_t* o = px.data() + ((size_t)Y * W + X) * 4;
        uint16_t s[4] = {0, 0, 0, 255};
        for (int k = 0; k < ch; k++) {
          if (depth == 8) s[k] = cur[(size_t)x * ch + k];
          else if (depth == 16) s[k] = (uint16_t)((cur[((size_t)x * ch + k) * 2] << 8) | cur[((size_t)x * ch + k) * 2 + 1]);
          else {
            const size_t bit = (size_t)x * depth;
            const int v = (cur[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
            s[k] = ctype == 3 ? (uint8_t)v : (uint8_t)(v * 255 / ((1 << depth) - 1));
          }
        }
        if (ctype == 3) {
          const size_t idx = s[0];
          if (idx * 3 + 2 < plte.size()) { o[0] = plte[idx * 3]; o[1] = plte[idx * 3 + 1]; o[2] = plte[idx * 3 + 2]; }
          else o[0] = o[1] = o[2] = 0;
        } else if (ch <= 2) { o[0] = o[1] = o[2] = s[0]; }
        else { o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; }
      }
      off += rb + 1;
    }
  }
  if (direct) return true;
  if (gray) {
    out.create((int)H, (int)W, SB_8UC1);
    for (size_t i = 0; i < (size_t)W * H; i++) {
      const long r = px[4 * i], g = px[4 * i + 1], b = px[4 * i + 2];
      // png_set_rgb_to_gray(png, 1, 0.299, 0.587): 15-bit coefficients 9797 / 19234 / 3737 (libpng truncates them); 8-bit samples
      // are not rounded, 16-bit samples are (and are converted before png_set_strip_16 drops the low byte); equal channels pass
      long v = r;
      if (colour_src && !(r == g && g == b)) v = depth == 16 ? (9797 * r + 19234 * g + 3737 * b + 16384) >> 15 : (9797 * r + 19234 * g + 3737 * b) >> 15;
      out.data[i] = (uint8_t)(depth == 16 ? v >> 8 : v);
    }
  } else {
    out.create((int)H, (int)W, SB_8UC3);
    const int sh = depth == 16 ? 8 : 0;  // png_set_strip_16: the high byte
    for (size_t i = 0; i < (size_t)W * H; i++) {
      out.data[3 * i] = (uint8_t)(px[4 * i + 2] >> sh); out.data[3 * i + 1] = (uint8_t)(px[4 * i + 1] >> sh); out.data[3 * i + 2] = (uint8_t)(px[4 * i] >> sh);
    }
  }
  return true;
}

// =================================================================================================== BMP
bool bmp_decode(const uint8_t* d, size_t n, Mat& out, bool gray, std::string& err) {
  auto le32 = [&](size_t o) { return (uint32_t)d[o] | (d[o + 1] << 8) | (d[o + 2] << 16) | ((uint32_t)d[o + 3] << 24); };
  auto le16 = [&](size_t o) { return (int)(d[o] | (d[o + 1] << 8)); };
  if (n < 54 || d[0] != 'B' || d[1] != 'M') { err = "not a BMP file"; return false; }
  const uint32_t offs = le32(10), hdr = le32(14);
  if (hdr < 40) { err = "unsupported BMP header"; return false; }
  const int W = (int)le32(18);
  int H = (int)le32(22);
  const int bpp = le16(28);
  const uint32_t compr = le32(30);
  const bool top_down = H < 0;
  if (top_down) H = -H;
  if (W <= 0 || H <= 0 || (bpp != 24 && bpp != 32) || (compr != 0 && !(compr == 3 && bpp == 32))) { err = "unsupported BMP format"; return false; }
  const size_t stride = ((size_t)W * (bpp / 8) + 3) & ~(size_t)3;
  if (offs + stride * H > n) { err = "truncated BMP"; return false; }
  out.create(H, W, gray ? SB_8UC1 : SB_8UC3);
  for (int y = 0; y < H; y++) {
    const uint8_t* src = d + offs + stride * (top_down ? y : H - 1 - y);
    uint8_t* o = out.ptr<uint8_t>(y);
    for (int x = 0; x < W; x++) {
      const int b = src[x * (bpp / 8)], g = src[x * (bpp / 8) + 1], r = src[x * (bpp / 8) + 2];
      if (gray) o[x] = (uint8_t)((r * 4899 + g * 9617 + b * 1868 + (1 << 13)) >> 14);
      else { o[3 * x] = (uint8_t)b; o[3 * x + 1] = (uint8_t)g; o[3 * x + 2] = (uint8_t)r; }
    }
  }
  return true;
}

thread_local std::string g_imread_error;

}  // namespace

const std::string& imread_error() { return g_imread_error; }

bool imdecode(const uint8_t* d, size_t n, Mat& out, bool grayscale) {
  out.release();
  g_imread_error.clear();
  if (n >= 2 && d[0] == 0xFF && d[1] == 0xD8) {
    Jpeg j;
    j.d = d; j.n = n;
    j.want_gray = grayscale;
    memset(j.qt, 0, sizeof j.qt);
    if (!j.parse() || !j.to_mat(out, grayscale)) { g_imread_error = j.err; out.release(); return false; }
    return true;
  }
  if (n >= 8 && d[0] == 0x89 && d[1] == 'P') {
    if (!png_decode(d, n, out, grayscale, g_imread_error)) { out.release(); return false; }
    return true;
  }
  if (n >= 2 && d[0] == 'B' && d[1] == 'M') {
    if (!bmp_decode(d, n, out, grayscale, g_imread_error)) { out.release(); return false; }
    return true;
  }
  g_imread_error = "unknown image format";
  return false;
}

bool imread(const std::string& path, Mat& out, bool grayscale) {
  out.release();
  FILE* fp = fopen(path.c_str(), "rb");
  if (!fp) { g_imread_error = "cannot open " + path; return false; }
  uint8_t magic[2] = {0, 0};
  const size_t got = fread(magic, 1, 2, fp);
  if (got == 2 && magic[0] == 'P' && (magic[1] == '5' || magic[1] == '6')) {
    fclose(fp);
    return imread_pnm(path, out, grayscale);
  }
  fseek(fp, 0, SEEK_END);
  const long sz = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  if (sz <= 0) { fclose(fp); g_imread_error = "empty file"; return false; }
  std::vector<uint8_t> buf((size_t)sz);
  const bool ok = fread(buf.data(), 1, buf.size(), fp) == buf.size();
  fclose(fp);
  if (!ok) { g_imread_error = "read error"; return false; }
  return imdecode(buf.data(), buf.size(), out, grayscale);
}

// ---------------------------------------------------------------------------------------------- ImagePrefetcher
struct ImagePrefetcher::Impl {
  struct Entry {
    std::string path;
    bool gray = false;
    int users = 0;
    bool done = false, ok = false;
    std::string err;
    Mat img;
  };
  std::vector<Entry> entries;                       // distinct (path, gray), in first-request order
  std::map<std::pair<std::string, bool>, size_t> index;
  std::mutex mu;
  std::condition_variable cv;
  std::atomic<size_t> next{0};
  std::vector<std::thread> threads;
  // Bounded look-ahead: at most `window` decoded images wait for their consumers (the reference holds one pair at a
  // time; a 20-view rig decoded all at once is > 1 GB of host memory).  An entry a consumer is already waiting for is
  // always decoded (i <= max_wanted), so the bound cannot dead-lock whatever order the consumers ask in.
  size_t live = 0, window = 16, max_wanted = 0;
  bool stop = false;  // set by the destructor: consumers are gone, waiting decoders leave
  void work() {
    for (;;) {
      const size_t i = next.fetch_add(1);
      if (i >= entries.size()) return;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&]() { return stop || live < window || i <= max_wanted; });
        if (stop) return;
        live++;
      }
      Mat m;
      bool ok = false;
      std::string e;
      try {
        ok = imread(entries[i].path, m, entries[i].gray);
        if (!ok) e = imread_error();
      } catch (const std::exception& ex) {  // e.g. std::bad_alloc: reported through get(), never terminates the process
        ok = false;
        m.release();
        e = std::string("decoding ") + entries[i].path + ": " + ex.what();
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        entries[i].img = m;
        entries[i].ok = ok;
        entries[i].err = e;
        entries[i].done = true;
        if (!ok || entries[i].users <= 0) live--;
      }
      cv.notify_all();
    }
  }
};

ImagePrefetcher::ImagePrefetcher() : impl_(new Impl) {}
ImagePrefetcher::~ImagePrefetcher() {
  {
    std::lock_guard<std::mutex> lk(impl_->mu);
    impl_->stop = true;
  }
  impl_->cv.notify_all();
  for (std::thread& t : impl_->threads) t.join();
  delete impl_;
}
bool ImagePrefetcher::active() const { return !impl_->entries.empty(); }

void ImagePrefetcher::start(const std::vector<std::pair<std::string, bool>>& requests, int n_threads) {
  if (active()) return;
  for (const auto& r : requests) {
    auto it = impl_->index.find(r);
    if (it == impl_->index.end()) {
      impl_->index[r] = impl_->entries.size();
      Impl::Entry e;
      e.path = r.first;
      e.gray = r.second;
      e.users = 1;
      impl_->entries.push_back(e);
    } else {
      impl_->entries[it->second].users++;
    }
  }
  if (const char* e = getenv("SB200_DECODE_WINDOW")) {
    if (atoi(e) > 0) impl_->window = (size_t)atoi(e);
  }
  if (n_threads < 1) n_threads = 1;
  if (impl_->window < (size_t)n_threads) impl_->window = (size_t)n_threads;
  if ((size_t)n_threads > impl_->entries.size()) n_threads = (int)impl_->entries.size();
  for (int t = 0; t < n_threads; t++) impl_->threads.emplace_back([this]() { impl_->work(); });
}

bool ImagePrefetcher::get(const std::string& path, bool grayscale, Mat& out, std::string* err) {
  out.release();
  auto it = impl_->index.find(std::make_pair(path, grayscale));
  if (it == impl_->index.end()) {
    if (err) *err = "not requested: " + path;
    return false;
  }
  std::unique_lock<std::mutex> lk(impl_->mu);
  Impl::Entry& e = impl_->entries[it->second];
  if (it->second > impl_->max_wanted) {
    impl_->max_wanted = it->second;
    impl_->cv.notify_all();
  }
  impl_->cv.wait(lk, [&]() { return e.done; });
  if (err) *err = e.err;
  const bool ok = e.ok;
  if (ok) out = e.img;
  if (--e.users == 0 && ok) {  // last consumer: the cache lets go of the pixels and a decoder may run further ahead
    e.img.release();
    impl_->live--;
    impl_->cv.notify_all();
  }
  return ok;
}

}  // namespace sbcv
