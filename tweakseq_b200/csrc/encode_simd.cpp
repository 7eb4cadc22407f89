// encode_simd.cpp -- see encode_simd.h.
#include "encode_simd.h"

#include <cstring>
#include <immintrin.h>

namespace tsq {

namespace {

inline bool in_drop_set(unsigned b) { return b == '-' || b == '.' || b == ' ' || (b >= 9 && b <= 13); }

}  // namespace

void encode_tables_finish(EncodeTables* t) {
  memset(t->letter_sym, 0, sizeof t->letter_sym);
  memset(t->sym_diag, 0, sizeof t->sym_diag);
  for (int k = 0; k < 26; k++) t->letter_sym[k] = t->lut['a' + k];
  t->other_sym = t->lut[0];
  bool ok = !in_drop_set(0) && t->other_sym < 32;
  for (int b = 0; b < 256 && ok; b++) {
    const uint8_t v = t->lut[b];
    if (v == 0xff) {
      ok = in_drop_set((unsigned)b) && t->diag_by_byte[b] == 0;
      continue;
    }
    const int u = b | 0x20;
    const bool letter = u >= 'a' && u <= 'z';
    ok = !in_drop_set((unsigned)b) && v < 32 && v == (letter ? t->letter_sym[u - 'a'] : t->other_sym) &&
         t->diag_by_byte[b] >= -128 && t->diag_by_byte[b] <= 127;
    if (ok) {
      // one S(x, x) per symbol: two bytes of the same symbol must agree
      const int8_t d = (int8_t)t->diag_by_byte[b];
      if (t->sym_diag[v] != 0 && t->sym_diag[v] != d) ok = false;
      t->sym_diag[v] = d;
    }
  }
  // a symbol whose S(x, x) is 0 and one that was never seen look alike above; check the finished table once more
  for (int b = 0; b < 256 && ok; b++)
    if (t->lut[b] != 0xff) ok = t->sym_diag[t->lut[b]] == t->diag_by_byte[b];
  t->vector_ok = ok;
}

size_t encode_residues_scalar(const EncodeTables& t, const char* s, size_t len, uint8_t* out, int64_t* self) {
  size_t k = 0;
  int64_t sum = 0;
  for (size_t i = 0; i < len; i++) {
    const unsigned char ch = (unsigned char)s[i];
    const uint8_t v = t.lut[ch];
    out[k] = v;
    k += (v != 0xff);
    sum += t.diag_by_byte[ch];
  }
  *self = sum;
  return k;
}

namespace {

#define TSQ_AVX2 __attribute__((target("avx2")))

struct Avx2Tables {
  __m256i sym_lo, sym_hi, dg_lo, dg_hi, other, c20, ca, c25, c15, cdash, cdot, c9, c4, ones8, ones16;
};

// one block: symbols of 32 bytes in *r_out, mask of the dropped ones returned, S(x, x) of the kept ones added to *acc
TSQ_AVX2 inline unsigned encode_block(const Avx2Tables& c, const __m256i v, __m256i* r_out, __m256i* acc) {
  const __m256i idx = _mm256_sub_epi8(_mm256_or_si256(v, c.c20), c.ca);                     // a..z -> 0..25
  const __m256i letter = _mm256_cmpeq_epi8(_mm256_min_epu8(idx, c.c25), idx);
  const __m256i r0 = _mm256_shuffle_epi8(c.sym_lo, idx), r1 = _mm256_shuffle_epi8(c.sym_hi, idx);
  __m256i r = _mm256_blendv_epi8(r0, r1, _mm256_cmpgt_epi8(idx, c.c15));
  r = _mm256_blendv_epi8(c.other, r, letter);
  // dropped bytes: '-', '.', ' ', 9..13
  const __m256i ws = _mm256_sub_epi8(v, c.c9);
  __m256i drop = _mm256_or_si256(_mm256_cmpeq_epi8(v, c.cdash), _mm256_cmpeq_epi8(v, c.cdot));
  drop = _mm256_or_si256(drop, _mm256_cmpeq_epi8(v, c.c20));
  drop = _mm256_or_si256(drop, _mm256_cmpeq_epi8(_mm256_min_epu8(ws, c.c4), ws));
  // S(x, x) of every kept byte
  const __m256i d0 = _mm256_shuffle_epi8(c.dg_lo, r), d1 = _mm256_shuffle_epi8(c.dg_hi, r);
  __m256i d = _mm256_blendv_epi8(d0, d1, _mm256_cmpgt_epi8(r, c.c15));
  d = _mm256_andnot_si256(drop, d);
  *acc = _mm256_add_epi32(*acc, _mm256_madd_epi16(_mm256_maddubs_epi16(c.ones8, d), c.ones16));
  *r_out = r;
  return (unsigned)_mm256_movemask_epi8(drop);
}

TSQ_AVX2 inline void avx2_tables(const EncodeTables& t, Avx2Tables* c) {
  c->sym_lo = _mm256_broadcastsi128_si256(_mm_load_si128(reinterpret_cast<const __m128i*>(t.letter_sym)));
  c->sym_hi = _mm256_broadcastsi128_si256(_mm_load_si128(reinterpret_cast<const __m128i*>(t.letter_sym + 16)));
  c->dg_lo = _mm256_broadcastsi128_si256(_mm_load_si128(reinterpret_cast<const __m128i*>(t.sym_diag)));
  c->dg_hi = _mm256_broadcastsi128_si256(_mm_load_si128(reinterpret_cast<const __m128i*>(t.sym_diag + 16)));
  c->other = _mm256_set1_epi8((char)t.other_sym);
  c->c20 = _mm256_set1_epi8(0x20), c->ca = _mm256_set1_epi8('a'), c->c25 = _mm256_set1_epi8(25), c->c15 = _mm256_set1_epi8(15);
  c->cdash = _mm256_set1_epi8('-'), c->cdot = _mm256_set1_epi8('.'), c->c9 = _mm256_set1_epi8(9), c->c4 = _mm256_set1_epi8(4);
  c->ones8 = _mm256_set1_epi8(1), c->ones16 = _mm256_set1_epi16(1);
}

// 32 x 0xff, then 32 x 0: loaded at offset rem, the first 32 - rem bytes of the vector are set
alignas(32) const uint8_t kHeadMask[64] = {255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255,
                                           255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255};

// len >= 32
TSQ_AVX2 inline size_t encode_avx2(const Avx2Tables& c, const char* s, size_t len, uint8_t* out, int64_t* self) {
  __m256i acc = _mm256_setzero_si256();
  int64_t sum = 0;
  size_t k = 0, i = 0, since_flush = 0;
  for (; i + 32 <= len; i += 32) {
    __m256i r;
    const unsigned dm = encode_block(c, _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i)), &r, &acc);
    if (dm == 0) {
      _mm256_storeu_si256(reinterpret_cast<__m256i*>(out + k), r);
      k += 32;
    } else {
      alignas(32) uint8_t tmp[32];
      _mm256_store_si256(reinterpret_cast<__m256i*>(tmp), r);
      for (unsigned j = 0; j < 32; j++) {
        out[k] = tmp[j];
        k += !((dm >> j) & 1u);
      }
    }
    if (++since_flush == (1u << 20)) {   // 4 x 127 per lane and block: far from 2^31
      alignas(32) int32_t lanes[8];
      _mm256_store_si256(reinterpret_cast<__m256i*>(lanes), acc);
      for (int q = 0; q < 8; q++) sum += lanes[q];
      acc = _mm256_setzero_si256();
      since_flush = 0;
    }
  }
  if (i < len) {
    // the last, partial block: the LAST 32 bytes of the input once more, with the 32 - rem bytes that were already
    // handled masked out of the score.  If nothing has been dropped so far (k == i) and the new bytes hold
    // no dropped one either, the block is stored over its own earlier output; else the new symbols go out one by one.
    const unsigned rem = (unsigned)(len - i);
    const __m256i head = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(kHeadMask + rem));
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + len - 32));
    // (masking the input to '-' keeps the masked bytes out of acc: '-' is dropped and scores nothing)
    const __m256i vm = _mm256_blendv_epi8(v, c.cdash, head);
    __m256i r;
    const unsigned dm = encode_block(c, vm, &r, &acc);
    if (k == i && (dm >> (32 - rem)) == 0) {
      const __m256i prev = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(out + len - 32));   // head: symbols already out
      _mm256_storeu_si256(reinterpret_cast<__m256i*>(out + len - 32), _mm256_blendv_epi8(r, prev, head));
      k += rem;
    } else {
      alignas(32) uint8_t tmp[32];
      _mm256_store_si256(reinterpret_cast<__m256i*>(tmp), r);
      for (unsigned j = 32 - rem; j < 32; j++) {
        out[k] = tmp[j];
        k += !((dm >> j) & 1u);
      }
    }
  }
  alignas(32) int32_t lanes[8];
  _mm256_store_si256(reinterpret_cast<__m256i*>(lanes), acc);
  for (int q = 0; q < 8; q++) sum += lanes[q];
  *self = sum;
  return k;
}

bool have_avx2() {
  __builtin_cpu_init();
  return __builtin_cpu_supports("avx2");
}
const bool kHaveAvx2 = have_avx2();

}  // namespace

bool encode_uses_avx2() { return kHaveAvx2; }

namespace {

// shorter than a block: through a '-'-padded copy
TSQ_AVX2 size_t encode_avx2_short(const Avx2Tables& c, const char* s, size_t len, uint8_t* out, int64_t* self) {
  alignas(32) uint8_t in[32], tmp[32];
  memset(in, '-', 32);
  memcpy(in, s, len);
  __m256i r, acc = _mm256_setzero_si256();
  const unsigned dm = encode_block(c, _mm256_load_si256(reinterpret_cast<const __m256i*>(in)), &r, &acc);
  _mm256_store_si256(reinterpret_cast<__m256i*>(tmp), r);
  size_t k = 0;
  for (unsigned j = 0; j < (unsigned)len; j++) {
    out[k] = tmp[j];
    k += !((dm >> j) & 1u);
  }
  alignas(32) int32_t lanes[8];
  _mm256_store_si256(reinterpret_cast<__m256i*>(lanes), acc);
  int64_t sum = 0;
  for (int q = 0; q < 8; q++) sum += lanes[q];
  *self = sum;
  return k;
}

TSQ_AVX2 void encode_many_avx2(const EncodeTables& t, EncodeJob* jobs, size_t n) {
  Avx2Tables c;
  avx2_tables(t, &c);
  for (size_t i = 0; i < n; i++) {
    EncodeJob& j = jobs[i];
    j.out_len = j.len >= 32 ? encode_avx2(c, j.s, j.len, j.out, &j.self)
                            : j.len > 0 ? encode_avx2_short(c, j.s, j.len, j.out, &j.self) : (j.self = 0, (size_t)0);
  }
}

}  // namespace

void encode_many(const EncodeTables& t, EncodeJob* jobs, size_t n) {
  if (kHaveAvx2 && t.vector_ok) {
    encode_many_avx2(t, jobs, n);
    return;
  }
  for (size_t i = 0; i < n; i++) jobs[i].out_len = encode_residues_scalar(t, jobs[i].s, jobs[i].len, jobs[i].out, &jobs[i].self);
}

size_t encode_residues(const EncodeTables& t, const char* s, size_t len, uint8_t* out, int64_t* self) {
  EncodeJob j = {s, len, out, 0, 0};
  encode_many(t, &j, 1);
  *self = j.self;
  return j.out_len;
}

}  // namespace tsq
