// encode_simd.h -- host-side residue encoder of tsq_set_sequences: input bytes -> symbols, gap and
// whitespace bytes dropped, self score S(x, x) summed on the way.  Letter map of
// tweakseq/Core/Annotations/Consensus.cpp:61-69 (the 256-entry table the caller passes in states it;
// the vector path restates it as two 16-entry shuffles and is checked against the table at start-up).
//
// Plain C++ (compiled by the host compiler, not nvcc): the AVX2 body sits behind a target attribute and
// is chosen at run time; without AVX2 the table loop runs.
#pragma once
#include <cstddef>
#include <cstdint>

namespace tsq {

struct EncodeTables {
  uint8_t lut[256];          // byte -> symbol, 0xff = dropped ('-', '.', whitespace)
  int32_t diag_by_byte[256]; // byte -> S(x, x) of its symbol (0 for dropped bytes)
  // vector form, filled by encode_tables_finish(): symbols of the letters a..z, the symbol of everything
  // else, and S(x, x) per symbol (int8: matrix entries are int8)
  alignas(16) uint8_t letter_sym[32];
  alignas(16) int8_t sym_diag[32];
  uint8_t other_sym;
  bool vector_ok;            // the vector form reproduces lut / diag_by_byte for all 256 bytes
};

// Derives the vector form from lut / diag_by_byte and verifies it byte by byte (vector_ok).
void encode_tables_finish(EncodeTables* t);

// One sequence of a batch: s[0, len) in; out (room for len bytes) and out_len, self = sum of S(x, x) back.
struct EncodeJob {
  const char* s;
  size_t len;
  uint8_t* out;
  size_t out_len;
  int64_t self;
};
void encode_many(const EncodeTables& t, EncodeJob* jobs, size_t n);

// out must have room for len bytes.  Returns the number of symbols written; *self = sum of S(x, x).
size_t encode_residues(const EncodeTables& t, const char* s, size_t len, uint8_t* out, int64_t* self);

// the table loop alone (what encode_residues falls back to): for tests and for the start-up check
size_t encode_residues_scalar(const EncodeTables& t, const char* s, size_t len, uint8_t* out, int64_t* self);

bool encode_uses_avx2();

}  // namespace tsq
