// gotoh16.cuh -- regime 1: inter-task packed 16-bit Gotoh kernel for sm_100a.
//
// One warp = one task = (a PAIR of query sequences A1,A2) x (32 subject sequences, one per
// lane).  The two queries ride in the two 16-bit halves of every 32-bit DP word, so each
// DPX instruction (VIMNMX3.U16x2 / VIADDMNMX.U16x2) advances two alignments.
//
// The DP is strip-mined: a lane keeps K columns of H and F in registers and walks down all
// rows of its subject, two rows at a time (row i+1 one column behind row i: two independent
// dependency chains per lane); the right-hand boundary column (H, E) of the strip goes to a
// per-lane scratch column in global memory and is read back, two rows ahead of use, by the next
// strip.  Lanes never exchange data, so the row loop has no barrier and no shuffle.
//
// Scores come from a per-warp, per-strip "query-pair profile" in shared memory:
//   prof[b][c] = S'(A1[j0+c], b) | S'(A2[j0+c], b) << 16          (b = subject letter)
// with row stride == 1 (mod 32): all lanes read the same column c of different rows b, so
// the bank is (b + c) mod 32 -- distinct letters hit distinct banks, equal letters broadcast.
// One conflict-free LDS.32 per packed cell, no ALU work for the lookup.
//
// Arithmetic (exactness argument in DESIGN.md section 4):
//   * values are stored biased and skewed:  v~(i,j) = v(i,j) + delta*(i+j) + BIAS, as
//     unsigned 16-bit.  delta = ceil(-min(S)/2) makes every substitution score S' = S+2*delta
//     non-negative, so "H_diag + S" and "H - (go+ge)" are plain 32-bit adds of two packed
//     halves with no carry between halves -> ptxas is free to place them on either integer
//     pipe (IMAD.IADD / VIADD), leaving 3 DPX instructions per packed cell:
//         t  = hd + S'                      (plain add)
//         h  = vimax3(t, E, F)              (ALU, DPX)
//         hg = h - goe'                     (plain add)
//         E  = viaddmax(E, -ge', hg)        (ALU, DPX)
//         F  = viaddmax(F, -ge', hg)        (ALU, DPX)
//     with ge' = ge - delta, goe' = go + ge - delta.
//   * -inf is never needed: E(i,0) and F(0,j) only ever feed max(x - ge, H - go - ge), which
//     equals H - go - ge for any x <= H - go, so the boundary E/F are seeded with H - goe'.
//
// Spec: SURVEY.md section 8c (frozen oracle spec).  Matrix/alphabet:
// tweakseq/Core/Annotations/Consensus.cpp:34-69.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tsq {

constexpr int kG16MaxRanges = 16;

struct G16Params {
  const uint32_t* dbw;         // subject residues, 32-way interleaved, 2 rows per word as 16-bit
                               // profile-row byte offsets (letter * STRIDE * 4), see the row loop
  const uint32_t* goff;        // word offset of each group of 32 sorted sequences
  const uint8_t* lin;          // residues, linear, sorted order
  const uint32_t* loff;        // start of each sorted sequence in lin
  const uint32_t* lens;        // sorted lengths
  const unsigned long long* task_prefix;  // [nq+1] cumulative chunk counts, per query pair
  unsigned long long* counter; // dynamic task cursor
  const int* cancel;           // device flag (set by a side-stream copy): != 0 stops task fetching
  uint2* bnd;                  // strip boundary scratch: [warp slot][row][lane] (H, E)
  const uint32_t* sbias;       // (nsym+1) x nsym biased scores S' (row nsym = padding = 0)
  int32_t* out;                // scores, packed upper triangle in sorted order
  unsigned long long ntasks;
  uint32_t bnd_rows;           // rows per warp slot in bnd
  uint32_t n_total;            // N of the triangle
  uint32_t lo, hi;             // the packed-16 eligible sorted range [lo, hi)
  uint32_t q_begin, q_end;     // query pairs of this launch (rows lo+2q, lo+2q+1)
  uint32_t nsym;
  uint32_t bias;               // BIAS
  int32_t delta;               // skew per anti-diagonal
  int32_t go;                  // gap open
  int32_t gep;                 // ge' = ge - delta
  uint32_t negge2;             // (-ge' mod 2^16) in both halves
  uint32_t goe2;               // goe' * 0x10001 (mod 2^32)
  // ---- results that leave while the launch is still running (tsq_stream_results, sorted order = final order) ----
  int16_t* out16;              // != nullptr: every score also as int16 (TSQ_FLAG_SCORES_I16; the host has checked the range)
  const int32_t* self;         // self scores S(x, x), sorted order
  double* out_dist;            // != nullptr: the distance 1 - S / min(S_ii, S_jj) goes out next to every score
                               //   (what finalize_kernel would compute from it afterwards, operation for operation)
  unsigned int* done;          // != nullptr: done[k] counts the finished tasks of row range k; a copy stream waits
                               //   for done[k] == its task count (stream memory operation) and copies the range out
  uint32_t nranges;
  unsigned long long range_end[kG16MaxRanges];   // tasks [range_end[k-1], range_end[k]) belong to range k
};

__device__ __forceinline__ unsigned long long tri_index(unsigned long long i, unsigned long long j,
                                                        unsigned long long n) {
  return i * n - i * (i + 1) / 2 + (j - i - 1);
}

template <int K>
struct G16Cfg {
  static constexpr int STRIDE = ((K + 31) & ~31) + 1;  // == 1 (mod 32)
};

// End of a task: the two scores of every lane, un-biased; when the results are streamed (tsq_stream_results) also
// their distances and the task's tick in its row range's counter.  Not inlined: once per task, and its temporaries
// stay out of the row loop's register allocation.
static __device__ __noinline__ void g16_task_results(const G16Params& p, unsigned long long task, int lane, bool valid, uint32_t A1,
                                              uint32_t j, uint32_t L1, uint32_t L2, uint32_t Ls, uint32_t res1, uint32_t res2) {
  const uint32_t A2 = A1 + 1;
  if (valid) {
    const int32_t base = -(int32_t)p.bias - p.delta * (int32_t)Ls;
    const int32_t s1 = (int32_t)res1 + base - p.delta * (int32_t)L1;
    const int32_t s2 = (int32_t)res2 + base - p.delta * (int32_t)L2;
    const unsigned long long i1 = tri_index(A1, j, p.n_total), i2 = tri_index(A2, j, p.n_total);
    p.out[i1] = s1;
    if (j > A2) p.out[i2] = s2;
    if (p.out16) {
      p.out16[i1] = (int16_t)s1;
      if (j > A2) p.out16[i2] = (int16_t)s2;
    }
    if (p.out_dist) {   // finalize_kernel's distance (tsq_device.cu), here because the rows leave before the launch ends
      const int32_t sj = p.self[j], sa = p.self[A1], sb = p.self[A2];
      const int32_t m1 = sa < sj ? sa : sj, m2 = sb < sj ? sb : sj;
      p.out_dist[i1] = m1 > 0 ? __dsub_rn(1.0, __ddiv_rn((double)s1, (double)m1)) : 1.0;
      if (j > A2) p.out_dist[i2] = m2 > 0 ? __dsub_rn(1.0, __ddiv_rn((double)s2, (double)m2)) : 1.0;
    }
  }
  if (p.done) {
    // every lane's results are visible device-wide before lane 0 counts the task as finished
    __threadfence();
    __syncwarp();
    if (lane == 0) {
      uint32_t k = 0;
      while (k + 1 < p.nranges && task >= p.range_end[k]) ++k;
      atomicAdd(p.done + k, 1u);
    }
  }
}

// NGE != 0: -ge' (both halves) is the compile-time constant NGE, so the two VIADDMNMX of a cell
// take it as an immediate (one register read less each); NGE == 0: taken from the parameters.
template <int K, int TPB, int MINB, uint32_t NGE>
__global__ void __launch_bounds__(TPB, MINB) gotoh16_kernel(const __grid_constant__ G16Params p) {
  constexpr int STRIDE = G16Cfg<K>::STRIDE;
  extern __shared__ uint32_t smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const uint32_t nsym = p.nsym;
  const uint32_t sbsz = (nsym + 1) * nsym;
  uint32_t* sb = smem;
  uint32_t* prof = smem + ((sbsz + 31) & ~31u) + wib * (nsym * STRIDE);
  for (uint32_t i = threadIdx.x; i < sbsz; i += TPB) sb[i] = p.sbias[i];
  __syncthreads();

  const uint32_t gw = blockIdx.x * (TPB / 32) + wib;
  uint2* const bnd = p.bnd + (size_t)gw * p.bnd_rows * 32 + lane;
  const uint32_t nge = NGE ? NGE : p.negge2;
  const uint32_t goe2 = p.goe2;
  const int32_t gep = p.gep;

  for (;;) {
    unsigned long long task = 0;
    if (lane == 0) {
      task = atomicAdd(p.counter, 1ULL);
      if (*reinterpret_cast<const volatile int*>(p.cancel) != 0) task = ~0ULL;   // "Stop" pressed
    }
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= p.ntasks) break;

    // task -> (query pair q, subject chunk); big tasks (long queries, long subjects) first
    uint32_t r;
    {
      uint32_t a = 0, b = p.q_end - p.q_begin;
      while (b - a > 1) {
        const uint32_t m = (a + b) >> 1;
        if (p.task_prefix[m] <= task) a = m; else b = m;
      }
      r = a;
    }
    const uint32_t q = p.q_end - 1 - r;
    const unsigned long long pr = p.task_prefix[r];
    const uint32_t nch = (uint32_t)(p.task_prefix[r + 1] - pr);
    const uint32_t chunk = nch - 1 - (uint32_t)(task - pr);
    const uint32_t A1 = p.lo + 2 * q, A2 = A1 + 1;
    const uint32_t L1 = p.lens[A1], L2 = p.lens[A2];
    const uint8_t* q1 = p.lin + p.loff[A1];
    const uint8_t* q2 = p.lin + p.loff[A2];
    const uint32_t j = A1 + 1 + chunk * 32 + lane;
    const bool valid = j < p.hi;
    const uint32_t Ls = valid ? p.lens[j] : 0u;
    const uint32_t* dbp = p.dbw + (valid ? (p.goff[j >> 5] + (j & 31)) : 0u);
    const uint32_t nstrips = (L2 + K - 1) / K;  // L1 <= L2 (sorted ascending)
    uint32_t res1 = 0, res2 = 0;

    for (uint32_t s = 0; s < nstrips; ++s) {
      const uint32_t j0 = s * K;
      // ---- profile of columns [j0, j0+K) for this query pair --------------------------
      __syncwarp();
      for (int c = lane; c < K; c += 32) {
        const uint32_t col = j0 + c;
        const uint32_t a1 = col < L1 ? q1[col] : nsym;
        const uint32_t a2 = col < L2 ? q2[col] : nsym;
        const uint32_t* r1 = sb + a1 * nsym;
        const uint32_t* r2 = sb + a2 * nsym;
        for (uint32_t b = 0; b < nsym; ++b) prof[b * STRIDE + c] = r1[b] | (r2[b] << 16);
      }
      __syncwarp();

      // ---- row 0 of the strip ---------------------------------------------------------
      uint32_t H[K], F[K];
      const int32_t h0 = (int32_t)p.bias - p.go - (int32_t)(j0 + 1) * gep;
#pragma unroll
      for (int c = 0; c < K; ++c) {
        const uint32_t v = (uint32_t)(h0 - c * gep) * 0x10001u;  // H~(0, j0+c+1), both halves
        H[c] = v;
        F[c] = v - goe2;                                         // F~(1, j0+c+1)
      }
      uint32_t hdiag = (j0 == 0) ? p.bias * 0x10001u
                                 : (uint32_t)((int32_t)p.bias - p.go - (int32_t)j0 * gep) * 0x10001u;
      const int32_t hl0 = (int32_t)p.bias - p.go;  // H~(i,0) = hl0 - i*ge'
      const bool last = (s + 1 == nstrips);

      // subject letters: the database stores, per residue, the BYTE OFFSET of its profile row
      // (letter * STRIDE * 4) as 16 bits, two rows per 32-bit word, right-aligned to an even
      // count (a pad entry leads an odd-length sequence): one word = one row pair, fetched two
      // pairs ahead (w1 next, w2 in flight).
      const char* const profb = reinterpret_cast<const char*>(prof);
      uint32_t widx = 2;
      uint32_t w1 = valid ? __ldg(dbp) : 0u;
      uint32_t w2 = valid ? __ldg(dbp + 32) : 0u;
      auto next_word = [&]() -> uint32_t {
        const uint32_t w = w1;
        w1 = w2;
        w2 = __ldg(dbp + (size_t)widx * 32);
        ++widx;
        return w;
      };
      // Strip 0 has no strip to its left: its boundary column (H~(i,0), E entering column 1) is
      // a formula.  Writing it into the scratch column first keeps the row loop free of a
      // first-strip case (a dozen predicated instructions per row pair otherwise).
      if (s == 0) {
        for (uint32_t r = 1; r <= Ls; ++r) {
          const uint32_t hl = (uint32_t)(hl0 - (int32_t)r * gep) * 0x10001u;
          bnd[(size_t)r * 32] = make_uint2(hl, hl - goe2);
        }
      }
      uint32_t i = 1;
      // ---- odd row count: row 1 alone, so that the main loop can take rows two at a time ----
      if (Ls & 1u) {
        const uint32_t* prow = reinterpret_cast<const uint32_t*>(profb + (next_word() >> 16));
        const uint2 lb = bnd[32];
        uint32_t E = lb.y;
        uint32_t t = hdiag + prow[0];
        hdiag = lb.x;
#pragma unroll
        for (int c = 0; c < K; ++c) {
          uint32_t tn = 0;
          if (c + 1 < K) tn = H[c] + prow[c + 1];
          const uint32_t h = __vimax3_u16x2(t, E, F[c]);
          H[c] = h;
          const uint32_t hg = h - goe2;
          E = __viaddmax_u16x2(E, nge, hg);
          F[c] = __viaddmax_u16x2(F[c], nge, hg);
          t = tn;
        }
        if (!last) bnd[32] = make_uint2(H[K - 1], E);
        i = 2;
      }

      // ---- main loop: rows i (A) and i+1 (B) together, B one column behind A --------------------
      // Two independent E chains per lane double the instruction-level parallelism: the
      // 3-instruction dependent chain of one cell (VIMNMX3 -> add -> VIADDMNMX, ~14 clk) is
      // overlapped with the other row's.  t = H_diag + S' is formed one column ahead, from the
      // old H value before the cell overwrites it, so no register copy carries the diagonal.
      uint2 na = make_uint2(0u, 0u), nb = make_uint2(0u, 0u);
      if (i < Ls) {
        na = bnd[(size_t)i * 32];
        nb = bnd[(size_t)(i + 1) * 32];
      }
      for (; i < Ls; i += 2) {
        const uint32_t wab = next_word();
        const uint32_t* prow_a = reinterpret_cast<const uint32_t*>(profb + (wab & 0xffffu));
        const uint32_t* prow_b = reinterpret_cast<const uint32_t*>(profb + (wab >> 16));
        const uint2 la = na, lb = nb;
        na = bnd[(size_t)(i + 2) * 32];   // rows of the next iteration (scratch has slack rows)
        nb = bnd[(size_t)(i + 3) * 32];
        uint32_t Ea = la.y, Eb = lb.y;
        uint32_t ta = hdiag + prow_a[0];      // diag of A(0) = H(i-1, j0)
        uint32_t tb = la.x + prow_b[0];       // diag of B(0) = H(i,   j0)
        hdiag = lb.x;                         // H(i+1, j0) for the next pair
        uint32_t ha_last = 0;
#pragma unroll
        for (int c = 0; c <= K; ++c) {
          if (c < K) {  // cell A(c) of row i
            uint32_t tn = 0;
            if (c + 1 < K) tn = H[c] + prow_a[c + 1];
            const uint32_t h = __vimax3_u16x2(ta, Ea, F[c]);
            H[c] = h;
            const uint32_t hg = h - goe2;
            Ea = __viaddmax_u16x2(Ea, nge, hg);
            F[c] = __viaddmax_u16x2(F[c], nge, hg);
            ta = tn;
            if (c == K - 1) ha_last = h;
          }
          if (c >= 1) {  // cell B(c-1) of row i+1
            uint32_t tn = 0;
            if (c < K) tn = H[c - 1] + prow_b[c];   // H(i, c-1), before B overwrites it
            const uint32_t h = __vimax3_u16x2(tb, Eb, F[c - 1]);
            H[c - 1] = h;
            const uint32_t hg = h - goe2;
            Eb = __viaddmax_u16x2(Eb, nge, hg);
            F[c - 1] = __viaddmax_u16x2(F[c - 1], nge, hg);
            tb = tn;
          }
        }
        if (!last) {
          bnd[(size_t)i * 32] = make_uint2(ha_last, Ea);
          bnd[(size_t)(i + 1) * 32] = make_uint2(H[K - 1], Eb);
        }
      }

      // ---- pick H(Ls, L1) / H(Ls, L2) if the query ends inside this strip --------------
      if (L1 > j0 && L1 <= j0 + K) {
        const int c1 = (int)(L1 - 1 - j0);
#pragma unroll
        for (int c = 0; c < K; ++c)
          if (c == c1) res1 = H[c] & 0xffffu;
      }
      if (L2 > j0 && L2 <= j0 + K) {
        const int c2 = (int)(L2 - 1 - j0);
#pragma unroll
        for (int c = 0; c < K; ++c)
          if (c == c2) res2 = H[c] >> 16;
      }
    }

    g16_task_results(p, task, lane, valid, A1, j, L1, L2, Ls, res1, res2);
  }
}

}  // namespace tsq
