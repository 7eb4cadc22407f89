// gotoh32.cuh -- regime 1b: inter-task 32-bit Gotoh kernel for sm_100a.
//
// The strip-mined, two-rows-at-a-time structure of gotoh16.cuh with one alignment per lane in plain
// signed 32-bit words (no packing, no bias, no skew): one warp = one task = (one query) x (32 subjects).
// Used where the packed 16-bit kernel cannot be: (a) gap/score parameters whose range bound leaves
// no room for 16 bits although the sequences are short, and (b) identity-aware scoring
// (SURVEY.md section 8f-2), where the host scales every score and gap penalty by M > max length and
// adds 1 to the score of identical residues, so that max() over the combined keys  score*M + nid
// picks the best score and, among co-optimal alignments, the one with most identities.
//
// Per cell: t = H_diag + S (add), h = VIMNMX3(t, E, F), hg = h - (go+ge) (add),
// E = VIADDMNMX(E, -ge, hg), F = VIADDMNMX(F, -ge, hg): three 32-bit DPX instructions.
// The subject database and the boundary scratch column are those of gotoh16.cuh.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "gotoh16.cuh"

namespace tsq {

// a*one + b with `one` an opaque 1: forces IMAD (FMA pipe) for the diagonal add; left to itself ptxas
// fuses that add into the DPX instruction (VIADDMNMX + VIMNMX instead of VIMNMX3): five ALU-pipe
// instructions per cell instead of four.
__device__ __forceinline__ int32_t g32_add_fma(int32_t a, int32_t one, int32_t b) {
  int32_t d;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
  return d;
}

struct G32Params {
  const uint32_t* dbw;         // subject database of gotoh16.cuh (16-bit profile-row byte offsets)
  const uint32_t* goff;
  const uint8_t* lin;
  const uint32_t* loff;
  const uint32_t* lens;
  const unsigned long long* task_prefix;  // [nq+1] cumulative chunk counts, per query
  unsigned long long* counter;
  const int* cancel;
  int2* bnd;                   // strip boundary scratch: [warp slot][row][lane] (H, E)
  const int32_t* smat;         // (nsym+1) x nsym scores (row nsym = padding = 0)
  int32_t* out;                // keys/scores, packed upper triangle in sorted order
  unsigned long long ntasks;
  uint32_t bnd_rows;
  uint32_t n_total;
  uint32_t lo, hi;             // eligible sorted range [lo, hi)
  uint32_t q_begin, q_end;     // queries of this launch: rows lo+q
  uint32_t nsym;
  int32_t go, ge;
  int32_t one;                 // 1, opaque to the compiler (see g32_add_fma)
};

template <int K, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) gotoh32_kernel(const __grid_constant__ G32Params p) {
  constexpr int STRIDE = G16Cfg<K>::STRIDE;
  extern __shared__ int32_t smem_g32[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const uint32_t nsym = p.nsym;
  const uint32_t sbsz = (nsym + 1) * nsym;
  int32_t* sm = smem_g32;
  int32_t* prof = smem_g32 + ((sbsz + 31) & ~31u) + wib * (nsym * STRIDE);
  for (uint32_t i = threadIdx.x; i < sbsz; i += TPB) sm[i] = p.smat[i];
  __syncthreads();

  const uint32_t gw = blockIdx.x * (TPB / 32) + wib;
  int2* const bnd = p.bnd + (size_t)gw * p.bnd_rows * 32 + lane;
  const int32_t go = p.go, ge = p.ge, goe = p.go + p.ge, nge = -p.ge, one = p.one;

  for (;;) {
    unsigned long long task = 0;
    if (lane == 0) {
      task = atomicAdd(p.counter, 1ULL);
      if (*reinterpret_cast<const volatile int*>(p.cancel) != 0) task = ~0ULL;
    }
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= p.ntasks) break;
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
    const uint32_t A1 = p.lo + q;
    const uint32_t L1 = p.lens[A1];
    const uint8_t* q1 = p.lin + p.loff[A1];
    const uint32_t j = A1 + 1 + chunk * 32 + lane;
    const bool valid = j < p.hi;
    const uint32_t Ls = valid ? p.lens[j] : 0u;
    const uint32_t* dbp = p.dbw + (valid ? (p.goff[j >> 5] + (j & 31)) : 0u);
    const uint32_t nstrips = (L1 + K - 1) / K;
    int32_t res = 0;

    for (uint32_t s = 0; s < nstrips; ++s) {
      const uint32_t j0 = s * K;
      __syncwarp();
      for (int c = lane; c < K; c += 32) {
        const uint32_t col = j0 + c;
        const uint32_t a1 = col < L1 ? q1[col] : nsym;
        const int32_t* r1 = sm + a1 * nsym;
        for (uint32_t b = 0; b < nsym; ++b) prof[b * STRIDE + c] = r1[b];
      }
      __syncwarp();

      int32_t H[K], F[K];
#pragma unroll
      for (int c = 0; c < K; ++c) {
        H[c] = -(go + (int32_t)(j0 + c + 1) * ge);
        F[c] = H[c] - goe;
      }
      int32_t hdiag = j0 == 0 ? 0 : -(go + (int32_t)j0 * ge);
      const bool last = (s + 1 == nstrips);

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
      if (s == 0) {  // column 0 of the matrix as the first strip's left boundary
        for (uint32_t r0 = 1; r0 <= Ls; ++r0) {
          const int32_t hl = -(go + (int32_t)r0 * ge);
          bnd[(size_t)r0 * 32] = make_int2(hl, hl - goe);
        }
      }
      uint32_t i = 1;
      if (Ls & 1u) {
        const int32_t* prow = reinterpret_cast<const int32_t*>(profb + (next_word() >> 16));
        const int2 lb = bnd[32];
        int32_t E = lb.y;
        int32_t t = g32_add_fma(hdiag, one, prow[0]);
        hdiag = lb.x;
#pragma unroll
        for (int c = 0; c < K; ++c) {
          int32_t tn = 0;
          if (c + 1 < K) tn = g32_add_fma(H[c], one, prow[c + 1]);
          const int32_t h = __vimax3_s32(t, E, F[c]);
          H[c] = h;
          const int32_t hg = h - goe;
          E = __viaddmax_s32(E, nge, hg);
          F[c] = __viaddmax_s32(F[c], nge, hg);
          t = tn;
        }
        if (!last) bnd[32] = make_int2(H[K - 1], E);
        i = 2;
      }
      int2 na = make_int2(0, 0), nb = make_int2(0, 0);
      if (i < Ls) {
        na = bnd[(size_t)i * 32];
        nb = bnd[(size_t)(i + 1) * 32];
      }
      for (; i < Ls; i += 2) {
        const uint32_t wab = next_word();
        const int32_t* prow_a = reinterpret_cast<const int32_t*>(profb + (wab & 0xffffu));
        const int32_t* prow_b = reinterpret_cast<const int32_t*>(profb + (wab >> 16));
        const int2 la = na, lb = nb;
        na = bnd[(size_t)(i + 2) * 32];
        nb = bnd[(size_t)(i + 3) * 32];
        int32_t Ea = la.y, Eb = lb.y;
        int32_t ta = g32_add_fma(hdiag, one, prow_a[0]);
        int32_t tb = g32_add_fma(la.x, one, prow_b[0]);
        hdiag = lb.x;
        int32_t ha_last = 0;
#pragma unroll
        for (int c = 0; c <= K; ++c) {
          if (c < K) {
            int32_t tn = 0;
            if (c + 1 < K) tn = g32_add_fma(H[c], one, prow_a[c + 1]);
            const int32_t h = __vimax3_s32(ta, Ea, F[c]);
            H[c] = h;
            const int32_t hg = h - goe;
            Ea = __viaddmax_s32(Ea, nge, hg);
            F[c] = __viaddmax_s32(F[c], nge, hg);
            ta = tn;
            if (c == K - 1) ha_last = h;
          }
          if (c >= 1) {
            int32_t tn = 0;
            if (c < K) tn = g32_add_fma(H[c - 1], one, prow_b[c]);
            const int32_t h = __vimax3_s32(tb, Eb, F[c - 1]);
            H[c - 1] = h;
            const int32_t hg = h - goe;
            Eb = __viaddmax_s32(Eb, nge, hg);
            F[c - 1] = __viaddmax_s32(F[c - 1], nge, hg);
            tb = tn;
          }
        }
        if (!last) {
          bnd[(size_t)i * 32] = make_int2(ha_last, Ea);
          bnd[(size_t)(i + 1) * 32] = make_int2(H[K - 1], Eb);
        }
      }
      if (L1 > j0 && L1 <= j0 + K) {
        const int c1 = (int)(L1 - 1 - j0);
#pragma unroll
        for (int c = 0; c < K; ++c)
          if (c == c1) res = H[c];
      }
    }
    if (valid) p.out[tri_index(A1, j, p.n_total)] = res;
  }
}

}  // namespace tsq
