// traceback.cuh -- one optimal global alignment of ONE pair, with its path (SURVEY.md 8f-2:
// "also enables emitting pairwise alignments").  Not the all-vs-all hot path: the matrix kernels
// never keep directions; this kernel recomputes the single pair the caller asks to see.
//
// Anti-diagonal sweep: cell (i, j) on diagonal d = i + j needs (i, j-1) and (i-1, j) from diagonal
// d-1 and (i-1, j-1) from d-2, so a diagonal is embarrassingly parallel and the threads meet at one
// barrier per diagonal.  Short pairs run on one CTA (__syncthreads); pairs whose diagonals are longer
// than a CTA run on one thread-block CLUSTER of 8 CTAs (8 192 threads on 8 SMs) that meets at the
// hardware cluster barrier (barrier.cluster arrive.release / wait.acquire), exchanging the rolling
// diagonals through L2 (ld.cg / st.cg: never a stale L1 line of another SM).
// Three rolling diagonals of H and two of E and F live in an L2-resident scratch indexed by the
// row i; one direction byte per cell goes to global memory, DIAGONAL-major (byte (d, i - ilo(d)) at
// d * ld + i - ilo(d), ld = min(m, n) + 1) so that a warp's 32 bytes are one 32-byte sector:
//     bits 0-1  where H(i,j) came from: 0 diagonal, 1 E (gap in the row sequence), 2 F
//     bit  2    E(i,j) opened from H(i,j-1) (ties prefer opening)      bit 3  same for F
// Ties in H prefer the diagonal, then E.  The CPU oracle (tsq_oracle_traceback) applies the same
// rules, so the emitted strings are identical byte for byte, not merely equally good.
// Thread 0 then walks the path back from (m, n) and writes the two gapped rows reversed.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tsq {

struct TbParams {
  const uint8_t* a;       // row sequence (encoded symbols), m residues
  const uint8_t* b;       // column sequence, n residues
  uint32_t m, n;
  const int32_t* smat;    // nsym x nsym plain scores
  uint32_t nsym;
  int32_t go, ge;
  int32_t* diag;          // scratch: 7 * (m + 1) ints
  uint8_t* dir;           // (m + n + 1) x (min(m, n) + 1) direction bytes, diagonal-major
  uint8_t* out_a;         // m + n bytes each: the gapped rows, symbols 0..nsym-1, 0xff = gap
  uint8_t* out_b;
  int32_t* info;          // [0] columns, [1] score
};

constexpr int TB_THREADS = 1024;
constexpr int TB_CLUSTER = 8;      // CTAs of the cluster variant (portable cluster size)

#ifdef TSQ_DEVICE_IMPL
// NCTA = 1: one CTA.  NCTA = TB_CLUSTER: launched as ONE cluster of NCTA CTAs (grid = NCTA).
template <int NCTA>
__global__ void __launch_bounds__(TB_THREADS) traceback_kernel(const __grid_constant__ TbParams p) {
  constexpr int NT = TB_THREADS * NCTA;
  const int tid = (int)threadIdx.x + TB_THREADS * (NCTA > 1 ? (int)blockIdx.x : 0);
  auto ld_ = [](const int32_t* q) -> int32_t { return NCTA > 1 ? __ldcg(q) : *q; };
  auto st_ = [](int32_t* q, int32_t v) { if (NCTA > 1) __stcg(q, v); else *q = v; };
  auto meet = []() {
    if (NCTA > 1) {
      asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
      __syncthreads();
    }
  };
  const int m = (int)p.m, n = (int)p.n;
  const int stride = m + 1;
  const int32_t NEG = -(1 << 29);
  const int32_t ge = p.ge, goe = p.go + p.ge;
  int32_t* Hc = p.diag;
  int32_t* Hp1 = Hc + stride;
  int32_t* Hp2 = Hp1 + stride;
  int32_t* Ec = Hp2 + stride;
  int32_t* Ep1 = Ec + stride;
  int32_t* Fc = Ep1 + stride;
  int32_t* Fp1 = Fc + stride;
  const size_t ld = (size_t)(m < n ? m : n) + 1;
  for (int d = 0; d <= m + n; ++d) {
    const int ilo = d > n ? d - n : 0;
    const int ihi = d < m ? d : m;
    uint8_t* const drow = p.dir + (size_t)d * ld - ilo;
    for (int i = ilo + tid; i <= ihi; i += NT) {
      const int j = d - i;
      int32_t H, E, F;
      uint32_t code;
      if (i == 0 && j == 0) {
        H = 0; E = NEG; F = NEG; code = 0;
      } else if (i == 0) {
        H = E = -p.go - j * ge; F = NEG; code = 1u | (j == 1 ? 4u : 0u);
      } else if (j == 0) {
        H = F = -p.go - i * ge; E = NEG; code = 2u | (i == 1 ? 8u : 0u);
      } else {
        const int32_t hl = ld_(Hp1 + i), hu = ld_(Hp1 + i - 1), hd = ld_(Hp2 + i - 1);
        const int32_t e_ext = ld_(Ep1 + i) - ge, e_open = hl - goe;
        const int32_t f_ext = ld_(Fp1 + i - 1) - ge, f_open = hu - goe;
        const int32_t dg = hd + __ldg(p.smat + (uint32_t)__ldg(p.a + i - 1) * p.nsym + __ldg(p.b + j - 1));
        const bool eo = e_open >= e_ext, fo = f_open >= f_ext;
        E = eo ? e_open : e_ext;
        F = fo ? f_open : f_ext;
        H = max(dg, max(E, F));
        code = (H == dg ? 0u : (H == E ? 1u : 2u)) | (eo ? 4u : 0u) | (fo ? 8u : 0u);
      }
      st_(Hc + i, H); st_(Ec + i, E); st_(Fc + i, F);
      drow[i] = (uint8_t)code;
    }
    meet();   // diagonal d complete and visible to every thread before anyone starts d + 1
    int32_t* t = Hp2; Hp2 = Hp1; Hp1 = Hc; Hc = t;
    t = Ep1; Ep1 = Ec; Ec = t;
    t = Fp1; Fp1 = Fc; Fc = t;
  }
  if (tid == 0) {
    // after the last rotation the corner (m, n) sits in Hp1[m]
    int i = m, j = n, state = 0;
    uint32_t k = 0;
    while (i > 0 || j > 0) {
      if (i == 0) { p.out_a[k] = 0xff; p.out_b[k] = p.b[j - 1]; --j; ++k; continue; }
      if (j == 0) { p.out_a[k] = p.a[i - 1]; p.out_b[k] = 0xff; --i; ++k; continue; }
      const int d = i + j;
      const uint32_t code = __ldcg(p.dir + (size_t)d * ld + (i - (d > n ? d - n : 0)));
      if (state == 0) {
        const uint32_t src = code & 3u;
        if (src == 0) { p.out_a[k] = p.a[i - 1]; p.out_b[k] = p.b[j - 1]; --i; --j; ++k; }
        else state = (int)src;
      } else if (state == 1) {
        p.out_a[k] = 0xff; p.out_b[k] = p.b[j - 1]; ++k;
        if (code & 4u) state = 0;
        --j;
      } else {
        p.out_a[k] = p.a[i - 1]; p.out_b[k] = 0xff; ++k;
        if (code & 8u) state = 0;
        --i;
      }
    }
    p.info[0] = (int32_t)k;
    p.info[1] = ld_(Hp1 + m);
  }
}
#endif  // TSQ_DEVICE_IMPL

}  // namespace tsq
