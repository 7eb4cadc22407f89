// wave32.cuh -- regime 2: intra-task anti-diagonal wavefront Gotoh kernel, 32-bit, sm_100a.
//
// For pairs the packed 16-bit kernel cannot take (a sequence longer than its range bound,
// e.g. 10-30 kb genomes: BASELINE.json configs[3]).  One warp = one pair.  The shorter
// sequence lies along the columns, the longer one along the rows.  Columns are processed in
// passes of 32*KW: lane l owns KW adjacent columns and keeps their H and F in registers.  The
// warp sweeps the rows as an anti-diagonal wavefront, two rows per lane per step (the same
// two-row interleave as gotoh16.cuh, for instruction-level parallelism): at step s lane l
// works on rows 2(s-l)+1 and 2(s-l)+2.  The right edge (H, E) of a lane's column block and the
// two subject letters travel to lane l+1 by warp shuffle; lane 31's right edge is the only
// thing a pass writes to memory (8 bytes per row per 32*KW columns) and lane 0 of the next
// pass reads it back.
//
// Scores: per-warp query profile in shared memory, laid out [letter][column-in-lane][lane]:
// lane l always reads bank l, so the lookup is conflict-free whatever letters the lanes hold.
//
// Per cell: t = H_diag + S (add), h = VIMNMX3(t, E, F), hg = h - (go+ge) (add),
// E = VIADDMNMX(E, -ge, hg), F = VIADDMNMX(F, -ge, hg): three 32-bit DPX instructions.
// -inf is replaced by the finite seeds H - go - ge exactly as in the packed kernel.
// Spec: SURVEY.md section 8c.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "gotoh16.cuh"

namespace tsq {

struct W32Params {
  const uint8_t* lin;          // residues, linear, sorted order
  const uint32_t* loff;        // start of each sorted sequence
  const uint32_t* lens;        // sorted lengths
  const uint2* pairs;          // tasks: (i, j) sorted indices, i < j, biggest first
  unsigned long long* counter; // dynamic task cursor
  const int* cancel;           // device flag (set by a side-stream copy): != 0 stops task fetching
  int2* bnd;                   // pass boundary scratch: [warp slot][row] (H, E)
  const int32_t* smat;         // (nsym+1) x nsym scores (row nsym = padding = 0)
  int32_t* out;                // scores, packed upper triangle in sorted order
  unsigned long long ntasks;
  uint32_t bnd_rows;
  uint32_t n_total;
  uint32_t nsym;
  int32_t go, ge;
  int32_t one;                 // 1, opaque to the compiler (see add_fma_pipe)
};

// a*one + b with `one` an opaque 1: forces IMAD (FMA pipe) for the diagonal add.  Left to itself
// ptxas fuses that add into the DPX instruction (VIADDMNMX + VIMNMX instead of VIMNMX3), which
// puts five ALU-pipe instructions in a cell; this way it is four (r01 profile: ALU pipe 73 %).
__device__ __forceinline__ int32_t add_fma_pipe(int32_t a, int32_t one, int32_t b) {
  int32_t d;
  asm("mad.lo.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
  return d;
}

template <int KW, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) wave32_kernel(const __grid_constant__ W32Params p) {
  constexpr int PW = 32 * KW;  // columns per pass
  extern __shared__ int32_t smem32[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const uint32_t nsym = p.nsym;
  const uint32_t sbsz = (nsym + 1) * nsym;
  int32_t* sm = smem32;
  int32_t* prof = smem32 + ((sbsz + 31) & ~31u) + (size_t)wib * nsym * PW;
  for (uint32_t i = threadIdx.x; i < sbsz; i += TPB) sm[i] = p.smat[i];
  __syncthreads();

  const uint32_t gw = blockIdx.x * (TPB / 32) + wib;
  int2* const bnd = p.bnd + (size_t)gw * p.bnd_rows;
  const int32_t go = p.go, ge = p.ge, goe = p.go + p.ge, nge = -p.ge;
  const int32_t one = p.one;

  for (;;) {
    unsigned long long task = 0;
    if (lane == 0) {
      task = atomicAdd(p.counter, 1ULL);
      if (*reinterpret_cast<const volatile int*>(p.cancel) != 0) task = ~0ULL;   // "Stop" pressed
    }
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= p.ntasks) break;
    const uint2 pr = p.pairs[task];
    // columns <- shorter (sorted index pr.x), rows <- longer (pr.y)
    const uint32_t n = p.lens[pr.x], m = p.lens[pr.y];
    const uint8_t* qa = p.lin + p.loff[pr.x];
    const uint8_t* sb = p.lin + p.loff[pr.y];
    const uint32_t npass = (n + PW - 1) / PW;
    const uint32_t npairs_rows = (m + 1) / 2;
    int32_t result = 0;

    for (uint32_t pass = 0; pass < npass; ++pass) {
      const uint32_t pcol0 = pass * PW;
      const bool firstp = (pass == 0), lastp = (pass + 1 == npass);
      // ---- profile of this pass: prof[b][c][l] = S(A[pcol0 + l*KW + c], b) -------------------
      __syncwarp();
      for (int c2 = 0; c2 < KW; ++c2) {  // lane l fills its own columns: stores hit bank l
        const uint32_t col = pcol0 + lane * KW + c2;
        const uint32_t a = col < n ? qa[col] : nsym;
        const int32_t* srow = sm + a * nsym;
        for (uint32_t b = 0; b < nsym; ++b) prof[b * PW + c2 * 32 + lane] = srow[b];
      }
      __syncwarp();
      const int32_t* myprof = prof + lane;

      // ---- row 0 of this lane's column block -------------------------------------------------
      const int32_t col0 = (int32_t)(pcol0 + lane * KW);  // columns col0+1 .. col0+KW (1-based)
      int32_t H[KW], F[KW];
#pragma unroll
      for (int c = 0; c < KW; ++c) {
        H[c] = -(go + (col0 + c + 1) * ge);
        F[c] = H[c] - goe;
      }
      int32_t hdiag = col0 == 0 ? 0 : -(go + col0 * ge);

      // values handed to the right-hand neighbour (produced in the previous step)
      int32_t oHa = 0, oEa = 0, oHb = 0, oEb = 0;
      uint32_t olet = 0;
      // lane 0: prefetched boundary rows and subject letters of the next step
      int2 nba = make_int2(0, 0), nbb = make_int2(0, 0);
      uint32_t nlet = 0;
      if (lane == 0) {
        nlet = (uint32_t)sb[0] | ((m > 1 ? (uint32_t)sb[1] : 0u) << 8);
        if (!firstp) {
          nba = bnd[1];
          nbb = bnd[2];
        }
      }
      const uint32_t nsteps = npairs_rows + 31;
      for (uint32_t s = 0; s < nsteps; ++s) {
        int32_t iHa = __shfl_up_sync(0xffffffffu, oHa, 1);
        int32_t iEa = __shfl_up_sync(0xffffffffu, oEa, 1);
        int32_t iHb = __shfl_up_sync(0xffffffffu, oHb, 1);
        int32_t iEb = __shfl_up_sync(0xffffffffu, oEb, 1);
        uint32_t let = __shfl_up_sync(0xffffffffu, olet, 1);
        if (lane == 0) {
          const uint32_t ra = 2 * s + 1;  // rows of this step for lane 0
          let = nlet;
          if (firstp) {
            iHa = -(go + (int32_t)ra * ge);
            iEa = iHa - goe;
            iHb = iHa - ge;
            iEb = iHb - goe;
          } else {
            iHa = nba.x; iEa = nba.y; iHb = nbb.x; iEb = nbb.y;
          }
          if (s + 1 < npairs_rows) {  // prefetch the next step's inputs
            const uint32_t r2 = ra + 2;  // 1-based row of next step's A
            nlet = (uint32_t)sb[r2 - 1] | ((r2 < m ? (uint32_t)sb[r2] : 0u) << 8);
            if (!firstp) {
              nba = bnd[r2];
              nbb = bnd[r2 + 1];
            }
          }
        }
        const int32_t ps = (int32_t)s - lane;
        const bool active = ps >= 0 && (uint32_t)ps < npairs_rows;
        olet = let;
        if (active) {
          const uint32_t ra = 2 * (uint32_t)ps + 1;
          const int32_t* prow_a = myprof + (let & 0xffu) * PW;
          if (ra == m) {
            // ---- last row of an odd-length subject: a single row ------------------------------
            int32_t E = iEa;
            int32_t t = add_fma_pipe(hdiag, one, prow_a[0]);
            hdiag = iHa;
#pragma unroll
            for (int c = 0; c < KW; ++c) {
              int32_t tn = 0;
              if (c + 1 < KW) tn = add_fma_pipe(H[c], one, prow_a[(c + 1) * 32]);
              const int32_t h = __vimax3_s32(t, E, F[c]);
              H[c] = h;
              const int32_t hg = h - goe;
              E = __viaddmax_s32(E, nge, hg);
              F[c] = __viaddmax_s32(F[c], nge, hg);
              t = tn;
            }
            oHa = H[KW - 1];
            oEa = E;
            if (lane == 31 && !lastp) bnd[ra] = make_int2(oHa, oEa);
          } else {
            // ---- rows ra (A) and ra+1 (B), B one column behind A --------------------------------
            const int32_t* prow_b = myprof + ((let >> 8) & 0xffu) * PW;
            int32_t Ea = iEa, Eb = iEb;
            int32_t ta = add_fma_pipe(hdiag, one, prow_a[0]);
            int32_t tb = add_fma_pipe(iHa, one, prow_b[0]);
            hdiag = iHb;
            int32_t ha_last = 0;
#pragma unroll
            for (int c = 0; c <= KW; ++c) {
              if (c < KW) {
                int32_t tn = 0;
                if (c + 1 < KW) tn = add_fma_pipe(H[c], one, prow_a[(c + 1) * 32]);
                const int32_t h = __vimax3_s32(ta, Ea, F[c]);
                H[c] = h;
                const int32_t hg = h - goe;
                Ea = __viaddmax_s32(Ea, nge, hg);
                F[c] = __viaddmax_s32(F[c], nge, hg);
                ta = tn;
                if (c == KW - 1) ha_last = h;
              }
              if (c >= 1) {
                int32_t tn = 0;
                if (c < KW) tn = add_fma_pipe(H[c - 1], one, prow_b[c * 32]);
                const int32_t h = __vimax3_s32(tb, Eb, F[c - 1]);
                H[c - 1] = h;
                const int32_t hg = h - goe;
                Eb = __viaddmax_s32(Eb, nge, hg);
                F[c - 1] = __viaddmax_s32(F[c - 1], nge, hg);
                tb = tn;
              }
            }
            oHa = ha_last; oEa = Ea; oHb = H[KW - 1]; oEb = Eb;
            if (lane == 31 && !lastp) {
              bnd[ra] = make_int2(oHa, oEa);
              bnd[ra + 1] = make_int2(oHb, oEb);
            }
          }
        }
      }
      // ---- H(m, n) sits in the lane that owns column n, after its last row ---------------------
      if (lastp) {
        const int32_t cn = (int32_t)n - 1 - col0;  // 0-based index inside this lane's block
        int32_t mine = 0;
        if (cn >= 0 && cn < KW) {
#pragma unroll
          for (int c = 0; c < KW; ++c)
            if (c == cn) mine = H[c];
        }
        const int owner = (int)((n - 1 - pcol0) / KW);
        result = __shfl_sync(0xffffffffu, mine, owner);
      }
      __syncwarp();
    }
    if (lane == 0) p.out[tri_index(pr.x, pr.y, p.n_total)] = result;
  }
}

}  // namespace tsq
