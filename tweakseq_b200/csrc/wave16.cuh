// wave16.cuh -- regime 2, packed: intra-task anti-diagonal wavefront Gotoh kernel with two
// alignments per 32-bit word (u16x2 DPX) and a moving base, sm_100a.
//
// Same wavefront as wave32.cuh (one warp per task, lane l owns KW adjacent columns, right edges
// handed to lane l+1 by warp shuffle; here FOUR rows -- two interleaved row pairs -- per lane per
// step, which halves the per-step hand-off overhead per cell), but a task is TWO queries
// (A1, A2, along the columns, one per 16-bit half) against one long subject (along the rows), so
// every DPX instruction advances two alignments -- what the packed inter-task kernel
// (gotoh16.cuh) does for sequences short enough to fit 16 bits outright.
//
// Long sequences do not fit 16 bits (|H| reaches 10^5), but the CELLS A WARP HOLDS AT ONE TIME do:
// by the Lipschitz property of alignment matrices (|H(i,j)-H(i,j-1)|, |H(i,j)-H(i-1,j)| <=
// max S + go + ge), everything in flight -- 32*KW columns by ~124+4R rows -- lies within a window
// D = (32*KW + 4*31 + 4*R + 16) * L of one reference cell, and E, F lie within go+ge of an H.
// The kernel therefore stores v - base(half) in unsigned 16 bits, where base is a warp-uniform
// 32-bit number per half that is re-centred every R steps on a reference cell (lane 16's first
// column): all live registers are shifted by the same packed constant, base absorbs the shift.
// max() and +constant commute with a uniform shift, so the arithmetic is exactly the 32-bit
// recurrence as long as nothing leaves [0, 65535]; the host only selects this kernel when
// D <= 30000 (tsq_api.cpp: wave16_ok) and falls back to wave32 otherwise.  Values that cross a
// pass boundary (lane 31 -> memory -> lane 0 of the next pass) are stored relative together with the
// base in force (one base pair per step) and re-based on load; final scores are made absolute.  As in gotoh16.cuh the stored value is skewed by
// delta*(i+j) so that substitution scores are non-negative and the two plain adds of a cell
// cannot carry between the halves; the skew is just part of what base tracks.
//
// The subject sequence is streamed through shared memory in 1 KB tiles by TMA bulk copies
// (cp.async.bulk + mbarrier, tma_stage.cuh) into a two-slot ring per warp; every lane reads its own
// two letters per step from the ring, so no letter travels by shuffle.
//
// Spec: SURVEY.md section 8c.  Exactness is tested bit for bit against the oracle and against
// wave32 (tests/test_gpu_parity.py).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "gotoh16.cuh"
#include "tma_stage.cuh"

namespace tsq {

struct W16Params {
  const uint8_t* lin;          // residues, linear, sorted order
  const uint32_t* loff;        // start of each sorted sequence
  const uint32_t* lens;        // sorted lengths
  const uint4* tasks;          // (i1, i2, j, 0): queries i1 <= i2 (i2 == i1: single), subject j
  unsigned long long* counter; // dynamic task cursor
  const int* cancel;           // device flag (set by a side-stream copy): != 0 stops task fetching
  int* fault;                  // device fault word (cancel + 1): raised when a TMA tile barrier times out
  uint32_t inject_fault;       // test hook: task 0 arms its first tile barrier without issuing the copy
  uint2* bnd;                  // pass boundary scratch per warp slot (w16_slot_elems): one record per step
  const uint32_t* sbias;       // (nsym+1) x nsym biased scores S' = S + 2*delta (row nsym = 0)
  int32_t* out;                // scores, packed upper triangle in sorted order
  unsigned long long ntasks;
  uint32_t bnd_rows;
  uint32_t n_total;
  uint32_t nsym;
  int32_t delta;               // skew per anti-diagonal
  int32_t go;                  // gap open
  int32_t gep;                 // ge' = ge - delta
  int32_t goep;                // goe' = go + ge - delta
  uint32_t negge2;             // (-ge' mod 2^16) in both halves
};

// Pass-boundary scratch of one warp: one 48-byte RECORD per step -- the (H, E) of the step's four rows 4q+1 .. 4q+4
// as two 16-byte vectors, then (base_lo, base_hi) in force when they were written, 8 bytes of padding.  Lane 31
// writes record s - 31 while lane 0 of the next pass reads record s + 1: ONE running pointer serves both, every
// access is that pointer plus a constant.  Size in uint2 elements (the type of W16Params::bnd).
constexpr uint32_t kW16RecBytes = 48;
__host__ __device__ inline size_t w16_slot_elems(uint32_t bnd_rows) {
  return ((size_t)(bnd_rows + 3) / 4 + 4) * (kW16RecBytes / 8);
}

__device__ __forceinline__ uint32_t pack_rel(int32_t a, int32_t base_lo, int32_t base_hi) {
  return ((uint32_t)(a - base_lo) & 0xffffu) | ((uint32_t)(a - base_hi) << 16);
}

__device__ __forceinline__ uint32_t comp4(const uint4& v, int k) {   // k is a compile-time constant after unrolling
  return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w;
}

template <int KW, int TPB, int MINB, uint32_t NGE>
__global__ void __launch_bounds__(TPB, MINB) wave16_kernel(const __grid_constant__ W16Params p) {
  constexpr int PW = 32 * KW;        // columns per pass
  constexpr uint32_t RB = 16;        // re-centre the base every RB steps (4*RB rows)
  constexpr int32_t CENTER = 32768;
  constexpr uint32_t TILE = 1024;    // subject bytes per TMA tile (two tiles per warp)
  constexpr uint32_t TSTEPS = TILE / 4;  // steps of lane 0 per tile (4 rows per step)
  extern __shared__ uint32_t smem16[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const uint32_t nsym = p.nsym;
  const uint32_t sbsz = (nsym + 1) * nsym;
  uint32_t* sb = smem16;
  uint32_t* prof = smem16 + ((sbsz + 31) & ~31u) + (size_t)wib * nsym * PW;
  // after the profiles: per warp two 1 KB subject tiles and their two mbarriers
  uint8_t* const tile_base = reinterpret_cast<uint8_t*>(smem16 + ((sbsz + 31) & ~31u) + (size_t)(TPB / 32) * nsym * PW);
  uint8_t* const stile = tile_base + (size_t)wib * (2 * TILE + 16);
  uint64_t* const tbar = reinterpret_cast<uint64_t*>(stile + 2 * TILE);
  uint32_t tph0 = 0, tph1 = 0;       // phase parity of the two tile barriers (warp-uniform)
  if (lane == 0) {
    mbar_init(&tbar[0], 1);
    mbar_init(&tbar[1], 1);
    fence_mbar_init();
  }
  for (uint32_t i = threadIdx.x; i < sbsz; i += TPB) sb[i] = p.sbias[i];
  __syncthreads();

  const uint32_t gw = blockIdx.x * (TPB / 32) + wib;
  char* const bslot = reinterpret_cast<char*>(p.bnd + (size_t)gw * w16_slot_elems(p.bnd_rows));   // record q at 48 q
  const uint32_t nge = NGE ? NGE : p.negge2;
  const int32_t go = p.go, gep = p.gep, goep = p.goep;
  const uint32_t goe2 = (uint32_t)goep * 0x10001u;
  const uint32_t gep2 = (uint32_t)gep * 0x10001u;

  for (;;) {
    unsigned long long task = 0;
    if (lane == 0) {
      task = atomicAdd(p.counter, 1ULL);
      if (*reinterpret_cast<const volatile int*>(p.cancel) != 0) task = ~0ULL;   // "Stop" pressed
      if (*reinterpret_cast<const volatile int*>(p.fault) != 0) task = ~0ULL;    // a warp hit a device fault: drain
    }
    task = __shfl_sync(0xffffffffu, task, 0);
    if (task >= p.ntasks) break;
    const uint4 tk = p.tasks[task];
    const uint32_t n1 = p.lens[tk.x], n2 = p.lens[tk.y], m = p.lens[tk.z];  // n1 <= n2
    const uint8_t* q1 = p.lin + p.loff[tk.x];
    const uint8_t* q2 = p.lin + p.loff[tk.y];
    const uint8_t* sq = p.lin + p.loff[tk.z];
    const uint32_t npass = (n2 + PW - 1) / PW;
    const uint32_t pass1 = (n1 - 1) / PW;   // pass in which query 1 ends
    const uint32_t nquads = (m + 3) / 4;    // steps per lane: four rows each
    int32_t res_lo = 0, res_hi = 0;
    bool dead = false;   // a tile barrier timed out (warp-uniform): abandon the task, the fault word is raised

    for (uint32_t pass = 0; pass < npass && !dead; ++pass) {
      const uint32_t pcol0 = pass * PW;
      const bool firstp = (pass == 0), lastp = (pass + 1 == npass);
      // ---- packed profile of this pass: prof[b][c][l] = S'(A1[col],b) | S'(A2[col],b) << 16 ------
      __syncwarp();
      for (int c2 = 0; c2 < KW; ++c2) {
        const uint32_t col = pcol0 + lane * KW + c2;
        const uint32_t a1 = col < n1 ? q1[col] : nsym;
        const uint32_t a2 = col < n2 ? q2[col] : nsym;
        const uint32_t* r1 = sb + a1 * nsym;
        const uint32_t* r2 = sb + a2 * nsym;
        // layout [letter][column group of 4][lane][4]: a lane's four adjacent columns are one 16-byte vector,
        // and the 8 lanes of a quarter-warp cover 128 contiguous bytes -> one conflict-free LDS.128 per four
        // packed cells, whatever the letters
        uint32_t* dst = prof + (c2 >> 2) * 128 + lane * 4 + (c2 & 3);
        for (uint32_t b = 0; b < nsym; ++b) dst[b * PW] = r1[b] | (r2[b] << 16);
      }
      __syncwarp();
      const uint32_t* myprof = prof + lane * 4;

      // ---- row 0 of this lane's column block; absolute skewed A(0,j) = -go - j*ge' ----------------
      const int32_t col0 = (int32_t)(pcol0 + lane * KW);
      int32_t base_lo = (-go - (int32_t)(pcol0 + PW / 2) * gep) - CENTER;
      int32_t base_hi = base_lo;
      uint32_t H[KW], F[KW];
#pragma unroll
      for (int c = 0; c < KW; ++c) {
        H[c] = pack_rel(-go - (col0 + c + 1) * gep, base_lo, base_hi);
        F[c] = H[c] - goe2;
      }
      uint32_t hdiag = pack_rel(col0 == 0 ? 0 : -go - col0 * gep, base_lo, base_hi);
      // first pass: column 0 of the matrix, A(r,0) = -go - r*ge', kept packed and relative for row 4s+1
      uint32_t colH = pack_rel(-go - gep, base_lo, base_hi);

      uint32_t oH[4] = {0u, 0u, 0u, 0u}, oE[4] = {0u, 0u, 0u, 0u};  // right edge of this step's four rows: to lane l+1
      uint4 nb0 = make_uint4(0u, 0u, 0u, 0u), nb1 = nb0;   // lane 0, later passes: (H, E) of the next step's rows 1,2 / 3,4
      int2 nbase = make_int2(0, 0);
      // ---- subject tiles: 2 x 1 KB ring in shared memory, filled by TMA bulk copies ----------------
      // Tile k covers rows [1024k, 1024k+1024) and lives in ring slot k & 1.  All tile bookkeeping is
      // WARP-UNIFORM (it depends on the step counter only): every lane polls the mbarrier, one
      // elected lane issues the copies.  (A wait loop inside a one-lane branch makes ptxas treat the
      // warp as possibly divergent at the shuffles below and route them through the slow
      // WARPSYNC.COLLECTIVE path: measured 1.8x slower.)  Every lane then reads its own four letters
      // per step straight from the ring: lane l, step s -> bytes 4(s-l) .. 4(s-l)+3.
      const uint32_t ntiles = (m + TILE - 1) / TILE;
      if (lane == 0) {
        fence_proxy_async_smem();
        mbar_arrive_expect_tx(&tbar[0], TILE);
        if (!(p.inject_fault && task == 0 && pass == 0)) bulk_copy_g2s(stile, sq, TILE, &tbar[0]);
        if (ntiles > 1) {
          mbar_arrive_expect_tx(&tbar[1], TILE);
          bulk_copy_g2s(stile + TILE, sq + TILE, TILE, &tbar[1]);
        }
        if (!firstp) {
          nb0 = *reinterpret_cast<const uint4*>(bslot);
          nb1 = *reinterpret_cast<const uint4*>(bslot + 16);
          nbase = *reinterpret_cast<const int2*>(bslot + 32);
        }
      }
      dead = !mbar_wait_warp(&tbar[0], tph0, p.fault);
      tph0 ^= 1u;
      if (dead) break;
      uint32_t nlet = lane == 0 ? *reinterpret_cast<const uint32_t*>(stile) : 0u;
      const uint32_t nsteps = nquads + 31;
      const uint32_t nfull = m / 4;                   // steps of a lane in which all four rows exist
      const uint32_t ps_last = (m & 3u) == 0 ? nfull - 1 : 0xffffffffu;   // the full step that holds the subject's last row
      int32_t ps = -lane;                             // this lane's own step: rows 4 ps + 1 .. 4 ps + 4
      uint32_t roff = (uint32_t)(4 * (1 - lane)) & (2 * TILE - 1);   // ring offset of this lane's NEXT four letters
      const char* const myprofb = reinterpret_cast<const char*>(prof) + lane * 16;   // this lane's profile vectors, as bytes
      char* rp = bslot;                               // record s of the boundary scratch (warp-uniform)
      for (uint32_t s = 0; s < nsteps; ++s) {
        const uint32_t let4 = nlet;
        char* const wp = rp - 31 * (int)kW16RecBytes;  // record s - 31: what lane 31 writes in this step
        // ---- everything that happens once in many steps, behind ONE test that is false on 15 steps of 16 (uniform):
        //      re-centring the base, waiting for the next subject tile, refilling the slot of the one before ----------
        if ((s & (RB - 1)) == 0) {
          if (s != 0 && s - 1 < nquads) {
            // re-centre the base on lane 16's first column (all lanes; the state is that of the end of step s - 1)
            const uint32_t ref = __shfl_sync(0xffffffffu, H[0], 16);
            const int32_t sh_lo = (int32_t)(ref & 0xffffu) - CENTER;
            const int32_t sh_hi = (int32_t)(ref >> 16) - CENTER;
            const uint32_t shift2 = (uint32_t)(sh_hi * 65536 + sh_lo);
#pragma unroll
            for (int c = 0; c < KW; ++c) {
              H[c] -= shift2;
              F[c] -= shift2;
            }
            hdiag -= shift2;
            colH -= shift2;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              oH[k] -= shift2;
              oE[k] -= shift2;
            }
            base_lo += sh_lo;
            base_hi += sh_hi;
          }
          const uint32_t tix = s / TSTEPS, tin = s & (TSTEPS - 1);
          // lane 0 enters tile tix + 1 in 16 steps (its letters are fetched one step ahead): the copy, issued 176
          // steps ago (or at the start of the pass), has to have landed
          if (tin == TSTEPS - 16 && tix + 1 < ntiles) {
            if ((tix + 1) & 1u) { dead = !mbar_wait_warp(&tbar[1], tph1, p.fault); tph1 ^= 1u; }
            else                { dead = !mbar_wait_warp(&tbar[0], tph0, p.fault); tph0 ^= 1u; }
            if (dead) break;
          }
          // lane 31 left tile tix - 1 at step 256 tix + 31: its slot is refilled with tile tix + 1
          if (tin == 64 && tix >= 1 && tix + 1 < ntiles && lane == 0) {
            fence_proxy_async_smem();
            uint64_t* br = &tbar[(tix + 1) & 1u];
            mbar_arrive_expect_tx(br, TILE);
            bulk_copy_g2s(stile + ((tix + 1) & 1u) * TILE, sq + (size_t)(tix + 1) * TILE, TILE, br);
          }
        }
        {
          // lanes that have not started yet must not touch the ring: their wrapped offsets can fall
          // into a slot a TMA copy is still writing (racecheck), and they need no letters anyway
          if (ps >= -1) nlet = *reinterpret_cast<const uint32_t*>(stile + roff);
          roff = (roff + 4u) & (2 * TILE - 1);
        }
        // ---- the left neighbour's right edge of the same four rows (its previous step), all lanes converged;
        //      lane 0 takes the pass's left boundary instead: column 0 of the matrix, A(r,0) = -go - r*ge', in the
        //      first pass, else what lane 31 of the previous pass left in the scratch, re-based ---------------------
        uint32_t iH[4], iE[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          iH[k] = __shfl_up_sync(0xffffffffu, oH[k], 1);
          iE[k] = __shfl_up_sync(0xffffffffu, oE[k], 1);
        }
        if (firstp) {
          if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              iH[k] = colH - (uint32_t)k * gep2;
              iE[k] = iH[k] - goe2;
            }
          }
          colH -= 4u * gep2;
        } else if (lane == 0) {
          const uint32_t d2 = (uint32_t)((nbase.y - base_hi) * 65536 + (nbase.x - base_lo));
          iH[0] = nb0.x + d2; iE[0] = nb0.y + d2; iH[1] = nb0.z + d2; iE[1] = nb0.w + d2;
          iH[2] = nb1.x + d2; iE[2] = nb1.y + d2; iH[3] = nb1.z + d2; iE[3] = nb1.w + d2;
          if (s + 1 < nquads) {
            nb0 = *reinterpret_cast<const uint4*>(rp + kW16RecBytes);        // record s + 1
            nb1 = *reinterpret_cast<const uint4*>(rp + kW16RecBytes + 16);
            nbase = *reinterpret_cast<const int2*>(rp + kW16RecBytes + 32);
          }
        }
        const bool lane_on = ps >= 0 && (uint32_t)ps < nquads;
        if ((uint32_t)ps < nfull) {
          // ---- the common case: all four rows exist.  Rows A..D together, each one column behind the one
          //      above: FOUR independent dependency chains per lane (the two-row body of the tail path below
          //      has two), which is what keeps the half-rate DPX pipe fed from three warps per scheduler ------
          const uint4* prow[4];   // the row's first vector: one PRMT (letter) + one IMAD (byte offset) per row
#pragma unroll
          for (int r = 0; r < 4; ++r)
            prow[r] = reinterpret_cast<const uint4*>(myprofb + __byte_perm(let4, 0u, 0x4440u + (uint32_t)r) * (uint32_t)(PW * 4));
          uint32_t E[4], t[4], hl[4] = {0u, 0u, 0u, 0u};
          uint4 v[4][KW / 4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            E[r] = iE[r];
            v[r][0] = prow[r][0];
          }
          t[0] = hdiag + v[0][0].x;       // diagonal of row A's first cell: the left lane's H of the row above
          t[1] = iH[0] + v[1][0].x;
          t[2] = iH[1] + v[2][0].x;
          t[3] = iH[2] + v[3][0].x;
          hdiag = iH[3];
#pragma unroll
          for (int c = 0; c < KW + 3; ++c) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const int k = c - r;          // row r works on its column k in this round
              if (k >= 0 && k < KW) {
                // scores of columns 4g .. 4g+3: one LDS.128, issued one column before its first use
                if (((k + 2) & 3) == 0 && k + 2 < KW)
                  v[r][(k + 2) >> 2] = prow[r][((k + 2) >> 2) * 32];
                uint32_t tn = 0;
                if (k + 1 < KW) tn = H[k] + comp4(v[r][(k + 1) >> 2], (k + 1) & 3);   // H[k]: still the row above
                const uint32_t h = __vimax3_u16x2(t[r], E[r], F[k]);
                H[k] = h;
                const uint32_t hg = h - goe2;
                E[r] = __viaddmax_u16x2(E[r], nge, hg);
                F[k] = __viaddmax_u16x2(F[k], nge, hg);
                t[r] = tn;
                if (k == KW - 1) hl[r] = h;
              }
            }
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            oH[r] = hl[r];
            oE[r] = E[r];
          }
          if (lane == 31 && !lastp) {
            *reinterpret_cast<uint4*>(wp) = make_uint4(oH[0], oE[0], oH[1], oE[1]);   // lane 31: record ps = s - 31
            *reinterpret_cast<uint4*>(wp + 16) = make_uint4(oH[2], oE[2], oH[3], oE[3]);
            *reinterpret_cast<int2*>(wp + 32) = make_int2(base_lo, base_hi);
          }
          if ((uint32_t)ps == ps_last) {   // the subject's last row: H(m, n) of a query that ends in this block, made absolute
            if (pass == pass1) {
              const int32_t c1 = (int32_t)n1 - 1 - col0;
              if (c1 >= 0 && c1 < KW) {
#pragma unroll
                for (int c = 0; c < KW; ++c)
                  if (c == c1) res_lo = (int32_t)(H[c] & 0xffffu) + base_lo;
              }
            }
            if (lastp) {
              const int32_t c2 = (int32_t)n2 - 1 - col0;
              if (c2 >= 0 && c2 < KW) {
#pragma unroll
                for (int c = 0; c < KW; ++c)
                  if (c == c2) res_hi = (int32_t)(H[c] >> 16) + base_hi;
              }
            }
          }
        } else {
        // ---- the last step of a lane whose subject length is not a multiple of four: two rows, then one ----
        if (lane == 31 && !lastp && lane_on) *reinterpret_cast<int2*>(wp + 32) = make_int2(base_lo, base_hi);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint32_t iHa = iH[2 * q], iEa = iE[2 * q], iHb = iH[2 * q + 1], iEb = iE[2 * q + 1];
          const uint32_t ra = 4 * (uint32_t)ps + 2 * q + 1;
          if (lane_on && ra <= m) {
            const uint32_t* prow_a = myprof + ((let4 >> (16 * q)) & 0xffu) * PW;
            if (ra == m) {
              // ---- last row of an odd-length subject ----------------------------------------------
              uint32_t E = iEa;
              uint4 va[KW / 4];
#pragma unroll
              for (int g = 0; g < KW / 4; ++g) va[g] = *reinterpret_cast<const uint4*>(prow_a + g * 128);
              uint32_t t = hdiag + va[0].x;
              hdiag = iHa;
#pragma unroll
              for (int c = 0; c < KW; ++c) {
                uint32_t tn = 0;
                if (c + 1 < KW) tn = H[c] + comp4(va[(c + 1) >> 2], (c + 1) & 3);
                const uint32_t h = __vimax3_u16x2(t, E, F[c]);
                H[c] = h;
                const uint32_t hg = h - goe2;
                E = __viaddmax_u16x2(E, nge, hg);
                F[c] = __viaddmax_u16x2(F[c], nge, hg);
                t = tn;
              }
              oH[2 * q] = H[KW - 1];
              oE[2 * q] = E;
              if (lane == 31 && !lastp) *reinterpret_cast<uint2*>(wp + 16 * q) = make_uint2(oH[2 * q], oE[2 * q]);
            } else {
              // ---- rows ra (A) and ra+1 (B), B one column behind A ----------------------------------
              const uint32_t* prow_b = myprof + ((let4 >> (16 * q + 8)) & 0xffu) * PW;
              uint32_t Ea = iEa, Eb = iEb;
              // substitution scores: one LDS.128 per four columns and row, fetched two columns ahead of use
              // (one register group per four columns, each live only around its own columns: no copies)
              uint4 va[KW / 4], vb[KW / 4];
              va[0] = *reinterpret_cast<const uint4*>(prow_a);
              vb[0] = *reinterpret_cast<const uint4*>(prow_b);
              uint32_t ta = hdiag + va[0].x;
              uint32_t tb = iHa + vb[0].x;
              hdiag = iHb;
              uint32_t ha_last = 0;
#pragma unroll
              for (int c = 0; c <= KW; ++c) {
                // columns 4g .. 4g+3 of both rows are first needed at c = 4g - 1 (row A) / c = 4g (row B)
                if ((c & 3) == 1 && c + 3 < KW) {
                  va[(c + 3) >> 2] = *reinterpret_cast<const uint4*>(prow_a + ((c + 3) >> 2) * 128);
                  vb[(c + 3) >> 2] = *reinterpret_cast<const uint4*>(prow_b + ((c + 3) >> 2) * 128);
                }
                if (c < KW) {
                  uint32_t tn = 0;
                  if (c + 1 < KW) tn = H[c] + comp4(va[(c + 1) >> 2], (c + 1) & 3);
                  const uint32_t h = __vimax3_u16x2(ta, Ea, F[c]);
                  H[c] = h;
                  const uint32_t hg = h - goe2;
                  Ea = __viaddmax_u16x2(Ea, nge, hg);
                  F[c] = __viaddmax_u16x2(F[c], nge, hg);
                  ta = tn;
                  if (c == KW - 1) ha_last = h;
                }
                if (c >= 1) {
                  uint32_t tn = 0;
                  if (c < KW) tn = H[c - 1] + comp4(vb[c >> 2], c & 3);
                  const uint32_t h = __vimax3_u16x2(tb, Eb, F[c - 1]);
                  H[c - 1] = h;
                  const uint32_t hg = h - goe2;
                  Eb = __viaddmax_u16x2(Eb, nge, hg);
                  F[c - 1] = __viaddmax_u16x2(F[c - 1], nge, hg);
                  tb = tn;
                }
              }
              oH[2 * q] = ha_last; oE[2 * q] = Ea; oH[2 * q + 1] = H[KW - 1]; oE[2 * q + 1] = Eb;
              if (lane == 31 && !lastp) {
                *reinterpret_cast<uint4*>(wp + 16 * q) = make_uint4(oH[2 * q], oE[2 * q], oH[2 * q + 1], oE[2 * q + 1]);
              }
            }
            // ---- the subject's last row: H(m, n) of a query that ends in this block, made absolute ----
            if (ra + 1 >= m) {
              if (pass == pass1) {
                const int32_t c1 = (int32_t)n1 - 1 - col0;
                if (c1 >= 0 && c1 < KW) {
#pragma unroll
                  for (int c = 0; c < KW; ++c)
                    if (c == c1) res_lo = (int32_t)(H[c] & 0xffffu) + base_lo;
                }
              }
              if (lastp) {
                const int32_t c2 = (int32_t)n2 - 1 - col0;
                if (c2 >= 0 && c2 < KW) {
#pragma unroll
                  for (int c = 0; c < KW; ++c)
                    if (c == c2) res_hi = (int32_t)(H[c] >> 16) + base_hi;
                }
              }
            }
          }
        }
        }
        ++ps;
        rp += kW16RecBytes;
      }
      __syncwarp();
    }
    if (dead) break;
    // results live in the lanes that own column n1 / n2
    const int own1 = (int)(((n1 - 1) % PW) / KW), own2 = (int)(((n2 - 1) % PW) / KW);
    res_lo = __shfl_sync(0xffffffffu, res_lo, own1);
    res_hi = __shfl_sync(0xffffffffu, res_hi, own2);
    if (lane == 0) {
      p.out[tri_index(tk.x, tk.z, p.n_total)] = res_lo - p.delta * (int32_t)(m + n1);
      if (tk.y != tk.x) p.out[tri_index(tk.y, tk.z, p.n_total)] = res_hi - p.delta * (int32_t)(m + n2);
    }
  }
}

}  // namespace tsq
