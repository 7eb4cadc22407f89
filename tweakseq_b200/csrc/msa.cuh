// msa.cuh -- progressive multiple alignment along the guide tree: the step after SURVEY.md 8f-1 that
// turns the distance matrix + UPGMA tree into the equal-length gapped rows Project::readNewAlignment
// (tweakseq/Core/Project.cpp:908-1032) ingests, i.e. what the external aligner launched at
// tweakseq/UI/SeqEditMainWin.cpp:1654-1660 hands back.  Not the all-vs-all hot path: n-1 dependent
// profile alignments, bound by the per-diagonal barrier like traceback.cuh.
//
// Spec (build-defined; restated independently, with explicit row lists, by tsq_oracle_msa): merge t
// aligns the column profiles of clusters X = left[t] (columns along i) and Y = right[t] (along j) with
// the Gotoh recurrence of SURVEY 8c over COLUMNS, int64 sum-of-pairs scores
//     sub(i, j) = sum_a sum_b cntX[i][a] * cntY[j][b] * S(a, b)          (a residue facing a gap: 0)
// and a gap of k columns costing |X| |Y| (go + k ge), end gaps penalised.  Tie rules of traceback.cuh:
// H prefers the diagonal, then E (gap in X), then F (gap in Y); a gap run is opened rather than
// extended.  Two single sequences therefore align exactly as tsq_align_pair aligns them.
//
// One CTA per merge, all merges of one tree level in one launch.  A merge is four phases separated by
// CTA barriers; every phase is a function of (task, thread id, thread count) with no barrier inside, so
// tests/msa_emul.cpp can run the very same phase code on the CPU, thread by thread, against the oracle:
//   1. the column score in two factors: for the cluster with MORE sequences ("big") the letter scores
//      P[b][col] = sum_a cnt_big[col][a] S(a, b); for the other one ("small") per column the list of
//      letters present with their counts.  sub(i, j) = sum over that list of count * P[letter][col_big]:
//      one multiply-add per distinct letter of the small side's column -- exactly one when it is a
//      single sequence, the common case of a guide tree over related sequences (a caterpillar).
//   2. the sweep, in TILES of 4 x 4 cells: one thread per tile, one step (and one barrier) per anti-diagonal of
//      tiles; a thread fetches its tile's 16 column scores first, then runs the cells out of registers.  What
//      crosses a tile edge -- H and F of the last finished row per column, H and E of the last finished column per
//      row, one corner per tile row -- sits in SHARED memory when it fits (msa_diag_bytes; global scratch
//      otherwise); one direction byte per cell, diagonal-major (traceback.cuh).  Scores are int32 when the
//      merge's range bound allows (msa_fits_narrow: half the instructions of int64 on a 32-bit datapath)
//   3. thread 0 walks the path back from (Lx, Ly): per merged column its X column and Y column or -1.
//      The direction bytes sit in L2, so it fetches the next 16 along the current run (diagonal, or a
//      gap run) at once and consumes them from registers: one round trip per run piece, not per column
//   4. all threads: column maps old -> merged for X and Y, merged counts = cntX[px] + cntY[py]
// Profiles are letter-major (c[a * cap + col]) so that a diagonal's threads read consecutive words.
// No rows are materialised per merge: a leaf's residues reach their final columns by composing the
// column maps up the tree once, at the end (msa_rows_phase).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define TSQ_HD __host__ __device__ __forceinline__
#else
#define TSQ_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define TSQ_UNROLL _Pragma("unroll")
#define TSQ_NO_UNROLL _Pragma("unroll 1")
#else
#define TSQ_UNROLL
#define TSQ_NO_UNROLL
#endif

namespace tsq {

struct MsaConst {
  const int32_t* smat;   // nsym x nsym plain substitution scores
  uint32_t nsym;
  int32_t go, ge;
};

struct MsaResult {       // written by the walk-back of one merge
  long long score;       // H(Lx, Ly)
  uint32_t len;          // columns of the merged alignment
  uint32_t pad;
};

struct MsaTask {         // one merge: X = left child, Y = right child
  const uint32_t* cx;    // residue counts per column, letter-major: cx[a * capx + col]
  const uint32_t* cy;
  uint32_t* cn;          // counts of the merged alignment, capn >= Lx + Ly columns
  uint32_t capx, capy, capn;
  uint32_t Lx, Ly;       // columns of X and Y
  uint32_t nx, ny;       // sequences in X and Y
  uint32_t narrow;       // every H, E, F of this merge fits 30 bits (msa_fits_narrow): sweep in int32, else int64
  uint32_t pad;
  uint32_t* mapx;        // Lx entries: X column -> merged column
  uint32_t* mapy;        // Ly entries
  long long* diag;       // scratch: msa_diag_bytes, used when the sweep's edge arrays do not fit shared memory
  int32_t* pbig;         // scratch: nsym * max(Lx, Ly): P[b * Lbig + col] of the side with more sequences
  uint32_t* lst;         // scratch: nsym * max(Lx, Ly): lst[k * Lsmall + col] = letter << 24 | count
  uint32_t* lnz;         // scratch: max(Lx, Ly): distinct letters in the small side's column
  uint8_t* dir;          // scratch: (Lx + Ly + 1) * (min(Lx, Ly) + 1) direction bytes, diagonal-major
  int32_t* path;         // scratch: 2 * (Lx + Ly): (X column, Y column) per merged column, last column first
  MsaResult* res;
};

struct MsaLeaf {         // one input sequence
  const uint8_t* sym;    // encoded residues
  uint32_t len;
  uint32_t cap;          // column capacity of c (>= len)
  uint32_t* c;           // its profile: c[a * cap + col] = (sym[col] == a)
};

struct MsaRows {         // final rows: every residue follows the column maps up to the root
  const MsaLeaf* leaves;             // n
  const uint32_t* parent;            // 2n-1 node ids, 0xffffffff = root
  const uint32_t* const* nodemap;    // per node: its column map inside its parent's merge
  uint8_t* out;                      // n x ncols characters, pre-filled with '-'
  uint32_t ncols;
  uint32_t n;
  char letters[24];
};

// "minus infinity" of a score type: far below every real value, and NEG - GE still representable
template <typename T> struct MsaNeg;
template <> struct MsaNeg<long long> { static constexpr long long v = -(1LL << 60); };
template <> struct MsaNeg<int32_t> { static constexpr int32_t v = -(1 << 30); };

// The sweep works on TILES of kMsaTile x kMsaTile cells, one thread per tile, the tiles of one anti-diagonal of tiles
// per step (r02; one cell per thread and one CTA barrier per cell diagonal before: 4x the barriers, and a thread's
// whole dependent chain -- list length -> letter -> letter score -> max -- between every two of them).  What
// crosses a tile edge lives in five small arrays: H and F of the last finished row, per column (Ly + 1 each), H and
// E of the last finished column, per row (Lx + 1 each), and per tile row the corner H(i0 - 1, j0 - 1) of its next tile.
constexpr int kMsaTile = 4;
// Column scores do not come from the tile's own thread: while the tile threads run step st, the CTA's OTHER warps
// compute the scores of step st + 1 (one cell per worker at a time: a big x big merge has ~20 letters per column,
// and that serial chain inside the tile thread was 4/5 of its step) into one of two buffers of 16 scores per tile.
TSQ_HD size_t msa_tiles_per_step(uint32_t Lx, uint32_t Ly) {
  const size_t a = ((size_t)Lx + kMsaTile - 1) / kMsaTile, b = ((size_t)Ly + kMsaTile - 1) / kMsaTile;
  return a < b ? a : b;
}
TSQ_HD size_t msa_diag_bytes(uint32_t Lx, uint32_t Ly, bool narrow) {
  return (2 * ((size_t)Ly + 1) + 2 * ((size_t)Lx + 1) + ((size_t)Lx + kMsaTile - 1) / kMsaTile + 1 +
          2 * msa_tiles_per_step(Lx, Ly) * kMsaTile * kMsaTile + 4) *
         (narrow ? sizeof(int32_t) : sizeof(long long));
}

// Direction codes in shared memory: 4 bits per cell, row-major, a row padded to a multiple of 4 cells so that the four
// codes of a tile row are one aligned 16-bit word.  90 000 cells (a 300 x 300 merge) take 45 KB; the walk-back then
// reads them at shared-memory latency instead of one L2 round trip per run piece (once the sweep was tiled, the
// walk was two thirds of a merge).  Merges whose codes do not fit keep the diagonal-major bytes in global scratch.
TSQ_HD uint32_t msa_code_row_words(uint32_t Ly) { return (Ly + 3u) / 4u; }          // 16-bit words per matrix row
TSQ_HD size_t msa_code_bytes(uint32_t Lx, uint32_t Ly) { return (size_t)Lx * msa_code_row_words(Ly) * 2; }
// code of cell (i, j), 1-based: from the 4-bit array when there is one, else from the diagonal-major bytes
TSQ_HD uint32_t msa_read_code(const uint16_t* codes, const uint8_t* dir, int i, int j, int n, size_t ld) {
  if (codes) return ((uint32_t)codes[(size_t)(i - 1) * msa_code_row_words((uint32_t)n) + (uint32_t)((j - 1) >> 2)] >> (4 * ((j - 1) & 3))) & 15u;
  const int d = i + j;
  return dir[(size_t)d * ld + (size_t)(i - (d > n ? d - n : 0))];
}

// bytes of the three column-score tables of a merge (msa_prep_phase): letter scores of the big side, letter lists
// and list lengths of the small side.  When they fit shared memory next to the rolling diagonals the sweep reads
// them from there: the dependent chain list length -> letter -> letter score is then three shared-memory loads per
// diagonal instead of three L2 round trips (which were most of the ~1 000 clocks a diagonal took in r01).
TSQ_HD size_t msa_round16(size_t x) { return (x + 15) & ~(size_t)15; }
// rows of the small side's letter lists: a column of n sequences holds at most min(nsym, n) distinct letters
TSQ_HD uint32_t msa_list_rows(uint32_t nsym, uint32_t nsmall) { return nsmall < nsym ? nsmall : nsym; }
TSQ_HD size_t msa_table_bytes(uint32_t Lx, uint32_t Ly, uint32_t nsym, uint32_t nsmall) {
  const size_t mx = Lx > Ly ? Lx : Ly;   // either side may be the big one: size for the longer
  return msa_round16((size_t)nsym * mx * 4) + msa_round16((size_t)msa_list_rows(nsym, nsmall) * mx * 4) + msa_round16(mx * 4);
}

// Whether the whole DP of a merge stays inside +-2^29, so that the sweep may run in int32 with -2^30 as
// minus infinity: |H|, |E|, |F| <= |X||Y| (max|S| (Lx + Ly) + 2 go + (Lx + Ly + 2) ge).
TSQ_HD bool msa_fits_narrow(uint32_t nx, uint32_t ny, uint32_t Lx, uint32_t Ly, int32_t max_abs_s, int32_t go, int32_t ge) {
  const long long w = (long long)nx * (long long)ny;
  const long long per = (long long)max_abs_s * ((long long)Lx + Ly) + 2LL * go + ((long long)Lx + Ly + 2) * ge;
  return w < (1LL << 29) && per < (1LL << 29) && w * per < (1LL << 29);
}

TSQ_HD void msa_leaf_phase(const MsaLeaf& l, uint32_t nsym, int tid, int nt) {
  for (uint32_t col = (uint32_t)tid; col < l.len; col += (uint32_t)nt) {
    const uint32_t s = l.sym[col];
    for (uint32_t a = 0; a < nsym; ++a) l.c[(size_t)a * l.cap + col] = (a == s) ? 1u : 0u;
  }
}

// Which side is factored into letter scores: the one with more sequences (ties: X).
TSQ_HD bool msa_big_is_x(const MsaTask& t) { return t.nx >= t.ny; }

// phase 1.  One thread per column: its counts are fetched with independent loads (one L2 round trip, not
// one per letter), then the letter scores come out of registers.
constexpr int kMsaMaxSym = 24;
TSQ_HD void msa_prep_phase(const MsaTask& t, const MsaConst& k, int tid, int nt) {
  const bool bx = msa_big_is_x(t);
  const uint32_t* cb = bx ? t.cx : t.cy;
  const uint32_t* cs = bx ? t.cy : t.cx;
  const uint32_t capb = bx ? t.capx : t.capy, caps = bx ? t.capy : t.capx;
  const uint32_t Lb = bx ? t.Lx : t.Ly, Ls = bx ? t.Ly : t.Lx;
  const uint32_t nsym = k.nsym;   // <= kMsaMaxSym
  for (uint32_t col = (uint32_t)tid; col < Lb; col += (uint32_t)nt) {
    uint32_t cnt[kMsaMaxSym];
TSQ_UNROLL
    for (int a = 0; a < kMsaMaxSym; ++a) cnt[a] = (uint32_t)a < nsym ? cb[(size_t)a * capb + col] : 0u;
    for (uint32_t b = 0; b < nsym; ++b) {   // b: a letter of the small side
      int32_t s = 0;
TSQ_UNROLL
      for (int a = 0; a < kMsaMaxSym; ++a)
        // S(x letter, y letter): the big side's letter is the row index when the big side is X
        if ((uint32_t)a < nsym) s += (int32_t)cnt[a] * k.smat[bx ? (uint32_t)a * nsym + b : b * nsym + (uint32_t)a];
      t.pbig[(size_t)b * Lb + col] = s;
    }
  }
  for (uint32_t col = (uint32_t)tid; col < Ls; col += (uint32_t)nt) {
    uint32_t cnt[kMsaMaxSym];
TSQ_UNROLL
    for (int a = 0; a < kMsaMaxSym; ++a) cnt[a] = (uint32_t)a < nsym ? cs[(size_t)a * caps + col] : 0u;
    uint32_t nz = 0;
TSQ_UNROLL
    for (int a = 0; a < kMsaMaxSym; ++a)
      if (cnt[a]) { t.lst[(size_t)nz * Ls + col] = ((uint32_t)a << 24) | cnt[a]; ++nz; }
    t.lnz[col] = nz;
  }
}

// phase 2.  Loop invariants of one merge's sweep, set up once per thread.
template <typename T>
struct MsaSweep {
  T GO, GE, GOE;
  int m, n;              // Lx, Ly
  int ntr, ntc;          // rows / columns of tiles
  uint32_t ld;           // min(m, n) + 1: one diagonal of direction bytes
  uint32_t Lb, Ls;       // columns of the big / small side
  bool bx;               // the big side is X
  bool single;           // the small side is ONE sequence: at most one letter per column, scored inside the tile thread
  const uint32_t* lnz;
  const uint32_t* lst;
  const int32_t* pbig;
  uint8_t* dir;          // direction bytes, diagonal-major (global scratch); used when codes == nullptr
  uint16_t* codes;       // direction codes, 4 bits per cell, row-major (shared memory), or nullptr
  T* Hrow;               // [n + 1] H of the last finished row, per column
  T* Frow;               // [n + 1] F of it
  T* Hcol;               // [m + 1] H of the last finished column, per row
  T* Ecol;               // [m + 1] E of it
  T* corner;             // [ntr + 1] H(i0 - 1, j0 - 1) of the next tile of a tile row
  T* subbuf;             // 2 x [tiles per step][16] column scores: step st reads half st & 1
  uint32_t subhalf;      // scores in one half
};

struct MsaTables {       // where the sweep reads the column-score tables from (the task's scratch, or shared memory)
  const int32_t* pbig;
  const uint32_t* lst;
  const uint32_t* lnz;
};

template <typename T>
TSQ_HD MsaSweep<T> msa_sweep_init(const MsaTask& t, const MsaConst& k, void* edge, const MsaTables* tab = nullptr, uint16_t* codes = nullptr) {
  MsaSweep<T> s;
  const long long w = (long long)t.nx * (long long)t.ny;
  s.GO = (T)(w * k.go); s.GE = (T)(w * k.ge); s.GOE = (T)(w * k.go + w * k.ge);
  s.m = (int)t.Lx; s.n = (int)t.Ly;
  s.ntr = (s.m + kMsaTile - 1) / kMsaTile; s.ntc = (s.n + kMsaTile - 1) / kMsaTile;
  s.ld = (t.Lx < t.Ly ? t.Lx : t.Ly) + 1;
  s.bx = msa_big_is_x(t);
  s.Lb = s.bx ? t.Lx : t.Ly; s.Ls = s.bx ? t.Ly : t.Lx;
  s.single = (s.bx ? t.ny : t.nx) == 1;
  s.lnz = tab ? tab->lnz : t.lnz;
  s.lst = tab ? tab->lst : t.lst;
  s.pbig = tab ? tab->pbig : t.pbig;
  s.dir = t.dir;
  s.codes = codes;
  s.Hrow = (T*)edge;
  s.Frow = s.Hrow + (s.n + 1);
  s.Hcol = s.Frow + (s.n + 1);
  s.Ecol = s.Hcol + (s.m + 1);
  s.corner = s.Ecol + (s.m + 1);
  s.subbuf = s.corner + (s.ntr + 1);
  s.subbuf += (4 - ((s.subbuf - (T*)edge) & 3)) & 3;   // a tile's 16 scores start on a 4-element boundary
  s.subhalf = (uint32_t)(msa_tiles_per_step(t.Lx, t.Ly) * kMsaTile * kMsaTile);
  return s;
}

// steps of the sweep: one per anti-diagonal of tiles
template <typename T>
TSQ_HD int msa_sweep_steps(const MsaSweep<T>& s) { return (s.ntr > 0 && s.ntc > 0) ? s.ntr + s.ntc - 1 : 0; }

// H(Lx, Ly) after the sweep
template <typename T>
TSQ_HD long long msa_final_score(const MsaSweep<T>& s) {
  return (long long)(s.m > 0 ? s.Hcol[s.m] : s.Hrow[s.n]);
}

// Row 0 and column 0 of the matrix, as the edges the first tiles read: H(0, j) = -GO - j GE, F(0, j) = -inf,
// H(i, 0) = -GO - i GE, E(i, 0) = -inf, H(0, 0) = 0.
template <typename T>
TSQ_HD void msa_edge_phase(const MsaSweep<T>& s, int tid, int nt) {
  constexpr T NEG = MsaNeg<T>::v;
  for (int j = tid; j <= s.n; j += nt) {
    s.Hrow[j] = j == 0 ? (T)0 : (T)(-s.GO - (T)j * s.GE);
    s.Frow[j] = NEG;
  }
  for (int i = tid; i <= s.m; i += nt) {
    s.Hcol[i] = i == 0 ? (T)0 : (T)(-s.GO - (T)i * s.GE);
    s.Ecol[i] = NEG;
  }
  for (int I = tid; I <= s.ntr; I += nt) s.corner[I] = I == 0 ? (T)0 : (T)(-s.GO - (T)(I * kMsaTile) * s.GE);   // H(I * tile, 0)
}

// The cells of one tile, row by row out of registers.  FULL: all kMsaTile x kMsaTile cells exist -- no guards, one
// straight-line block in which the scheduler interleaves the independent cells of the tile's own anti-diagonals (a
// thread's dependent chain is 7 cells long, not 16).
// CODES4: the direction codes go to the 4-bit array in shared memory, one 16-bit store per tile row; else one byte per
// cell to the diagonal-major array in global scratch.  A template argument, not a test per cell: a branch inside the
// block would end the scheduler's window at every cell and serialise the tile.
template <typename T, bool FULL, bool CODES4>
TSQ_HD void msa_tile_cells(const MsaSweep<T>& s, int I, int i0, int j0, int ih, int jw, const T (&sub)[kMsaTile][kMsaTile]) {
  constexpr int TB = kMsaTile;
  // ---- edges in ----
  T th[TB], tf[TB];      // H and F of the row above, per tile column
  T hl[TB], el[TB];      // H and E of the column to the left, per tile row
TSQ_UNROLL
  for (int c = 0; c < TB; ++c) {
    th[c] = (FULL || c < jw) ? s.Hrow[j0 + c] : (T)0;
    tf[c] = (FULL || c < jw) ? s.Frow[j0 + c] : (T)0;
  }
TSQ_UNROLL
  for (int r = 0; r < TB; ++r) {
    hl[r] = (FULL || r < ih) ? s.Hcol[i0 + r] : (T)0;
    el[r] = (FULL || r < ih) ? s.Ecol[i0 + r] : (T)0;
  }
  T dcorner = s.corner[I];                               // H(i0 - 1, j0 - 1)
  s.corner[I] = FULL ? th[TB - 1] : s.Hrow[j0 + jw - 1];  // H(i0 - 1, j0 + jw - 1): the corner of tile (I, J + 1)
  // direction bytes: cell (i, j) at dir[d * ld + i - max(0, d - n)], d = i + j
  const size_t ld = s.ld;
  // ---- cells ----
TSQ_UNROLL
  for (int r = 0; r < TB; ++r) {
    if (FULL || r < ih) {
      const int i = i0 + r;
      T h = hl[r], e = el[r];              // H(i, j0 - 1), E(i, j0 - 1)
      T dg0 = dcorner;
      dcorner = h;                         // H(i, j0 - 1): the diagonal of row i + 1's first cell
      uint32_t rowcodes = 0;               // the row's codes, 4 bits each
TSQ_UNROLL
      for (int c = 0; c < TB; ++c) {
        if (FULL || c < jw) {
          const int j = j0 + c;
          const T e_ext = e - s.GE, e_open = h - s.GOE;
          const T f_ext = tf[c] - s.GE, f_open = th[c] - s.GOE;
          const T dg = dg0 + sub[r][c];
          const bool eo = e_open >= e_ext, fo = f_open >= f_ext;
          const T E = eo ? e_open : e_ext;
          const T F = fo ? f_open : f_ext;
          T H = dg;
          if (E > H) H = E;
          if (F > H) H = F;
          const uint32_t code = (H == dg ? 0u : (H == E ? 1u : 2u)) | (eo ? 4u : 0u) | (fo ? 8u : 0u);
          if (CODES4) {
            rowcodes |= code << (4 * c);
          } else {
            const int d = i + j;
            s.dir[(size_t)d * ld + (size_t)(i - (d > s.n ? d - s.n : 0))] = (uint8_t)code;
          }
          dg0 = th[c];                     // H(i - 1, j): the diagonal of (i, j + 1)
          th[c] = H; tf[c] = F;
          h = H; e = E;
        }
      }
      s.Hcol[i] = h;                       // H(i, j0 + jw - 1), E of it
      s.Ecol[i] = e;
      if (CODES4) s.codes[(size_t)(i - 1) * msa_code_row_words((uint32_t)s.n) + (uint32_t)((j0 - 1) >> 2)] = (uint16_t)rowcodes;
    }
  }
  // ---- edges out ----
TSQ_UNROLL
  for (int c = 0; c < TB; ++c)
    if (FULL || c < jw) { s.Hrow[j0 + c] = th[c]; s.Frow[j0 + c] = tf[c]; }
}

// The column scores of step st, sub(i, j) = sum over the small side's letters of count * letter score of the big
// side: worker w of nw takes the cells w, w + nw, ... of the step's tiles (16 per tile, row-major; cells off the
// matrix stay unwritten and unread).
template <typename T>
TSQ_HD void msa_sub_phase(const MsaSweep<T>& s, int st, int w, int nw) {
  constexpr int TB = kMsaTile;
  if (s.single || st >= msa_sweep_steps<T>(s)) return;   // (a single sequence's scores come out of the tile thread)
  const int Ilo = st >= s.ntc ? st - s.ntc + 1 : 0;
  const int Ihi = st < s.ntr ? st : s.ntr - 1;
  T* const out = s.subbuf + (size_t)(st & 1) * s.subhalf;
  const int ncell = (Ihi - Ilo + 1) * TB * TB;
  for (int f = w; f < ncell; f += nw) {
    const int I = Ilo + (f >> 4), r = (f >> 2) & 3, c = f & 3;
    const int i = I * TB + 1 + r, j = (st - I) * TB + 1 + c;
    if (i > s.m || j > s.n) continue;
    const uint32_t colb = (uint32_t)(s.bx ? i - 1 : j - 1), cols = (uint32_t)(s.bx ? j - 1 : i - 1);
    const uint32_t nz = s.lnz[cols];
    const uint32_t* le = s.lst + cols;
    const int32_t* pb = s.pbig + colb;
    T sub = 0;
    TSQ_NO_UNROLL
    for (uint32_t q = 0; q < nz; ++q, le += s.Ls) {
      const uint32_t e = *le;
      sub += (T)(e & 0xffffffu) * (T)pb[(e >> 24) * s.Lb];
    }
    out[f] = sub;
  }
}

// One step: the tiles (I, J) with I + J == st; thread tid takes tile row Ilo + tid (+ nt, ...).  Rows i0 .. i0 + ih - 1,
// columns j0 .. j0 + jw - 1 of the matrix (1-based cells).  The tile's column scores were left by msa_sub_phase.
template <typename T>
TSQ_HD void msa_tile_phase(const MsaSweep<T>& s, int st, int tid, int nt) {
  constexpr int TB = kMsaTile;
  const int Ilo = st >= s.ntc ? st - s.ntc + 1 : 0;
  const int Ihi = st < s.ntr ? st : s.ntr - 1;
  const T* const in = s.subbuf + (size_t)(st & 1) * s.subhalf;
  for (int I = Ilo + tid; I <= Ihi; I += nt) {
    const int J = st - I;
    const int i0 = I * TB + 1, j0 = J * TB + 1;
    const int ih = s.m - i0 + 1 < TB ? s.m - i0 + 1 : TB;
    const int jw = s.n - j0 + 1 < TB ? s.n - j0 + 1 : TB;
    const bool full = ih == TB && jw == TB;
    T sub[TB][TB];
    if (s.single) {
      // one sequence on the small side: one letter (or none) per small-side column, 4 loads + 16 loads + 16 products,
      // straight-line.  The small side runs along j when the big side is X, along i otherwise; a0 / b0: first
      // small-side / big-side column of the tile, na / nb: how many of each exist
      const uint32_t a0 = (uint32_t)(s.bx ? j0 : i0) - 1, b0 = (uint32_t)(s.bx ? i0 : j0) - 1;
      const int na = s.bx ? jw : ih, nb = s.bx ? ih : jw;
      T sa[TB][TB];      // [small-side column][big-side column]
      uint32_t e[TB];
TSQ_UNROLL
      for (int a = 0; a < TB; ++a) e[a] = ((full || a < na) && s.lnz[a0 + a] != 0u) ? s.lst[a0 + a] : 0u;   // count 0: scores 0
TSQ_UNROLL
      for (int a = 0; a < TB; ++a) {
        const int32_t* pb = s.pbig + (size_t)(e[a] >> 24) * s.Lb + b0;
        const T cnt = (T)(e[a] & 0xffffffu);
TSQ_UNROLL
        for (int b = 0; b < TB; ++b) sa[a][b] = (full || b < nb) ? cnt * (T)pb[b] : (T)0;
      }
TSQ_UNROLL
      for (int r = 0; r < TB; ++r)
TSQ_UNROLL
        for (int c = 0; c < TB; ++c) sub[r][c] = s.bx ? sa[c][r] : sa[r][c];
    } else {
      const T* const mine = in + (size_t)(I - Ilo) * (TB * TB);
TSQ_UNROLL
      for (int r = 0; r < TB; ++r)
TSQ_UNROLL
        for (int c = 0; c < TB; ++c) sub[r][c] = (full || (r < ih && c < jw)) ? mine[r * TB + c] : (T)0;
    }
    if (s.codes) {
      if (full) msa_tile_cells<T, true, true>(s, I, i0, j0, ih, jw, sub);
      else msa_tile_cells<T, false, true>(s, I, i0, j0, ih, jw, sub);
    } else {
      if (full) msa_tile_cells<T, true, false>(s, I, i0, j0, ih, jw, sub);
      else msa_tile_cells<T, false, false>(s, I, i0, j0, ih, jw, sub);
    }
  }
}

// How a CTA (or the emulation) of nt threads splits into tile threads and column-score workers: the first tw threads
// run the tiles; the rest, if there is at least a warp of them, compute the next step's column scores meanwhile.
// Otherwise everybody does both, one after the other.
TSQ_HD int msa_tile_threads(uint32_t Lx, uint32_t Ly, int nt) {
  const int tw = (int)((msa_tiles_per_step(Lx, Ly) + 31) & ~(size_t)31);
  return nt - tw >= 32 ? tw : nt;
}

// phase 3, one thread.  score = H(Lx, Ly) (msa_final_score).
TSQ_HD void msa_walk_phase(const MsaTask& t, long long score, const uint16_t* codes4 = nullptr) {
  constexpr int B = 16;   // direction bytes fetched per round trip
  const int m = (int)t.Lx, n = (int)t.Ly;
  const size_t ld = (size_t)(m < n ? m : n) + 1;
  int i = m, j = n, state = 0;
  uint32_t k = 0;
  while (i > 0 && j > 0) {
    // the next B cells along the current run: down the diagonal, or along the gap run
    const int di = state == 1 ? 0 : 1, dj = state == 2 ? 0 : 1;
    uint32_t codes[B];
TSQ_UNROLL
    for (int q = 0; q < B; ++q) {
      const int ii = i - q * di, jj = j - q * dj;
      uint32_t c = 0;
      if (ii > 0 && jj > 0) c = msa_read_code(codes4, t.dir, ii, jj, n, ld);
      codes[q] = c;
    }
    bool turned = false;
TSQ_UNROLL
    for (int q = 0; q < B; ++q) {
      if (turned || i == 0 || j == 0) continue;
      const uint32_t code = codes[q];
      if (state == 0) {
        const uint32_t src = code & 3u;
        if (src == 0) { t.path[2 * k] = i - 1; t.path[2 * k + 1] = j - 1; --i; --j; ++k; }
        else { state = (int)src; turned = true; }       // same cell again, as the start of a gap run
      } else if (state == 1) {            // gap in X: the merged column takes Y's column only
        t.path[2 * k] = -1; t.path[2 * k + 1] = j - 1; ++k;
        if (code & 4u) { state = 0; turned = true; }
        --j;
      } else {                            // gap in Y
        t.path[2 * k] = i - 1; t.path[2 * k + 1] = -1; ++k;
        if (code & 8u) { state = 0; turned = true; }
        --i;
      }
    }
  }
  while (j > 0) { t.path[2 * k] = -1; t.path[2 * k + 1] = j - 1; --j; ++k; }
  while (i > 0) { t.path[2 * k] = i - 1; t.path[2 * k + 1] = -1; --i; ++k; }
  t.res->len = k;
  t.res->pad = 0;
  t.res->score = score;
}

// phase 4
TSQ_HD void msa_build_phase(const MsaTask& t, const MsaConst& k, int tid, int nt) {
  const uint32_t len = t.res->len;
  const uint32_t nsym = k.nsym;
  for (uint32_t c = (uint32_t)tid; c < len; c += (uint32_t)nt) {
    const uint32_t s = len - 1 - c;
    const int32_t xi = t.path[2 * s], yj = t.path[2 * s + 1];
    if (xi >= 0) t.mapx[xi] = c;
    if (yj >= 0) t.mapy[yj] = c;
    uint32_t v[kMsaMaxSym];
TSQ_UNROLL
    for (int a = 0; a < kMsaMaxSym; ++a) {   // all loads before the first store: they overlap
      uint32_t sum = 0;
      if ((uint32_t)a < nsym) {
        if (xi >= 0) sum += t.cx[(size_t)a * t.capx + (uint32_t)xi];
        if (yj >= 0) sum += t.cy[(size_t)a * t.capy + (uint32_t)yj];
      }
      v[a] = sum;
    }
TSQ_UNROLL
    for (int a = 0; a < kMsaMaxSym; ++a)
      if ((uint32_t)a < nsym) t.cn[(size_t)a * t.capn + c] = v[a];
  }
}

// final rows: leaf r, its residues tid, tid + nt, ...
TSQ_HD void msa_rows_phase(const MsaRows& p, uint32_t r, int tid, int nt) {
  const MsaLeaf& l = p.leaves[r];
  for (uint32_t q = (uint32_t)tid; q < l.len; q += (uint32_t)nt) {
    uint32_t col = q, node = r;
    for (;;) {
      const uint32_t up = p.parent[node];
      if (up == 0xffffffffu) break;
      col = p.nodemap[node][col];
      node = up;
    }
    p.out[(size_t)r * p.ncols + col] = (uint8_t)p.letters[l.sym[q]];
  }
}

// phase 3 by ONE WARP (the device's walk-back).  The serial walk above is a chain of round trips consumed one code at a
// time by one thread: a third of a merge's time once the sweep was tiled.  Here lane q fetches the direction codes of
// the cells q and q + 32 steps ahead along the current run (down the diagonal, or along the gap run), a ballot finds
// where the run ends, and all lanes before that point write their path entries at once.  Same state machine, same
// path, entry for entry.  The pieces below are per-lane or warp-uniform functions of plain values, so that the CPU
// emulation (tests/msa_emul.cpp) runs the same code with the ballots done by a loop over 32 lanes and compares the
// path with the serial walk's; msa_walk_warp is the device's glue around them.
struct MsaWalk {
  int i, j, state;       // current cell, 0 = on the diagonal run, 1 = gap in X (moving along j), 2 = gap in Y
  uint32_t k;            // path entries written
};

// lane q: codes of the cells q and q + 32 steps ahead along the current run (valid = the cell is on the matrix)
TSQ_HD void msa_walk_fetch(const MsaTask& t, const uint16_t* codes4, const MsaWalk& w, int lane, uint32_t (&code)[2], bool (&valid)[2]) {
  const int n = (int)t.Ly;
  const size_t ld = (size_t)(t.Lx < t.Ly ? t.Lx : t.Ly) + 1;
  const int di = w.state == 1 ? 0 : 1, dj = w.state == 2 ? 0 : 1;
  for (int h = 0; h < 2; ++h) {
    const int q = lane + 32 * h;
    const int ii = w.i - q * di, jj = w.j - q * dj;
    valid[h] = ii > 0 && jj > 0;
    code[h] = valid[h] ? msa_read_code(codes4, t.dir, ii, jj, n, ld) : 0u;
  }
}

// a cell on the matrix ends the run by a TURN: (state 0) it opens a gap run; (gap states) its flag says the run was
// opened here -- that cell still belongs to the run
TSQ_HD bool msa_walk_turn(int state, uint32_t code) {
  return state == 0 ? (code & 3u) != 0u : state == 1 ? (code & 4u) != 0u : (code & 8u) != 0u;
}

// from the two ballots of a half (stops: off the matrix or a turn; turns): the first stopping lane s (32: none), whether
// it is a turn, and how many path entries the half yields
TSQ_HD int msa_walk_count(int state, uint32_t stops, uint32_t turns, int* s_out, bool* s_turn_out) {
  int s = 32;
  if (stops) {
    s = 0;
    while (!((stops >> s) & 1u)) ++s;
  }
  const bool s_turn = s < 32 && ((turns >> s) & 1u);
  *s_out = s;
  *s_turn_out = s_turn;
  return (state != 0 && s_turn) ? s + 1 : s;
}

// lane < count writes the path entry of its cell of the half (i, j: the half's first cell)
TSQ_HD void msa_walk_store(const MsaTask& t, const MsaWalk& w, int lane, int count) {
  if (lane >= count) return;
  const int di = w.state == 1 ? 0 : 1, dj = w.state == 2 ? 0 : 1;
  const int ii = w.i - lane * di, jj = w.j - lane * dj;
  t.path[2 * (w.k + (uint32_t)lane)] = w.state == 1 ? -1 : ii - 1;
  t.path[2 * (w.k + (uint32_t)lane) + 1] = w.state == 2 ? -1 : jj - 1;
}

// after a half: move on; returns whether the run went through all 32 cells (the second half may follow)
TSQ_HD bool msa_walk_advance(MsaWalk& w, int s, bool s_turn, int count, uint32_t code_at_s) {
  const int di = w.state == 1 ? 0 : 1, dj = w.state == 2 ? 0 : 1;
  w.k += (uint32_t)count;
  w.i -= count * di;
  w.j -= count * dj;
  if (s >= 32) return true;
  if (s_turn) w.state = w.state == 0 ? (int)(code_at_s & 3u) : 0;   // (off the matrix: i or j is 0 now, the walk's loop ends)
  return false;
}

// the rest of the longer side faces gaps: lane's share of the two tails; then the totals
TSQ_HD void msa_walk_tails(const MsaTask& t, const MsaWalk& w, int lane) {
  for (int q = lane; q < w.j; q += 32) { t.path[2 * (w.k + (uint32_t)q)] = -1; t.path[2 * (w.k + (uint32_t)q) + 1] = w.j - 1 - q; }
  const uint32_t k2 = w.k + (uint32_t)(w.j > 0 ? w.j : 0);
  for (int q = lane; q < w.i; q += 32) { t.path[2 * (k2 + (uint32_t)q)] = w.i - 1 - q; t.path[2 * (k2 + (uint32_t)q) + 1] = -1; }
}
TSQ_HD void msa_walk_finish(const MsaTask& t, const MsaWalk& w, long long score) {
  t.res->len = w.k + (uint32_t)(w.j > 0 ? w.j : 0) + (uint32_t)(w.i > 0 ? w.i : 0);
  t.res->pad = 0;
  t.res->score = score;
}

#ifdef TSQ_DEVICE_IMPL
__device__ __forceinline__ void msa_walk_warp(const MsaTask& t, long long score, int lane, const uint16_t* codes4) {
  MsaWalk w{(int)t.Lx, (int)t.Ly, 0, 0u};
  while (w.i > 0 && w.j > 0) {
    uint32_t code[2];
    bool valid[2];
    msa_walk_fetch(t, codes4, w, lane, code, valid);
    bool more = true;   // warp-uniform
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (!more) break;
      const bool turn = valid[h] && msa_walk_turn(w.state, code[h]);
      const uint32_t stops = __ballot_sync(0xffffffffu, !valid[h] || turn);
      const uint32_t turns = __ballot_sync(0xffffffffu, turn);
      int s;
      bool s_turn;
      const int count = msa_walk_count(w.state, stops, turns, &s, &s_turn);
      msa_walk_store(t, w, lane, count);
      const uint32_t code_at_s = __shfl_sync(0xffffffffu, code[h], s & 31);
      more = msa_walk_advance(w, s, s_turn, count, code_at_s);
    }
  }
  msa_walk_tails(t, w, lane);
  if (lane == 0) msa_walk_finish(t, w, score);
}

__global__ void __launch_bounds__(128) msa_leaf_kernel(const MsaLeaf* leaves, uint32_t n, uint32_t nsym) {
  for (uint32_t r = blockIdx.x; r < n; r += gridDim.x) msa_leaf_phase(leaves[r], nsym, (int)threadIdx.x, (int)blockDim.x);
}

// The sweep of one merge with its edge arrays at `edge`.  Called once with the shared-memory array and once with
// the global scratch, so that each copy of the loop knows its address space (LDS/STS with 32-bit offsets instead
// of generic 64-bit addressing).
template <typename T>
__device__ __forceinline__ long long msa_sweep_cta(const MsaTask& t, const MsaConst& k, T* edge, const MsaTables* tab, uint16_t* codes, int tid, int nt) {
  const MsaSweep<T> sw = msa_sweep_init<T>(t, k, edge, tab, codes);
  const int tw = msa_tile_threads(t.Lx, t.Ly, nt);
  const bool split = tw < nt;                 // warps tw/32 .. are column-score workers
  msa_edge_phase<T>(sw, tid, nt);
  msa_sub_phase<T>(sw, 0, tid, nt);
  __syncthreads();
  const int steps = msa_sweep_steps<T>(sw);
  for (int st = 0; st < steps; ++st) {
    if (tid < tw) msa_tile_phase<T>(sw, st, tid, tw);
    else msa_sub_phase<T>(sw, st + 1, tid - tw, nt - tw);
    __syncthreads();   // the tiles of step st and the column scores of st + 1 complete and visible to the whole CTA
    if (!split) {
      msa_sub_phase<T>(sw, st + 1, tid, nt);
      __syncthreads();
    }
  }
  return msa_final_score<T>(sw);
}

// One merge by one CTA, in score type T.  smem_bytes: dynamic shared memory of the launch.  A merge whose edge
// arrays fit uses it for them (its global scratch otherwise), and one whose column-score tables fit behind them
// copies them there after the prep phase.
template <typename T>
__device__ __forceinline__ void msa_merge_cta(const MsaTask& t, const MsaConst& k, uint32_t smem_bytes, T* smem) {
  const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
  msa_prep_phase(t, k, tid, nt);
  __syncthreads();
  long long score;
  uint16_t* codes = nullptr;   // direction codes in shared memory, where they fit too
  const size_t db = msa_round16(msa_diag_bytes(t.Lx, t.Ly, sizeof(T) == 4));
  if (db <= (size_t)smem_bytes) {
    const uint32_t Lb = msa_big_is_x(t) ? t.Lx : t.Ly, Ls = msa_big_is_x(t) ? t.Ly : t.Lx;
    const uint32_t nsmall = msa_big_is_x(t) ? t.ny : t.nx;
    const size_t pb = msa_round16((size_t)k.nsym * Lb * 4), lb = msa_round16((size_t)msa_list_rows(k.nsym, nsmall) * Ls * 4), nb = msa_round16((size_t)Ls * 4);
    if (db + pb + lb + nb <= (size_t)smem_bytes) {
      char* base = reinterpret_cast<char*>(smem) + db;
      int32_t* s_pbig = reinterpret_cast<int32_t*>(base);
      uint32_t* s_lst = reinterpret_cast<uint32_t*>(base + pb);
      uint32_t* s_lnz = reinterpret_cast<uint32_t*>(base + pb + lb);
      if (db + pb + lb + nb + msa_round16(msa_code_bytes(t.Lx, t.Ly)) <= (size_t)smem_bytes)
        codes = reinterpret_cast<uint16_t*>(base + pb + lb + nb);
      for (uint32_t q = (uint32_t)tid; q < k.nsym * Lb; q += (uint32_t)nt) s_pbig[q] = t.pbig[q];
      for (uint32_t q = (uint32_t)tid; q < Ls; q += (uint32_t)nt) {
        const uint32_t nz = t.lnz[q];
        s_lnz[q] = nz;
        for (uint32_t r = 0; r < nz; ++r) s_lst[(size_t)r * Ls + q] = t.lst[(size_t)r * Ls + q];   // only the entries in use
      }
      __syncthreads();
      const MsaTables tab{s_pbig, s_lst, s_lnz};
      score = msa_sweep_cta<T>(t, k, smem, &tab, codes, tid, nt);
    } else {
      score = msa_sweep_cta<T>(t, k, smem, nullptr, nullptr, tid, nt);
    }
  } else {
    score = msa_sweep_cta<T>(t, k, reinterpret_cast<T*>(t.diag), nullptr, nullptr, tid, nt);
  }
  if (tid < 32) msa_walk_warp(t, score, tid, codes);
  __syncthreads();
  msa_build_phase(t, k, tid, nt);
}

// MAXT: the largest block size the variant is compiled for (registers per thread follow from it).
template <int MAXT>
__global__ void __launch_bounds__(MAXT) msa_merge_kernel(const MsaTask* tasks, const MsaConst k, uint32_t smem_bytes) {
  extern __shared__ long long msa_shared_diag[];
  const MsaTask t = tasks[blockIdx.x];
  if (t.narrow) msa_merge_cta<int32_t>(t, k, smem_bytes, reinterpret_cast<int32_t*>(msa_shared_diag));
  else msa_merge_cta<long long>(t, k, smem_bytes, msa_shared_diag);
}

__global__ void __launch_bounds__(256) msa_rows_kernel(const __grid_constant__ MsaRows p) {
  for (uint32_t r = blockIdx.x; r < p.n; r += gridDim.x) msa_rows_phase(p, r, (int)threadIdx.x, (int)blockDim.x);
}
#endif  // TSQ_DEVICE_IMPL

}  // namespace tsq
