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
//   2. anti-diagonal sweep, one barrier per diagonal: H (3 rolling diagonals), E, F (2 each) indexed by
//      i, in SHARED memory when 7 (Lx + 1) scores fit (an L2 round trip per diagonal otherwise); one
//      direction byte per cell, diagonal-major (traceback.cuh).  Scores are int32 when the merge's
//      range bound allows (msa_fits_narrow: half the instructions of int64 on a 32-bit datapath)
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
  long long* diag;       // scratch: 7 * (Lx + 1) scores, used when the rolling diagonals do not fit shared memory
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

// bytes of the 7 rolling diagonals (3 H, 2 E, 2 F) of a merge with Lx columns along i
TSQ_HD size_t msa_diag_bytes(uint32_t Lx, bool narrow) { return 7 * ((size_t)Lx + 1) * (narrow ? sizeof(int32_t) : sizeof(long long)); }

// bytes of the three column-score tables of a merge (msa_prep_phase): letter scores of the big side, letter lists
// and list lengths of the small side.  When they fit shared memory next to the rolling diagonals the sweep reads
// them from there: the dependent chain list length -> letter -> letter score is then three shared-memory loads per
// diagonal instead of three L2 round trips (which were most of the ~1 000 clocks a diagonal took in r01).
TSQ_HD size_t msa_round16(size_t x) { return (x + 15) & ~(size_t)15; }
TSQ_HD size_t msa_table_bytes(uint32_t Lx, uint32_t Ly, uint32_t nsym) {
  const size_t mx = Lx > Ly ? Lx : Ly;   // either side may be the big one: size for the longer
  return msa_round16((size_t)nsym * mx * 4) + msa_round16((size_t)nsym * mx * 4) + msa_round16(mx * 4);
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
  uint32_t stride;       // m + 1: one rolling diagonal
  uint32_t ld;           // min(m, n) + 1: one diagonal of direction bytes
  uint32_t Lb, Ls;       // columns of the big / small side
  bool bx;               // the big side is X
  const uint32_t* lnz;
  const uint32_t* lst;
  const int32_t* pbig;
  uint8_t* dir;
  T* diag;               // 3 H, 2 E, 2 F diagonals
};

struct MsaTables {       // where the sweep reads the column-score tables from (the task's scratch, or shared memory)
  const int32_t* pbig;
  const uint32_t* lst;
  const uint32_t* lnz;
};

template <typename T>
TSQ_HD MsaSweep<T> msa_sweep_init(const MsaTask& t, const MsaConst& k, void* diag, const MsaTables* tab = nullptr) {
  MsaSweep<T> s;
  const long long w = (long long)t.nx * (long long)t.ny;
  s.GO = (T)(w * k.go); s.GE = (T)(w * k.ge); s.GOE = (T)(w * k.go + w * k.ge);
  s.m = (int)t.Lx; s.n = (int)t.Ly;
  s.stride = t.Lx + 1;
  s.ld = (t.Lx < t.Ly ? t.Lx : t.Ly) + 1;
  s.bx = msa_big_is_x(t);
  s.Lb = s.bx ? t.Lx : t.Ly; s.Ls = s.bx ? t.Ly : t.Lx;
  s.lnz = tab ? tab->lnz : t.lnz;
  s.lst = tab ? tab->lst : t.lst;
  s.pbig = tab ? tab->pbig : t.pbig;
  s.dir = t.dir;
  s.diag = (T*)diag;
  return s;
}

// H(Lx, Ly) after the sweep: diagonal Lx + Ly sits in H buffer (Lx + Ly) % 3
template <typename T>
TSQ_HD long long msa_final_score(const MsaSweep<T>& s) {
  return (long long)s.diag[(size_t)((s.m + s.n) % 3) * s.stride + (size_t)s.m];
}

// One diagonal d in [0, Lx + Ly]; hc = d % 3 (the caller counts it along: no division per diagonal).
template <typename T>
TSQ_HD void msa_diag_phase(const MsaSweep<T>& s, int d, int hc, int tid, int nt) {
  constexpr T NEG = MsaNeg<T>::v;
  const int h1 = hc ? hc - 1 : 2, h2 = h1 ? h1 - 1 : 2;   // buffers of diagonals d-1, d-2
  const int par = d & 1;
  T* const Hc = s.diag + (size_t)hc * s.stride;
  const T* const Hp1 = s.diag + (size_t)h1 * s.stride;
  const T* const Hp2 = s.diag + (size_t)h2 * s.stride;
  T* const Ec = s.diag + (size_t)(3 + par) * s.stride;
  const T* const Ep1 = s.diag + (size_t)(4 - par) * s.stride;
  T* const Fc = s.diag + (size_t)(5 + par) * s.stride;
  const T* const Fp1 = s.diag + (size_t)(6 - par) * s.stride;
  const int ilo = d > s.n ? d - s.n : 0;
  const int ihi = d < s.m ? d : s.m;
  uint8_t* const drow = s.dir + (size_t)d * s.ld - ilo;
  for (int i = ilo + tid; i <= ihi; i += nt) {
    const int j = d - i;
    T H, E, F;
    uint32_t code;
    if (i != 0 && j != 0) {
      const uint32_t colb = (uint32_t)(s.bx ? i - 1 : j - 1), cols = (uint32_t)(s.bx ? j - 1 : i - 1);
      const uint32_t nz = s.lnz[cols];
      const uint32_t* le = s.lst + cols;
      const int32_t* pb = s.pbig + colb;
      T sub = 0;
      TSQ_NO_UNROLL   // usually one to three letters: an unrolled body costs more in remainder branches than it saves
      for (uint32_t q = 0; q < nz; ++q, le += s.Ls) {
        const uint32_t e = *le;
        sub += (T)(e & 0xffffffu) * (T)pb[(e >> 24) * s.Lb];   // nsym * Lb words: a 32-bit offset
      }
      const T hl = Hp1[i], hu = Hp1[i - 1], hd = Hp2[i - 1];
      const T e_ext = Ep1[i] - s.GE, e_open = hl - s.GOE;
      const T f_ext = Fp1[i - 1] - s.GE, f_open = hu - s.GOE;
      const T dg = hd + sub;
      const bool eo = e_open >= e_ext, fo = f_open >= f_ext;
      E = eo ? e_open : e_ext;
      F = fo ? f_open : f_ext;
      H = dg;
      if (E > H) H = E;
      if (F > H) H = F;
      code = (H == dg ? 0u : (H == E ? 1u : 2u)) | (eo ? 4u : 0u) | (fo ? 8u : 0u);
    } else if (i == 0 && j == 0) {
      H = 0; E = NEG; F = NEG; code = 0;
    } else if (i == 0) {
      H = E = (T)(-s.GO - (T)j * s.GE); F = NEG; code = 1u | (j == 1 ? 4u : 0u);
    } else {
      H = F = (T)(-s.GO - (T)i * s.GE); E = NEG; code = 2u | (i == 1 ? 8u : 0u);
    }
    Hc[i] = H; Ec[i] = E; Fc[i] = F;
    drow[i] = (uint8_t)code;
  }
}

// phase 3, one thread.  score = H(Lx, Ly) (msa_final_score).
TSQ_HD void msa_walk_phase(const MsaTask& t, long long score) {
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
      if (ii > 0 && jj > 0) {
        const int d = ii + jj;
        c = t.dir[(size_t)d * ld + (size_t)(ii - (d > n ? d - n : 0))];
      }
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

#ifdef TSQ_DEVICE_IMPL
__global__ void __launch_bounds__(128) msa_leaf_kernel(const MsaLeaf* leaves, uint32_t n, uint32_t nsym) {
  for (uint32_t r = blockIdx.x; r < n; r += gridDim.x) msa_leaf_phase(leaves[r], nsym, (int)threadIdx.x, (int)blockDim.x);
}

// One merge by one CTA, in score type T.  smem_bytes: dynamic shared memory of the launch; a merge whose 7
// rolling diagonals fit uses it, any other its global scratch.
// The sweep of one merge with its rolling diagonals at `diag`.  Called once with the shared-memory array
// and once with the global scratch, so that each copy of the loop knows its address space (LDS/STS with
// 32-bit offsets instead of generic 64-bit addressing).
template <typename T>
__device__ __forceinline__ long long msa_sweep_cta(const MsaTask& t, const MsaConst& k, T* diag, const MsaTables* tab, int tid, int nt) {
  const MsaSweep<T> sw = msa_sweep_init<T>(t, k, diag, tab);
  const int last = sw.m + sw.n;
  int hc = 0;
  for (int d = 0; d <= last; ++d) {
    msa_diag_phase<T>(sw, d, hc, tid, nt);
    hc = hc == 2 ? 0 : hc + 1;
    __syncthreads();   // diagonal d complete and visible to the whole CTA before d + 1 starts
  }
  return msa_final_score<T>(sw);
}

// One merge by one CTA, in score type T.  smem_bytes: dynamic shared memory of the launch.  A merge whose 7 rolling
// diagonals fit uses it for them (its global scratch otherwise), and one whose column-score tables fit behind the
// diagonals copies them there after the prep phase.
template <typename T>
__device__ __forceinline__ void msa_merge_cta(const MsaTask& t, const MsaConst& k, uint32_t smem_bytes, T* smem) {
  const int tid = (int)threadIdx.x, nt = (int)blockDim.x;
  msa_prep_phase(t, k, tid, nt);
  __syncthreads();
  long long score;
  const size_t db = msa_round16(msa_diag_bytes(t.Lx, sizeof(T) == 4));
  if (db <= (size_t)smem_bytes) {
    const uint32_t Lb = msa_big_is_x(t) ? t.Lx : t.Ly, Ls = msa_big_is_x(t) ? t.Ly : t.Lx;
    const size_t pb = msa_round16((size_t)k.nsym * Lb * 4), lb = msa_round16((size_t)k.nsym * Ls * 4), nb = msa_round16((size_t)Ls * 4);
    if (db + pb + lb + nb <= (size_t)smem_bytes) {
      char* base = reinterpret_cast<char*>(smem) + db;
      int32_t* s_pbig = reinterpret_cast<int32_t*>(base);
      uint32_t* s_lst = reinterpret_cast<uint32_t*>(base + pb);
      uint32_t* s_lnz = reinterpret_cast<uint32_t*>(base + pb + lb);
      for (uint32_t q = (uint32_t)tid; q < k.nsym * Lb; q += (uint32_t)nt) s_pbig[q] = t.pbig[q];
      for (uint32_t q = (uint32_t)tid; q < Ls; q += (uint32_t)nt) {
        const uint32_t nz = t.lnz[q];
        s_lnz[q] = nz;
        for (uint32_t r = 0; r < nz; ++r) s_lst[(size_t)r * Ls + q] = t.lst[(size_t)r * Ls + q];   // only the entries in use
      }
      __syncthreads();
      const MsaTables tab{s_pbig, s_lst, s_lnz};
      score = msa_sweep_cta<T>(t, k, smem, &tab, tid, nt);
    } else {
      score = msa_sweep_cta<T>(t, k, smem, nullptr, tid, nt);
    }
  } else {
    score = msa_sweep_cta<T>(t, k, reinterpret_cast<T*>(t.diag), nullptr, tid, nt);
  }
  if (tid == 0) msa_walk_phase(t, score);
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
