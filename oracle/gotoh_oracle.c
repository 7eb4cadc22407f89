/*
 * gotoh_oracle.c -- scalar CPU restatement of the all-vs-all Gotoh distance-matrix path.
 *
 * TEST INFRASTRUCTURE ONLY (see gotoh_oracle.h).  PARITY UNPINNED by the reference: there is
 * no alignment arithmetic in groundstate/tweakseq to restate; the spec
 * implemented here is the one frozen in SURVEY.md section 8c.  What anchors it outside this repository are seven
 * published optima (tests/published_vectors.py: Durbin et al. 1998 fig. 2.5, the Needleman-Wunsch worked example,
 * the Biopython tutorial's affine example, the sample datasets of Rosalind's GLOB, GAFF, GCON on BLOSUM62 and EDIT);
 * everything else that pins it is listed in DESIGN.md section 3.2.
 *
 * What follows the reference:
 *   - matrix values + residue order      tweakseq/Core/Annotations/Consensus.cpp:34-59
 *   - letter -> index map (J,O,U,X -> X)  tweakseq/Core/Annotations/Consensus.cpp:61-69
 *   - non-letters are "not a residue"     tweakseq/Core/Annotations/Consensus.cpp:99-105
 *   - 7-bit residue cells, '-' kept by the editor and therefore stripped here
 *                                         tweakseq/Core/Sequence.cpp:57-69, Sequence.h:36-39
 */
#include "gotoh_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* Rows/columns in the order A R N D C Q E G H I L K M F P S T W Y V B Z X. */
const int8_t tsq_oracle_blosum62[23 * 23] = {
    /* A */ 4,  -1, -2, -2, 0,  -1, -1, 0,  -2, -1, -1, -1, -1, -2, -1, 1,  0,  -3, -2, 0,  -2, -1, 0,
    /* R */ -1, 5,  0,  -2, -3, 1,  0,  -2, 0,  -3, -2, 2,  -1, -3, -2, -1, -1, -3, -2, -3, -1, 0,  -1,
    /* N */ -2, 0,  6,  1,  -3, 0,  0,  0,  1,  -3, -3, 0,  -2, -3, -2, 1,  0,  -4, -2, -3, 3,  0,  -1,
    /* D */ -2, -2, 1,  6,  -3, 0,  2,  -1, -1, -3, -4, -1, -3, -3, -1, 0,  -1, -4, -3, -3, 4,  1,  -1,
    /* C */ 0,  -3, -3, -3, 9,  -3, -4, -3, -3, -1, -1, -3, -1, -2, -3, -1, -1, -2, -2, -1, -3, -3, -2,
    /* Q */ -1, 1,  0,  0,  -3, 5,  2,  -2, 0,  -3, -2, 1,  0,  -3, -1, 0,  -1, -2, -1, -2, 0,  3,  -1,
    /* E */ -1, 0,  0,  2,  -4, 2,  5,  -2, 0,  -3, -3, 1,  -2, -3, -1, 0,  -1, -3, -2, -2, 1,  4,  -1,
    /* G */ 0,  -2, 0,  -1, -3, -2, -2, 6,  -2, -4, -4, -2, -3, -3, -2, 0,  -2, -2, -3, -3, -1, -2, -1,
    /* H */ -2, 0,  1,  -1, -3, 0,  0,  -2, 8,  -3, -3, -1, -2, -1, -2, -1, -2, -2, 2,  -3, 0,  0,  -1,
    /* I */ -1, -3, -3, -3, -1, -3, -3, -4, -3, 4,  2,  -3, 1,  0,  -3, -2, -1, -3, -1, 3,  -3, -3, -1,
    /* L */ -1, -2, -3, -4, -1, -2, -3, -4, -3, 2,  4,  -2, 2,  0,  -3, -2, -1, -2, -1, 1,  -4, -3, -1,
    /* K */ -1, 2,  0,  -1, -3, 1,  1,  -2, -1, -3, -2, 5,  -1, -3, -1, 0,  -1, -3, -2, -2, 0,  1,  -1,
    /* M */ -1, -1, -2, -3, -1, 0,  -2, -3, -2, 1,  2,  -1, 5,  0,  -2, -1, -1, -1, -1, 1,  -3, -1, -1,
    /* F */ -2, -3, -3, -3, -2, -3, -3, -3, -1, 0,  0,  -3, 0,  6,  -4, -2, -2, 1,  3,  -1, -3, -3, -1,
    /* P */ -1, -2, -2, -1, -3, -1, -1, -2, -2, -3, -3, -1, -2, -4, 7,  -1, -1, -4, -3, -2, -2, -1, -2,
    /* S */ 1,  -1, 1,  0,  -1, 0,  0,  0,  -1, -2, -2, 0,  -1, -2, -1, 4,  1,  -3, -2, -2, 0,  0,  0,
    /* T */ 0,  -1, 0,  -1, -1, -1, -1, -2, -2, -1, -1, -1, -1, -2, -1, 1,  5,  -2, -2, 0,  -1, -1, 0,
    /* W */ -3, -3, -4, -4, -2, -2, -3, -2, -2, -3, -2, -3, -1, 1,  -4, -3, -2, 11, 2,  -3, -4, -3, -2,
    /* Y */ -2, -2, -2, -3, -2, -1, -2, -3, 2,  -1, -1, -2, -1, 3,  -3, -2, -2, 2,  7,  -1, -3, -2, -1,
    /* V */ 0,  -3, -3, -3, -1, -2, -2, -3, -3, 3,  1,  -2, 1,  -1, -2, -2, 0,  -3, -1, 4,  -3, -2, -1,
    /* B */ -2, -1, 3,  4,  -3, 0,  1,  -1, 0,  -3, -4, 0,  -3, -3, -2, 0,  -1, -4, -3, -3, 4,  1,  -1,
    /* Z */ -1, 0,  0,  1,  -3, 3,  4,  -2, 0,  -3, -3, 1,  -1, -3, -1, 0,  -1, -3, -2, -2, 1,  4,  -1,
    /* X */ 0,  -1, -1, -1, -2, -1, -1, -1, -1, -1, -1, -1, -1, -1, -2, 0,  0,  -2, -1, -1, -1, -1, -1,
};

/* A C G T N.  The reference has no nucleotide matrix (Core/DNA.h is colours only); these
 * EDNAFULL-style values are build-defined and this table is authoritative. */
const int8_t tsq_oracle_dna[5 * 5] = {
    /* A */ 5,  -4, -4, -4, -2,
    /* C */ -4, 5,  -4, -4, -2,
    /* G */ -4, -4, 5,  -4, -2,
    /* T */ -4, -4, -4, 5,  -2,
    /* N */ -2, -2, -2, -2, -1,
};

/* 'A'..'Z' -> row of tsq_oracle_blosum62 (Consensus.cpp:61-69). */
static const uint8_t protein_index[26] = {
    /* A */ 0,  /* B */ 20, /* C */ 4,  /* D */ 3,  /* E */ 6,  /* F */ 13, /* G */ 7,
    /* H */ 8,  /* I */ 9,  /* J */ 22, /* K */ 11, /* L */ 10, /* M */ 12, /* N */ 2,
    /* O */ 22, /* P */ 14, /* Q */ 5,  /* R */ 1,  /* S */ 15, /* T */ 16, /* U */ 22,
    /* V */ 19, /* W */ 17, /* X */ 22, /* Y */ 18, /* Z */ 21,
};

int tsq_oracle_nsym(int alphabet) { return alphabet == TSQ_ORACLE_NUCLEOTIDE ? 5 : 23; }

const int8_t *tsq_oracle_matrix(int alphabet) {
  return alphabet == TSQ_ORACLE_NUCLEOTIDE ? tsq_oracle_dna : tsq_oracle_blosum62;
}

static int is_gap_or_space(unsigned char c) {
  return c == '-' || c == '.' || c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' ||
         c == '\f';
}

size_t tsq_oracle_encode(int alphabet, const char *in, size_t n, uint8_t *out) {
  size_t k = 0;
  for (size_t i = 0; i < n; i++) {
    unsigned char c = (unsigned char)in[i];
    if (is_gap_or_space(c)) continue;
    if (c >= 'a' && c <= 'z') c = (unsigned char)(c - 'a' + 'A');
    if (alphabet == TSQ_ORACLE_NUCLEOTIDE) {
      uint8_t v = 4;
      if (c == 'A') v = 0;
      else if (c == 'C') v = 1;
      else if (c == 'G') v = 2;
      else if (c == 'T' || c == 'U') v = 3;
      out[k++] = v;
    } else {
      out[k++] = (c >= 'A' && c <= 'Z') ? protein_index[c - 'A'] : 22;
    }
  }
  return k;
}

static inline int32_t max2(int32_t a, int32_t b) { return a > b ? a : b; }

int32_t tsq_oracle_gotoh(const uint8_t *a, int m, const uint8_t *b, int n, const int8_t *mat,
                         int nsym, int go, int ge) {
  if (m == 0 && n == 0) return 0;
  if (m == 0) return -(go + n * ge);
  if (n == 0) return -(go + m * ge);
  const int32_t NEG = INT32_MIN / 2;
  const int32_t goe = go + ge;
  int32_t stackbuf[2 * 1025];
  int32_t *H = stackbuf;
  if (n > 1024) {
    H = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(n + 1));
    if (!H) return INT32_MIN;
  }
  int32_t *F = H + (n + 1);
  H[0] = 0;
  for (int j = 1; j <= n; j++) {
    H[j] = -(go + j * ge);
    F[j] = NEG;
  }
  for (int i = 1; i <= m; i++) {
    const int8_t *srow = mat + (size_t)a[i - 1] * nsym;
    int32_t diag = H[0];
    int32_t left = -(go + i * ge); /* H[i][0] */
    int32_t E = NEG;               /* E[i][0] */
    H[0] = left;
    for (int j = 1; j <= n; j++) {
      E = max2(E - ge, left - goe);
      int32_t f = max2(F[j] - ge, H[j] - goe);
      int32_t h = diag + srow[b[j - 1]];
      h = max2(h, max2(E, f));
      diag = H[j];
      H[j] = h;
      F[j] = f;
      left = h;
    }
  }
  int32_t r = H[n];
  if (H != stackbuf) free(H);
  return r;
}

int32_t tsq_oracle_self_score(const uint8_t *a, int m, const int8_t *mat, int nsym) {
  int32_t s = 0;
  for (int k = 0; k < m; k++) s += mat[(size_t)a[k] * nsym + a[k]];
  return s;
}

double tsq_oracle_distance(int32_t s_ij, int32_t s_ii, int32_t s_jj) {
  int32_t mn = s_ii < s_jj ? s_ii : s_jj;
  if (mn <= 0) return 1.0;
  double q = (double)s_ij / (double)mn;
  return 1.0 - q;
}

/* ---- Kimura-corrected identity distance (SURVEY.md 8f-2 "+ Kimura") ----
 * ClustalW corrects the observed protein distance D = 1 - identities / min(len) for multiple substitutions with
 * Kimura's formula  d = -ln(1 - D - D^2 / 5),  valid for D < 0.75 (above it ClustalW switches to a lookup table that
 * the reference does not hold: those pairs are an error, never a guess).  Bit-equality between CPU and GPU cannot
 * rest on two libm's, so the logarithm is part of the spec: ln x = e ln 2 + 2 z (1 + w/3 + w^2/5 + ... + w^11/23)
 * with x = m 2^e, m in [sqrt(1/2), sqrt(2)), z = (m - 1)/(m + 1), w = z z, evaluated by Horner's rule in IEEE double
 * operations in exactly this order, no fused multiply-add (this file is built with -ffp-contract=off; the kernel
 * uses the _rn intrinsics).  Within 2 ulp of the true logarithm on the range used (tests/test_identity.py). */
double tsq_oracle_ln(double x) {
  union { double d; uint64_t u; } v;
  v.d = x;
  int e = (int)((v.u >> 52) & 0x7ff) - 1022;                 /* x = m 2^e, m in [1/2, 1) */
  v.u = (v.u & 0x000fffffffffffffull) | 0x3fe0000000000000ull;
  double m = v.d;
  if (m < 0.70710678118654752440) {
    m = m * 2.0;
    e -= 1;
  }
  const double z = (m - 1.0) / (m + 1.0);
  const double w = z * z;
  double p = 1.0 / 23.0;
  p = p * w + 1.0 / 21.0;
  p = p * w + 1.0 / 19.0;
  p = p * w + 1.0 / 17.0;
  p = p * w + 1.0 / 15.0;
  p = p * w + 1.0 / 13.0;
  p = p * w + 1.0 / 11.0;
  p = p * w + 1.0 / 9.0;
  p = p * w + 1.0 / 7.0;
  p = p * w + 1.0 / 5.0;
  p = p * w + 1.0 / 3.0;
  p = p * w + 1.0;
  const double lnm = (2.0 * z) * p;
  return (double)e * 0.693147180559945309417 + lnm;
}

/* identities, shorter length -> (corrected distance, ok).  ok = 0: D >= 0.75, the formula does not apply. */
double tsq_oracle_kimura(int32_t identities, int32_t min_len, int *ok) {
  if (ok) *ok = 1;
  if (min_len <= 0) {           /* no residues to compare: the uncorrected convention, D = 1 */
    if (ok) *ok = 0;
    return 1.0;
  }
  const double D = 1.0 - (double)identities / (double)min_len;
  if (!(D < 0.75)) {
    if (ok) *ok = 0;
    return D;
  }
  const double t = D * D;
  const double u = t / 5.0;
  const double a = 1.0 - D;
  const double arg = a - u;
  return 0.0 - tsq_oracle_ln(arg);
}

uint64_t tsq_oracle_pair_index(uint64_t i, uint64_t j, uint64_t n) {
  return i * n - i * (i + 1) / 2 + (j - i - 1);
}

/* ---- multithreaded drivers (also the CPU baseline timed by bench.py) ---- */

typedef struct {
  const uint8_t *seqs;
  const uint64_t *offs;
  const uint32_t *lens;
  uint32_t n;
  const int8_t *mat;
  int nsym, go, ge;
  uint64_t begin, end;
  const uint32_t *pi, *pj; /* explicit list, or NULL for the packed range */
  int32_t *out;
  uint64_t next; /* atomic chunk cursor */
  uint64_t chunk; /* pairs handed out per fetch */
  uint64_t cells; /* atomic */
} job_t;


static void unpack_pair(uint64_t p, uint64_t n, uint32_t *pi, uint32_t *pj) {
  /* invert p = i*n - i(i+1)/2 + (j-i-1) by walking rows from a float guess */
  double nn = (double)n;
  double disc = (2.0 * nn - 1.0) * (2.0 * nn - 1.0) - 8.0 * (double)p;
  uint64_t i = 0;
  if (disc > 0) {
    double r = ((2.0 * nn - 1.0) - __builtin_sqrt(disc)) / 2.0;
    if (r > 0) i = (uint64_t)r;
  }
  if (i >= n - 1) i = n - 2;
  while (i > 0 && tsq_oracle_pair_index(i, i + 1, n) > p) i--;
  while (i + 2 < n && tsq_oracle_pair_index(i + 1, i + 2, n) <= p) i++;
  uint64_t j = p - tsq_oracle_pair_index(i, i + 1, n) + i + 1;
  *pi = (uint32_t)i;
  *pj = (uint32_t)j;
}

static void *worker(void *arg) {
  job_t *jb = (job_t *)arg;
  uint64_t cells = 0;
  for (;;) {
    uint64_t s = __atomic_fetch_add(&jb->next, jb->chunk, __ATOMIC_RELAXED);
    if (s >= jb->end) break;
    uint64_t e = s + jb->chunk < jb->end ? s + jb->chunk : jb->end;
    uint32_t i = 0, j = 0;
    if (!jb->pi) unpack_pair(s, jb->n, &i, &j);
    for (uint64_t p = s; p < e; p++) {
      if (jb->pi) {
        i = jb->pi[p];
        j = jb->pj[p];
      }
      jb->out[p - jb->begin] =
          tsq_oracle_gotoh(jb->seqs + jb->offs[i], (int)jb->lens[i], jb->seqs + jb->offs[j],
                           (int)jb->lens[j], jb->mat, jb->nsym, jb->go, jb->ge);
      cells += (uint64_t)jb->lens[i] * jb->lens[j];
      if (!jb->pi) {
        if (++j >= jb->n) {
          i++;
          j = i + 1;
        }
      }
    }
  }
  __atomic_fetch_add(&jb->cells, cells, __ATOMIC_RELAXED);
  return NULL;
}

static uint64_t run_job(job_t *jb, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  /* small jobs of long pairs: hand out single pairs so that every thread gets work */
  uint64_t per = (jb->end - jb->begin) / ((uint64_t)nthreads * 4);
  jb->chunk = per < 1 ? 1 : (per > 64 ? 64 : per);
  pthread_t th[256];
  int started = 0;
  for (int t = 1; t < nthreads; t++)
    if (pthread_create(&th[started], NULL, worker, jb) == 0) started++;
  worker(jb);
  for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
  return jb->cells;
}

uint64_t tsq_oracle_all_pairs(const uint8_t *seqs, const uint64_t *offs, const uint32_t *lens,
                              uint32_t n, const int8_t *mat, int nsym, int go, int ge,
                              uint64_t pair_begin, uint64_t pair_end, int32_t *out,
                              int nthreads) {
  if (n < 2 || pair_end <= pair_begin) return 0;
  job_t jb;
  memset(&jb, 0, sizeof jb);
  jb.seqs = seqs; jb.offs = offs; jb.lens = lens; jb.n = n;
  jb.mat = mat; jb.nsym = nsym; jb.go = go; jb.ge = ge;
  jb.begin = pair_begin; jb.end = pair_end; jb.out = out; jb.next = pair_begin;
  return run_job(&jb, nthreads);
}

uint64_t tsq_oracle_pair_list(const uint8_t *seqs, const uint64_t *offs, const uint32_t *lens,
                              const int8_t *mat, int nsym, int go, int ge, const uint32_t *pi,
                              const uint32_t *pj, uint64_t npairs, int32_t *out, int nthreads) {
  if (npairs == 0) return 0;
  job_t jb;
  memset(&jb, 0, sizeof jb);
  jb.seqs = seqs; jb.offs = offs; jb.lens = lens; jb.n = 0;
  jb.mat = mat; jb.nsym = nsym; jb.go = go; jb.ge = ge;
  jb.begin = 0; jb.end = npairs; jb.pi = pi; jb.pj = pj; jb.out = out; jb.next = 0;
  return run_job(&jb, nthreads);
}

/* ---- UPGMA guide tree, naive O(n^3) ---- */
void tsq_oracle_upgma(const double *packed, uint32_t n, uint32_t *left, uint32_t *right, double *height) {
  if (n < 2) return;
  double *D = (double *)malloc(sizeof(double) * (size_t)n * n);
  uint32_t *size = (uint32_t *)malloc(sizeof(uint32_t) * n), *node = (uint32_t *)malloc(sizeof(uint32_t) * n);
  unsigned char *act = (unsigned char *)malloc(n);
  for (uint32_t i = 0; i < n; i++) {
    size[i] = 1; node[i] = i; act[i] = 1;
    for (uint32_t j = i + 1; j < n; j++) {
      double v = packed[tsq_oracle_pair_index(i, j, n)];
      D[(size_t)i * n + j] = v;
      D[(size_t)j * n + i] = v;
    }
  }
  for (uint32_t t = 0; t + 1 < n; t++) {
    int found = 0;
    uint32_t a = 0, b = 0;
    double best = 0;
    for (uint32_t i = 0; i < n; i++) {
      if (!act[i]) continue;
      for (uint32_t j = i + 1; j < n; j++) {
        if (!act[j]) continue;
        double v = D[(size_t)i * n + j];
        if (!found || v < best) { found = 1; best = v; a = i; b = j; }   /* scan order = tie order */
      }
    }
    left[t] = node[a]; right[t] = node[b];
    volatile double half = best * 0.5;
    height[t] = half;
    double da = (double)size[a], db = (double)size[b], ds = (double)(size[a] + size[b]);
    for (uint32_t k = 0; k < n; k++) {
      if (!act[k] || k == a || k == b) continue;
      volatile double x = da * D[(size_t)a * n + k];
      volatile double y = db * D[(size_t)b * n + k];
      volatile double sum = x + y;
      double nd = sum / ds;
      D[(size_t)a * n + k] = nd;
      D[(size_t)k * n + a] = nd;
    }
    act[b] = 0; size[a] += size[b]; node[a] = n + t;
  }
  free(D); free(size); free(node); free(act);
}

/* ---- identity-aware score ---- */
void tsq_oracle_gotoh_id(const uint8_t *a, int m, const uint8_t *b, int n, const int8_t *mat, int nsym,
                         int go, int ge, int32_t *score, int32_t *identities) {
  *identities = 0;
  if (m == 0 || n == 0) {
    *score = tsq_oracle_gotoh(a, m, b, n, mat, nsym, go, ge);
    return;
  }
  const int64_t M = (int64_t)1 << 32;
  const int64_t NEG = INT64_MIN / 4;
  const int64_t gok = go * M, gek = ge * M, goek = gok + gek;
  int64_t *H = (int64_t *)malloc(sizeof(int64_t) * 2 * (size_t)(n + 1));
  int64_t *F = H + (n + 1);
  H[0] = 0;
  for (int j = 1; j <= n; j++) { H[j] = -(gok + j * gek); F[j] = NEG; }
  for (int i = 1; i <= m; i++) {
    const int8_t *srow = mat + (size_t)a[i - 1] * nsym;
    int64_t diag = H[0], left = -(gok + i * gek), E = NEG;
    H[0] = left;
    for (int j = 1; j <= n; j++) {
      int64_t e1 = E - gek, e2 = left - goek;
      E = e1 > e2 ? e1 : e2;
      int64_t f1 = F[j] - gek, f2 = H[j] - goek;
      int64_t f = f1 > f2 ? f1 : f2;
      int64_t h = diag + srow[b[j - 1]] * M + (a[i - 1] == b[j - 1] ? 1 : 0);
      if (E > h) h = E;
      if (f > h) h = f;
      diag = H[j]; H[j] = h; F[j] = f; left = h;
    }
  }
  int64_t key = H[n];
  int64_t sc = key >> 32;          /* floor division by 2^32 */
  *score = (int32_t)sc;
  *identities = (int32_t)(key - sc * M);
  free(H);
}

/* ---- one optimal alignment with its path ("emitting pairwise alignments") ----
 * Full H, E, F matrices, row by row; the walk back from (m, n) re-derives every decision from the
 * stored values.  Tie rules (build-defined, shared with the CUDA kernel): H prefers the diagonal,
 * then E (gap in a), then F (gap in b); a gap run that can equally be opened or extended was opened.
 * out_a / out_b: m + n + 1 bytes each, symbols or 0xff for a gap, written first column first.
 * Returns the number of columns; *score = H(m, n). */
uint32_t tsq_oracle_traceback(const uint8_t *a, int m, const uint8_t *b, int n, const int8_t *mat, int nsym,
                              int go, int ge, uint8_t *out_a, uint8_t *out_b, int32_t *score) {
  const int32_t NEG = -(1 << 29);
  const int goe = go + ge;
  const size_t ld = (size_t)n + 1, cells = ((size_t)m + 1) * ld;
  int32_t *H = (int32_t *)malloc(sizeof(int32_t) * 3 * cells);
  int32_t *E = H + cells, *F = E + cells;
  H[0] = 0; E[0] = NEG; F[0] = NEG;
  for (int j = 1; j <= n; j++) { H[j] = E[j] = -go - j * ge; F[j] = NEG; }
  for (int i = 1; i <= m; i++) {
    int32_t *h = H + (size_t)i * ld, *e = E + (size_t)i * ld, *f = F + (size_t)i * ld;
    const int32_t *hu = h - ld, *fu = f - ld;
    const int8_t *srow = mat + (size_t)a[i - 1] * nsym;
    h[0] = f[0] = -go - i * ge; e[0] = NEG;
    for (int j = 1; j <= n; j++) {
      int32_t e1 = e[j - 1] - ge, e2 = h[j - 1] - goe;
      int32_t f1 = fu[j] - ge, f2 = hu[j] - goe;
      int32_t dg = hu[j - 1] + srow[b[j - 1]];
      e[j] = e2 >= e1 ? e2 : e1;
      f[j] = f2 >= f1 ? f2 : f1;
      int32_t best = dg;
      if (e[j] > best) best = e[j];
      if (f[j] > best) best = f[j];
      h[j] = best;
    }
  }
  *score = H[(size_t)m * ld + n];
  uint8_t *ra = (uint8_t *)malloc(2 * ((size_t)m + n) + 2), *rb = ra + (size_t)m + n + 1;
  uint32_t k = 0;
  int i = m, j = n, state = 0;
  while (i > 0 || j > 0) {
    if (i == 0) { ra[k] = 0xff; rb[k] = b[j - 1]; j--; k++; continue; }
    if (j == 0) { ra[k] = a[i - 1]; rb[k] = 0xff; i--; k++; continue; }
    const size_t at = (size_t)i * ld + j;
    if (state == 0) {
      if (H[at] == H[at - ld - 1] + mat[(size_t)a[i - 1] * nsym + b[j - 1]]) { ra[k] = a[i - 1]; rb[k] = b[j - 1]; i--; j--; k++; }
      else state = (H[at] == E[at]) ? 1 : 2;
    } else if (state == 1) {
      ra[k] = 0xff; rb[k] = b[j - 1]; k++;
      if (E[at] == H[at - 1] - goe) state = 0;
      j--;
    } else {
      ra[k] = a[i - 1]; rb[k] = 0xff; k++;
      if (F[at] == H[at - ld] - goe) state = 0;
      i--;
    }
  }
  for (uint32_t c = 0; c < k; c++) { out_a[c] = ra[k - 1 - c]; out_b[c] = rb[k - 1 - c]; }
  free(ra); free(H);
  return k;
}

/* ---- consensus annotation: Consensus.cpp:80-161 restated ---- */
void tsq_oracle_consensus(const char *const *rows, uint32_t nrows, uint32_t ncols, double plurality, char *out) {
  if (nrows == 0) { for (uint32_t c = 0; c < ncols; c++) out[c] = '?'; return; }
  double *scores = (double *)malloc(sizeof(double) * nrows);
  unsigned char *rindex = (unsigned char *)malloc((size_t)nrows * ncols);
  for (uint32_t s = 0; s < nrows; s++)                       /* :95-106 */
    for (uint32_t c = 0; c < ncols; c++) {
      int idx = (int)(unsigned char)rows[s][c] - 65;
      rindex[(size_t)s * ncols + c] = (idx < 0 || idx > 25) ? 99 : protein_index[idx];
    }
  for (uint32_t c = 0; c < ncols; c++) {                     /* :108-152 */
    double hiScore = 0.0, riMatches = 0;
    uint32_t riHiScore = 0;
    for (uint32_t ri = 0; ri < nrows; ri++) {
      scores[ri] = 0.0;
      double matches = 0.0;
      for (uint32_t rj = 0; rj < nrows; rj++) {
        if (ri == rj) continue;
        double weight = 1.0;
        int resi = rindex[(size_t)ri * ncols + c], resj = rindex[(size_t)rj * ncols + c];
        if (resi == 99 && resj == 99) {
          scores[ri] += weight;
          if (weight > 0) matches += weight;
        } else if (resi == 99 || resj == 99) {
          scores[ri] += -4 * weight;
        } else {
          double tmp = tsq_oracle_blosum62[resi * 23 + resj] * weight;
          scores[ri] += tmp;
          if (tmp > 0) matches += weight;
        }
      }
      if (ri == 0) { hiScore = scores[ri]; riHiScore = ri; riMatches = matches; }
      else if (scores[ri] > hiScore) { hiScore = scores[ri]; riHiScore = ri; riMatches = matches; }
    }
    out[c] = riMatches >= plurality ? rows[riHiScore][c] : '?';
  }
  free(scores); free(rindex);
}

/* ---- progressive multiple alignment along a guide tree (the step after SURVEY.md 8f-1) ----
 * What the external aligner does with the tree (tweakseq/Core/ClustalO.cpp:48-52 run) and what
 * Project::readNewAlignment (tweakseq/Core/Project.cpp:908-1032) ingests: equal-length gapped rows.
 * Build-defined spec (the reference holds no aligner): the n-1 merges are applied in order; merge t
 * aligns the alignments of clusters left[t] (X, its columns along i) and right[t] (Y, along j) by the
 * Gotoh recurrence of SURVEY 8c over COLUMNS, with integer sum-of-pairs scores
 *     sub(i, j) = sum over rows x of X, y of Y with residues in columns i, j of S(x_i, y_j)
 * (a residue facing a gap scores 0) and a gap of k columns costing |X| |Y| (go + k ge); end gaps are
 * penalised.  For two single sequences this is exactly the pairwise alignment of tsq_oracle_traceback,
 * tie rules included (diagonal, then gap in X, then gap in Y; open rather than extend).  The merged
 * cluster lists X's rows, then Y's.  This restatement keeps explicit row lists and recounts every
 * column from the rows; all arithmetic is int64.
 * rows_out: n x (*ncols_out) bytes, row r = submitted sequence r, symbols or 0xff for a gap (malloc'd,
 * caller frees); merge_scores (n-1 entries or NULL) receives H(Lx, Ly) of every merge.  Returns 0. */
typedef struct {
  uint32_t nmem, L;
  uint32_t *mem;
  uint8_t *rows;   /* nmem x L */
} msa_cluster;

static int64_t msa_sub(const uint32_t *cx, const uint32_t *cy, const int8_t *mat, int nsym) {
  int64_t s = 0;
  for (int a = 0; a < nsym; a++) {
    if (!cx[a]) continue;
    for (int b = 0; b < nsym; b++)
      if (cy[b]) s += (int64_t)cx[a] * cy[b] * mat[a * nsym + b];
  }
  return s;
}

static uint32_t *msa_counts(const msa_cluster *c, int nsym) {
  uint32_t *cnt = (uint32_t *)calloc((size_t)(c->L ? c->L : 1) * nsym, sizeof(uint32_t));
  for (uint32_t r = 0; r < c->nmem; r++)
    for (uint32_t k = 0; k < c->L; k++) {
      uint8_t v = c->rows[(size_t)r * c->L + k];
      if (v != 0xff) cnt[(size_t)k * nsym + v]++;
    }
  return cnt;
}

int tsq_oracle_msa(const uint8_t *seqs, const uint64_t *offs, const uint32_t *lens, uint32_t n, const int8_t *mat,
                   int nsym, int go, int ge, const uint32_t *left, const uint32_t *right, uint8_t **rows_out,
                   uint32_t *ncols_out, int64_t *merge_scores) {
  *rows_out = NULL; *ncols_out = 0;
  if (n == 0) return 0;
  const int64_t NEG = -((int64_t)1 << 60);
  msa_cluster *cl = (msa_cluster *)calloc(2 * (size_t)n - 1, sizeof(msa_cluster));
  for (uint32_t r = 0; r < n; r++) {
    cl[r].nmem = 1; cl[r].L = lens[r];
    cl[r].mem = (uint32_t *)malloc(sizeof(uint32_t)); cl[r].mem[0] = r;
    cl[r].rows = (uint8_t *)malloc(lens[r] ? lens[r] : 1);
    memcpy(cl[r].rows, seqs + offs[r], lens[r]);
  }
  for (uint32_t t = 0; t + 1 < n; t++) {
    msa_cluster *X = &cl[left[t]], *Y = &cl[right[t]], *Z = &cl[n + t];
    const int m = (int)X->L, q = (int)Y->L;
    uint32_t *cx = msa_counts(X, nsym), *cy = msa_counts(Y, nsym);
    const int64_t w = (int64_t)X->nmem * Y->nmem, GE = w * ge, GO = w * go, GOE = GO + GE;
    const size_t ld = (size_t)q + 1, cells = ((size_t)m + 1) * ld;
    int64_t *H = (int64_t *)malloc(sizeof(int64_t) * 3 * cells), *E = H + cells, *F = E + cells;
    int64_t *S = (int64_t *)malloc(sizeof(int64_t) * cells);   /* sub(i, j), kept for the walk back */
    H[0] = 0; E[0] = NEG; F[0] = NEG;
    for (int j = 1; j <= q; j++) { H[j] = E[j] = -GO - j * GE; F[j] = NEG; }
    for (int i = 1; i <= m; i++) {
      int64_t *h = H + (size_t)i * ld, *e = E + (size_t)i * ld, *f = F + (size_t)i * ld;
      const int64_t *hu = h - ld, *fu = f - ld;
      h[0] = f[0] = -GO - i * GE; e[0] = NEG;
      for (int j = 1; j <= q; j++) {
        const int64_t sub = msa_sub(cx + (size_t)(i - 1) * nsym, cy + (size_t)(j - 1) * nsym, mat, nsym);
        S[(size_t)i * ld + j] = sub;
        const int64_t e1 = e[j - 1] - GE, e2 = h[j - 1] - GOE;
        const int64_t f1 = fu[j] - GE, f2 = hu[j] - GOE;
        const int64_t dg = hu[j - 1] + sub;
        e[j] = e2 >= e1 ? e2 : e1;
        f[j] = f2 >= f1 ? f2 : f1;
        int64_t best = dg;
        if (e[j] > best) best = e[j];
        if (f[j] > best) best = f[j];
        h[j] = best;
      }
    }
    if (merge_scores) merge_scores[t] = H[(size_t)m * ld + q];
    /* walk back: per merged column (last first) the X column and Y column it takes, -1 = gap */
    int *pi = (int *)malloc(sizeof(int) * 2 * ((size_t)m + q + 1)), *pj = pi + m + q + 1;
    uint32_t k = 0;
    int i = m, j = q, state = 0;
    while (i > 0 || j > 0) {
      if (i == 0) { pi[k] = -1; pj[k] = j - 1; j--; k++; continue; }
      if (j == 0) { pi[k] = i - 1; pj[k] = -1; i--; k++; continue; }
      const size_t at = (size_t)i * ld + j;
      if (state == 0) {
        if (H[at] == H[at - ld - 1] + S[at]) { pi[k] = i - 1; pj[k] = j - 1; i--; j--; k++; }
        else state = (H[at] == E[at]) ? 1 : 2;
      } else if (state == 1) {
        pi[k] = -1; pj[k] = j - 1; k++;
        if (E[at] == H[at - 1] - GOE) state = 0;
        j--;
      } else {
        pi[k] = i - 1; pj[k] = -1; k++;
        if (F[at] == H[at - ld] - GOE) state = 0;
        i--;
      }
    }
    Z->nmem = X->nmem + Y->nmem; Z->L = k;
    Z->mem = (uint32_t *)malloc(sizeof(uint32_t) * Z->nmem);
    memcpy(Z->mem, X->mem, sizeof(uint32_t) * X->nmem);
    memcpy(Z->mem + X->nmem, Y->mem, sizeof(uint32_t) * Y->nmem);
    Z->rows = (uint8_t *)malloc((size_t)Z->nmem * (k ? k : 1));
    for (uint32_t c = 0; c < k; c++) {
      const int xi = pi[k - 1 - c], yj = pj[k - 1 - c];
      for (uint32_t r = 0; r < X->nmem; r++)
        Z->rows[(size_t)r * k + c] = xi < 0 ? 0xff : X->rows[(size_t)r * X->L + xi];
      for (uint32_t r = 0; r < Y->nmem; r++)
        Z->rows[(size_t)(X->nmem + r) * k + c] = yj < 0 ? 0xff : Y->rows[(size_t)r * Y->L + yj];
    }
    free(pi); free(S); free(H); free(cx); free(cy);
    free(X->rows); free(X->mem); X->rows = NULL; X->mem = NULL;
    free(Y->rows); free(Y->mem); Y->rows = NULL; Y->mem = NULL;
  }
  msa_cluster *R = &cl[n == 1 ? 0 : 2 * (size_t)n - 2];
  const uint32_t L = R->L;
  uint8_t *out = (uint8_t *)malloc((size_t)n * (L ? L : 1));
  for (uint32_t r = 0; r < R->nmem; r++) memcpy(out + (size_t)R->mem[r] * L, R->rows + (size_t)r * L, L);
  free(R->rows); free(R->mem); free(cl);
  *rows_out = out; *ncols_out = L;
  return 0;
}

void tsq_oracle_free(void *p) { free(p); }
