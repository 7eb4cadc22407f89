/*
 * gotoh_simd.c -- inter-sequence SIMD statement (int16 lanes) of the oracle's recurrence, for the CPU BASELINE.
 *
 * TEST INFRASTRUCTURE ONLY, like gotoh_oracle.c: bench.py's cpu_baseline leg times it beside the scalar port so
 * that the GPU/CPU ratio is quoted against a competent CPU kernel, and tests/test_oracle.py checks it bit for bit
 * against the scalar oracle.  Nothing under tweakseq_b200/ or host/ may call it.
 *
 * Same spec as tsq_oracle_gotoh (matrix of tweakseq/Core/Annotations/Consensus.cpp:34-59).
 * Layout: one query against 32 subjects at a time, one subject per 16-bit lane (what the CUDA kernel does with one
 * subject per thread).  Per batch of subjects the substitution scores are laid out once as T[column][query letter] =
 * vector over the 32 subjects, then every query row is five vector operations per cell.  No -infinity: the boundary
 * gap states are seeded with H - (go + ge), which the recurrence cannot tell from -infinity.
 * Pairs whose scores could leave int16 (long sequences) go through the scalar routine.
 * GCC vector extensions: the same source compiles to AVX-512BW, AVX2 or plain SSE2 (target_clones picks at load);
 * built as C++ (oracle/Makefile) because only g++ accepts the element-wise vector ternary.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

extern "C" {
#include "gotoh_oracle.h"
}

#define LANES 32
typedef int16_t vec16 __attribute__((vector_size(2 * LANES), aligned(2 * LANES)));

static inline vec16 vmax(vec16 a, vec16 b) { return a > b ? a : b; }
static inline vec16 vsplat(int v) {
  vec16 r;
  for (int k = 0; k < LANES; k++) r[k] = (int16_t)v;
  return r;
}

typedef struct {
  const uint8_t *seqs;
  const uint64_t *offs;
  const uint32_t *lens;
  uint32_t n;
  const int8_t *mat;
  int nsym, go, ge;
  uint32_t row_begin, row_end;
  uint64_t base;          /* packed index of (row_begin, row_begin + 1) */
  int32_t *out;
  uint32_t nbatch, qchunk, nqchunks;
  uint64_t next, ntasks;  /* task = (subject batch, chunk of queries) */
  uint64_t cells;
  pthread_mutex_t mu;
} simd_job;

static inline uint64_t tri(uint64_t i, uint64_t j, uint64_t n) { return i * n - i * (i + 1) / 2 + (j - i - 1); }

/* one query (length m >= 1) against the batch whose tables are in T; H, E: (lmax + 1) vectors of scratch */
__attribute__((target_clones("arch=x86-64-v4", "arch=x86-64-v3", "default")))
static void query_vs_batch(const uint8_t *q, int m, const vec16 *T, int lmax, int go, int ge, vec16 *H, vec16 *E) {
  const vec16 vge = vsplat(ge), vgoe = vsplat(go + ge);
  H[0] = vsplat(0);
  for (int c = 1; c <= lmax; c++) {
    H[c] = vsplat(-(go + c * ge));
    E[c] = H[c] - vgoe;
  }
  for (int r = 1; r <= m; r++) {
    const vec16 *Ta = T + (size_t)q[r - 1] * (size_t)lmax;   /* T[letter][column] */
    vec16 diag = H[0];
    const vec16 hleft = vsplat(-(go + r * ge));
    H[0] = hleft;
    vec16 F = hleft - vgoe;
    for (int c = 1; c <= lmax; c++) {
      const vec16 e = E[c];
      const vec16 h = vmax(vmax(diag + Ta[c - 1], e), F);
      diag = H[c];
      H[c] = h;
      const vec16 t = h - vgoe;
      E[c] = vmax(e - vge, t);
      F = vmax(F - vge, t);
    }
  }
}

static void *simd_worker(void *arg) {
  simd_job *jb = (simd_job *)arg;
  vec16 *T = NULL, *H = NULL, *E = NULL;
  size_t capT = 0, capH = 0;
  uint64_t cells = 0;
  for (;;) {
    pthread_mutex_lock(&jb->mu);
    const uint64_t task = jb->next++;
    pthread_mutex_unlock(&jb->mu);
    if (task >= jb->ntasks) break;
    const uint32_t b = (uint32_t)(task / jb->nqchunks), qc = (uint32_t)(task % jb->nqchunks);
    const uint32_t j0 = b * LANES, j1 = j0 + LANES < jb->n ? j0 + LANES : jb->n;
    uint32_t i0 = jb->row_begin + qc * jb->qchunk, i1 = i0 + jb->qchunk;
    if (i1 > jb->row_end) i1 = jb->row_end;
    if (i1 > j1 - 1) i1 = j1 - 1;        /* only queries i with some j > i in the batch */
    if (i0 >= i1) continue;
    int lmax = 0;
    for (uint32_t j = j0; j < j1; j++)
      if ((int)jb->lens[j] > lmax) lmax = (int)jb->lens[j];
    if (lmax == 0) lmax = 1;
    const size_t needT = (size_t)jb->nsym * (size_t)lmax, needH = (size_t)lmax + 1;
    if (needT > capT) {
      free(T);
      T = (vec16 *)aligned_alloc(2 * LANES, needT * sizeof(vec16));
      capT = needT;
    }
    if (needH > capH) {
      free(H);
      free(E);
      H = (vec16 *)aligned_alloc(2 * LANES, needH * sizeof(vec16));
      E = (vec16 *)aligned_alloc(2 * LANES, needH * sizeof(vec16));
      capH = needH;
    }
    /* T[a][c][lane] = S(a, subject_lane[c]); columns past a subject's end score 0 (never read back) */
    for (int a = 0; a < jb->nsym; a++)
      for (int c = 0; c < lmax; c++) {
        vec16 v = vsplat(0);
        for (uint32_t j = j0; j < j1; j++)
          if (c < (int)jb->lens[j]) v[j - j0] = jb->mat[a * jb->nsym + jb->seqs[jb->offs[j] + c]];
        T[(size_t)a * lmax + c] = v;
      }
    int smax = 0;
    for (int k = 0; k < jb->nsym * jb->nsym; k++) {
      const int v = jb->mat[k] < 0 ? -jb->mat[k] : jb->mat[k];
      if (v > smax) smax = v;
    }
    for (uint32_t i = i0; i < i1; i++) {
      const int m = (int)jb->lens[i];
      const uint8_t *q = jb->seqs + jb->offs[i];
      const long bound = (long)3 * jb->go + (long)(m + lmax + 2) * jb->ge + (long)smax * (m > lmax ? m : lmax) + 64;
      const int simd_ok = m > 0 && bound < 32000;
      if (simd_ok) query_vs_batch(q, m, T, lmax, jb->go, jb->ge, H, E);
      for (uint32_t j = (i + 1 > j0 ? i + 1 : j0); j < j1; j++) {
        const int l = (int)jb->lens[j];
        int32_t s;
        if (simd_ok && l > 0) s = H[l][j - j0];
        else s = tsq_oracle_gotoh(q, m, jb->seqs + jb->offs[j], l, jb->mat, jb->nsym, jb->go, jb->ge);
        jb->out[tri(i, j, jb->n) - jb->base] = s;
        cells += (uint64_t)m * (uint64_t)l;
      }
    }
  }
  free(T);
  free(H);
  free(E);
  pthread_mutex_lock(&jb->mu);
  jb->cells += cells;
  pthread_mutex_unlock(&jb->mu);
  return NULL;
}

/* All pairs (i, j), row_begin <= i < row_end, i < j < n; out[packed(i, j) - packed(row_begin, row_begin + 1)].
 * Returns the number of DP cells. */
extern "C" uint64_t tsq_oracle_rows_simd(const uint8_t *seqs, const uint64_t *offs, const uint32_t *lens, uint32_t n,
                              const int8_t *mat, int nsym, int go, int ge, uint32_t row_begin, uint32_t row_end,
                              int32_t *out, int nthreads) {
  if (n < 2 || row_begin >= row_end) return 0;
  if (row_end > n - 1) row_end = n - 1;
  if (row_begin >= row_end) return 0;
  simd_job jb;
  memset(&jb, 0, sizeof jb);
  jb.seqs = seqs; jb.offs = offs; jb.lens = lens; jb.n = n;
  jb.mat = mat; jb.nsym = nsym; jb.go = go; jb.ge = ge;
  jb.row_begin = row_begin; jb.row_end = row_end;
  jb.base = tri(row_begin, row_begin + 1, n);
  jb.out = out;
  jb.nbatch = (n + LANES - 1) / LANES;
  jb.qchunk = 32;
  jb.nqchunks = (row_end - row_begin + jb.qchunk - 1) / jb.qchunk;
  jb.ntasks = (uint64_t)jb.nbatch * jb.nqchunks;
  pthread_mutex_init(&jb.mu, NULL);
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t th[256];
  int started = 0;
  for (int t = 1; t < nthreads; t++)
    if (pthread_create(&th[started], NULL, simd_worker, &jb) == 0) started++;
  simd_worker(&jb);
  for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
  pthread_mutex_destroy(&jb.mu);
  return jb.cells;
}
