/*
 * gotoh_oracle.h -- CPU oracle for the all-vs-all Gotoh distance-matrix path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under tweakseq_b200/ or host/ may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and there only as the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference (groundstate/tweakseq) ships no alignment arithmetic and no
 * test vectors for this path; it shells out to an external clustalo
 * binary (tweakseq/Core/ClustalO.cpp:48-52, tweakseq/UI/SeqEditMainWin.cpp:1654-1660) that is
 * neither vendored nor pinned.  The oracle therefore restates the frozen spec of SURVEY.md
 * section 8c; what it takes from the reference is the substitution matrix and letter map
 * (tweakseq/Core/Annotations/Consensus.cpp:34-69) and the residue-cell filtering rule
 * (tweakseq/Core/Sequence.cpp:57-69).  It is pinned against hand-derived known answers and
 * an independent numpy formulation in tests/.  What the reference does implement on this path
 * (consensus, FASTA reader / writer, cell filter, tool wrappers) is compiled from its own sources
 * into oracle/_ref/ (oracle/Makefile) and pins those pieces directly; see DESIGN.md section 3.1.
 */
#ifndef TSQ_GOTOH_ORACLE_H
#define TSQ_GOTOH_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TSQ_ORACLE_PROTEIN 0
#define TSQ_ORACLE_NUCLEOTIDE 1

/* 23x23, order ARNDCQEGHILKMFPSTWYVBZX (Consensus.cpp:34-59). */
extern const int8_t tsq_oracle_blosum62[23 * 23];
/* 5x5, order ACGTN: match +5, mismatch -4, N vs ACGT -2, N vs N -1. */
extern const int8_t tsq_oracle_dna[5 * 5];

/* Number of symbols of an alphabet (23 or 5). */
int tsq_oracle_nsym(int alphabet);
/* Built-in matrix of an alphabet. */
const int8_t *tsq_oracle_matrix(int alphabet);

/* ASCII -> symbol indices; returns the encoded length (gaps '-', '.', whitespace removed). */
size_t tsq_oracle_encode(int alphabet, const char *in, size_t n, uint8_t *out);

/* Global affine-gap (Gotoh) score H[m][n]; a gap of length k costs go + k*ge. */
int32_t tsq_oracle_gotoh(const uint8_t *a, int m, const uint8_t *b, int n,
                         const int8_t *mat, int nsym, int go, int ge);

/* Sum of mat[a_k][a_k]. */
int32_t tsq_oracle_self_score(const uint8_t *a, int m, const int8_t *mat, int nsym);

/* d = 1 - s_ij / min(s_ii, s_jj); 1.0 if that minimum is <= 0. */
double tsq_oracle_distance(int32_t s_ij, int32_t s_ii, int32_t s_jj);

/* Kimura-corrected identity distance -ln(1 - D - D^2/5), D = 1 - identities/min_len, for D < 0.75 (*ok = 0 and the
 * uncorrected D above: ClustalW's table for that range is not restated).  tsq_oracle_ln is the logarithm the spec
 * fixes operation by operation, so that CPU and GPU agree bit for bit. */
double tsq_oracle_ln(double x);
double tsq_oracle_kimura(int32_t identities, int32_t min_len, int *ok);

/* Packed upper-triangle index of (i,j), i<j, over n sequences. */
uint64_t tsq_oracle_pair_index(uint64_t i, uint64_t j, uint64_t n);

/*
 * Scores of the packed pairs [pair_begin, pair_end) into out[0 .. pair_end-pair_begin),
 * using nthreads POSIX threads.  seqs = concatenated encoded residues, offs[i] = start of
 * sequence i, lens[i] = its length.  Returns the number of DP cells evaluated.
 */
uint64_t tsq_oracle_all_pairs(const uint8_t *seqs, const uint64_t *offs, const uint32_t *lens,
                              uint32_t n, const int8_t *mat, int nsym, int go, int ge,
                              uint64_t pair_begin, uint64_t pair_end, int32_t *out,
                              int nthreads);

/* Scores of an explicit list of pairs (pi[k], pj[k]); same conventions. */
uint64_t tsq_oracle_pair_list(const uint8_t *seqs, const uint64_t *offs, const uint32_t *lens,
                              const int8_t *mat, int nsym, int go, int ge, const uint32_t *pi,
                              const uint32_t *pj, uint64_t npairs, int32_t *out, int nthreads);

/*
 * The CPU BASELINE kernel (gotoh_simd.c): the same recurrence with one subject per int16 SIMD lane, 32 subjects
 * per query pass (AVX-512BW / AVX2 / SSE2 by target_clones).  All pairs (i, j) with row_begin <= i < row_end and
 * i < j < n; out[packed(i, j) - packed(row_begin, row_begin + 1)].  Bit-identical to tsq_oracle_all_pairs
 * (tests/test_oracle.py); pairs whose scores could leave int16 run through the scalar routine.  Returns DP cells.
 */
uint64_t tsq_oracle_rows_simd(const uint8_t *seqs, const uint64_t *offs, const uint32_t *lens, uint32_t n,
                              const int8_t *mat, int nsym, int go, int ge, uint32_t row_begin, uint32_t row_end,
                              int32_t *out, int nthreads);

/*
 * Identity-aware score: the Gotoh score of tsq_oracle_gotoh() and, among all
 * alignments that reach it, the largest number of columns pairing identical symbols.  Restated as
 * one DP over 64-bit keys  score * 2^32 + identities  (every substitution score and gap penalty
 * scaled by 2^32, +1 for identical symbols).  Either length 0: score as tsq_oracle_gotoh, 0 identities.
 */
void tsq_oracle_gotoh_id(const uint8_t *a, int m, const uint8_t *b, int n, const int8_t *mat, int nsym,
                         int go, int ge, int32_t *score, int32_t *identities);

/*
 * One optimal alignment with its path.  out_a / out_b: m + n + 1 bytes each,
 * symbols or 0xff for a gap.  Returns the column count; *score = the Gotoh score.  Tie rules:
 * diagonal, then gap in a, then gap in b; open rather than extend.
 */
uint32_t tsq_oracle_traceback(const uint8_t *a, int m, const uint8_t *b, int n, const int8_t *mat, int nsym,
                              int go, int ge, uint8_t *out_a, uint8_t *out_b, int32_t *score);

/*
 * Consensus annotation: line-by-line restatement of Consensus::calculate,
 * tweakseq/Core/Annotations/Consensus.cpp:80-161, doubles and all: rows[r] has ncols characters
 * (already `& 0xff`), out gets ncols characters, '?' where the winner's matches < plurality.
 */
void tsq_oracle_consensus(const char *const *rows, uint32_t nrows, uint32_t ncols, double plurality, char *out);

/*
 * UPGMA guide tree of a packed fp64 distance matrix (spec in
 * tweakseq_b200/csrc/upgma.cuh): naive O(n^3) restatement.  Step t merges the active slot pair
 * (a < b) of smallest distance (ties: smallest a, then smallest b), the merged cluster keeps slot a,
 * d(a,k) <- (|a| d(a,k) + |b| d(b,k)) / (|a|+|b|) with four separately rounded operations;
 * left[t], right[t] = node ids merged (leaves 0..n-1, node of step t = n+t), height[t] = d(a,b)/2.
 */
void tsq_oracle_upgma(const double *packed, uint32_t n, uint32_t *left, uint32_t *right, double *height);

/*
 * Progressive multiple alignment along a guide tree (the step after SURVEY.md 8f-1; spec at the
 * definition in gotoh_oracle.c).  seqs/offs/lens as for tsq_oracle_all_pairs (submitted order); left/right
 * = the n-1 merges (leaves 0..n-1, node of merge t = n+t, e.g. from tsq_oracle_upgma).  *rows_out =
 * malloc'd n x *ncols_out bytes, row r = sequence r, symbols or 0xff for a gap: release it with
 * tsq_oracle_free.  merge_scores: n-1 profile-alignment scores or NULL.  Returns 0.
 */
int tsq_oracle_msa(const uint8_t *seqs, const uint64_t *offs, const uint32_t *lens, uint32_t n, const int8_t *mat,
                   int nsym, int go, int ge, const uint32_t *left, const uint32_t *right, uint8_t **rows_out,
                   uint32_t *ncols_out, int64_t *merge_scores);
void tsq_oracle_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
