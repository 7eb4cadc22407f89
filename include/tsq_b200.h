/*
 * tsq_b200.h -- C ABI of libtsqb200.so, the B200-native all-vs-all Gotoh distance-matrix
 * backend for tweakseq.
 *
 * This is the boundary a tweakseq maintainer binds (INTEGRATION.md shows the stub).  It
 * replaces, for the pairwise-distance stage only, the process boundary through which
 * tweakseq reaches its aligner today:
 *
 *   tweakseq/Core/AlignmentTool.h:36-71        the wrapper interface (argv builder)
 *   tweakseq/Core/ClustalO.cpp:48-52           makeCommand(): the clustalo argv
 *   tweakseq/UI/SeqEditMainWin.cpp:1654-1660   QProcess::start(exec, args)  <- replaced
 *   tweakseq/UI/SeqEditMainWin.cpp:803-812     alignmentStop(): kill()      <- cancel flag
 *   tweakseq/UI/SeqEditMainWin.cpp:822-834     stdout/stderr -> MessageWin  <- log callback
 *   tweakseq/UI/SeqEditMainWin.cpp:836-861     exit code / status           <- int status
 *
 * Conventions: plain C99 types only; every function returns an int status (0 = TSQ_OK,
 * negative = error) unless noted; no exception or abort crosses this boundary; the caller
 * owns inputs and the context handle, the library owns outputs (valid until the next
 * tsq_run/tsq_compute/tsq_set_sequences/tsq_destroy on the same context).  A context is not
 * re-entrant: use one context per thread.  There is NO CPU fallback: without an sm_100
 * device tsq_create fails with TSQ_ERR_NO_DEVICE.
 *
 * Scoring spec: global alignment, end gaps penalised, affine gaps
 * (a gap of length k costs gap_open + k*gap_extend), substitution matrix BLOSUM62 in the
 * order and with the values of tweakseq/Core/Annotations/Consensus.cpp:34-69, or a caller
 * supplied symmetric matrix.  distance(i,j) = 1 - S(i,j)/min(S(i,i),S(j,j)), 1 if that
 * minimum is <= 0.  Packed upper triangle, row-major: index(i,j) = i*n - i*(i+1)/2 + (j-i-1)
 * for i < j, over the sequences in the order they were submitted.
 */
#ifndef TSQ_B200_H
#define TSQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSQ_VERSION_MAJOR 0
#define TSQ_VERSION_MINOR 2

/* status codes */
#define TSQ_OK 0
#define TSQ_ERR_INVALID (-1)    /* bad argument / bad parameter block */
#define TSQ_ERR_NO_DEVICE (-2)  /* no CUDA device of compute capability 10.x */
#define TSQ_ERR_CUDA (-3)       /* a CUDA runtime call failed; see tsq_last_error */
#define TSQ_ERR_NOMEM (-4)      /* host or device allocation failed */
#define TSQ_ERR_CANCELLED (-5)  /* the cancel flag was raised (SeqEditMainWin.cpp:803-812) */
#define TSQ_ERR_STATE (-6)      /* call out of order (e.g. results before a run) */
#define TSQ_ERR_IO (-7)         /* FASTA / distance-matrix file could not be read/written */
#define TSQ_ERR_MATRIX (-8)     /* substitution matrix not symmetric or out of range */
#define TSQ_ERR_RANGE (-9)      /* scores would not fit the 32-bit kernels */

#define TSQ_PROTEIN 0    /* 23 symbols ARNDCQEGHILKMFPSTWYVBZX, Consensus.cpp:34-69 */
#define TSQ_NUCLEOTIDE 1 /* 5 symbols ACGTN (U->T, other letters -> N) */
#define TSQ_ALPHABET_AUTO (-1) /* tsq_run_fasta only: decided from the residues of the file, as clustalo does without
                                  --seqtype (>= 90 % of the letters in ACGTUN = nucleotide); tweakseq projects hold
                                  either kind (SequenceFile::DNA / ::Proteins, Core/SequenceFile.h) */

/* flags */
#define TSQ_FLAG_FORCE_S32 1u     /* never use the packed 16-bit kernel */
#define TSQ_FLAG_NO_DISTANCES 2u  /* skip the fp64 distance pass (scores only) */
#define TSQ_FLAG_NO_WAVE16 4u     /* long sequences: 32-bit wavefront kernel only, not the packed one */
#define TSQ_FLAG_IDENTITY 8u      /* identity-aware scoring: among the optimal alignments of a
                                     pair take the one with most identical residue pairs, report that count
                                     (tsq_identities) and make the distance 1 - identities/min(len_i, len_j),
                                     the ClustalW pairwise distance.  Scores are unchanged. */

#define TSQ_FLAG_MSA_OUT 16u       /* tsq_run_fasta only: write the multiple alignment (tsq_msa) to fout, in tree
                                     order (the guide tree goes to <fout>.dnd only with TSQ_FLAG_KEEP_TREE) */
#define TSQ_FLAG_KEEP_DISTMAT 32u /* with TSQ_FLAG_MSA_OUT: also write the distance matrix, to <fout>.distmat
                                     (n^2 numbers of text; an external aligner writes it only on request too) */

#define TSQ_FLAG_INPUT_ORDER 64u  /* with TSQ_FLAG_MSA_OUT: rows in the order of the input file instead of tree order
                                     (clustalo's --output-order=input-order) */

#define TSQ_FLAG_KEEP_TREE 128u    /* with TSQ_FLAG_MSA_OUT: also write the guide tree, to <fout>.dnd (without MSA_OUT the
                                     tree always accompanies the matrix: the two files clustalo takes) */

#define TSQ_FLAG_KIMURA 256u       /* with TSQ_FLAG_IDENTITY: the distance is Kimura-corrected, -ln(1 - D - D^2/5) with
                                     D = 1 - identities/min(len) (ClustalW's correction for multiple substitutions).  The
                                     formula holds for D < 0.75; beyond it ClustalW reads a table the reference does not
                                     hold, so a job with such a pair fails with TSQ_ERR_RANGE instead of guessing.  The
                                     logarithm is evaluated by a fixed sequence of IEEE double operations (bit-identical
                                     on CPU oracle and GPU; within 2 ulp of ln). */
#define TSQ_FLAG_SCORES_I16 512u   /* scores come back as int16 (tsq_scores16; tsq_scores is refused): half the bytes on the
                                     PCIe link and in host memory (10 GB instead of 20 GB at
                                     configs[4]).  Exact or refused: tsq_upload returns TSQ_ERR_RANGE when a score of the
                                     job could leave [-32767, 32767].  Not with TSQ_FLAG_IDENTITY, not with
                                     tsq_set_result_buffers */

typedef struct tsq_ctx tsq_ctx;

typedef struct tsq_params {
  uint32_t struct_size;  /* = sizeof(tsq_params); lets the struct grow compatibly */
  int32_t alphabet;      /* TSQ_PROTEIN | TSQ_NUCLEOTIDE */
  int32_t gap_open;      /* >= 0; <0 selects the default (protein 11, nucleotide 10) */
  int32_t gap_extend;    /* >= 0; <0 selects the default (1) */
  const int8_t *matrix;  /* nsym x nsym row-major, symmetric, or NULL = built-in */
  int32_t device;        /* CUDA ordinal of the device this context computes on */
  int32_t part_rank;     /* this context's share of the pair space: rank ... */
  int32_t part_world;    /* ... of world (1 = everything).  See tsq_partition. */
  uint32_t flags;        /* TSQ_FLAG_* */
  int32_t n_devices;     /* 0 or 1: the one device above.  2, 4, 8, ...: this context drives devices
                            device .. device + n_devices - 1 of the box from one process (one host thread and
                            one stream per device); the sorted rows are cut into n_devices slabs by the same
                            planner as part_rank/part_world (which must then be 0/1), every device finalizes
                            its own slab and copies it over its OWN PCIe link into the one pinned host result.
                            -1 = every usable device of the box (tsq_device_count).  SURVEY.md section 8b/8e. */
} tsq_params;

typedef struct tsq_stats {
  uint64_t n_sequences;
  uint64_t n_pairs;          /* pairs this context computed (its partition) */
  uint64_t cells;            /* sum len_i*len_j over those pairs */
  uint64_t cells_s16;        /* of which in the packed 16-bit inter-task kernel */
  uint64_t cells_s32;        /* of which in the 32-bit wavefront kernel */
  double kernel_ms;          /* CUDA-event time of the last tsq_compute (device only) */
  double upload_ms;          /* host encode/sort/pack + H2D of the last tsq_upload */
  double download_ms;        /* D2H of the last tsq_download */
  double gcups_kernel;       /* cells / kernel_ms */
  uint32_t launches;         /* kernels launched by the last tsq_compute */
  uint32_t sm_count;
  uint32_t strip_width;      /* K of the 16-bit kernel variant used */
  uint32_t upload_launches;  /* kernels launched by the last tsq_upload (subject database build) */
  uint64_t h2d_bytes;        /* bytes the last tsq_upload copied host -> device */
  uint64_t d2h_bytes;        /* bytes the last tsq_download copied device -> host */
  double tree_ms;            /* CUDA-event time of the last tsq_guide_tree (device only) */
  double msa_ms;             /* wall time of the last tsq_msa (plan, kernels, copies) */
  double encode_ms;          /* host time of the last tsq_set_sequences[_flat] (gap stripping, letter map, self scores) */
} tsq_stats;

/* progress in [0,1]; msg may be NULL.  Return value ignored. */
typedef void (*tsq_progress_cb)(void *user, double fraction, const char *msg);
/* one line of log text, no trailing newline (what MessageWin::addMessage would receive) */
typedef void (*tsq_log_cb)(void *user, const char *line);

int tsq_version(int *major, int *minor);
/* textual version, what AlignmentTool::version() shows (ClustalO.cpp:100-111) */
const char *tsq_version_string(void);
/* static description of a status code */
const char *tsq_status_string(int status);
/* Number of usable (compute capability 10.x) devices; <0 = status code. */
int tsq_device_count(void);

void tsq_default_params(tsq_params *p);
/*
 * Protein or nucleotide?  What clustalo decides without --seqtype, and what TSQ_ALPHABET_AUTO means in
 * tsq_run_fasta: TSQ_NUCLEOTIDE when at least 90 % of the letters are ACGTUN, else TSQ_PROTEIN (also for no
 * letters at all).  Host only; for callers that hold the residues in memory.
 */
int tsq_detect_alphabet(const char *const *residues, const uint32_t *lengths, uint32_t n);
/*
 * The encoding tsq_set_sequences applies, for callers that want the symbols themselves (and for the tests
 * that pin the vectorised encoder against the oracle): letter map of Consensus.cpp:61-69, case-insensitive;
 * '-', '.' and whitespace dropped; any other byte -> X / N.  `out` needs room for `len` bytes; *out_len
 * receives the number of symbols.  *self_score (may be NULL) = sum of S(x, x) under the alphabet's default
 * matrix.  Host only, no device.  TSQ_ERR_INVALID for a bad alphabet or null pointers.
 */
int tsq_encode(int alphabet, const char *residues, uint64_t len, uint8_t *out, uint64_t *out_len, int64_t *self_score);
int tsq_create(tsq_ctx **out, const tsq_params *params);
int tsq_destroy(tsq_ctx *ctx);
/* Last error text of this context ("" if none).  Never NULL. */
const char *tsq_last_error(const tsq_ctx *ctx);

/*
 * Residues as ASCII letters, one array per sequence, lengths[i] bytes each (no NUL needed).
 * Case-insensitive; '-', '.', and whitespace are dropped (an aligned project reaches the
 * tool with gaps still in place: Sequence.cpp:57-69); any other byte maps to X (protein)
 * or N (nucleotide).  The library copies and encodes; the caller keeps ownership.
 */
int tsq_set_sequences(tsq_ctx *ctx, const char *const *residues, const uint32_t *lengths,
                      uint32_t n);

/* Same, from ONE contiguous host buffer: sequence i is residues[offsets[i] .. offsets[i+1]). */
int tsq_set_sequences_flat(tsq_ctx *ctx, const char *residues, const uint64_t *offsets, uint32_t n);

/* The three stages of a run, separately callable (bench.py times them separately). */
int tsq_upload(tsq_ctx *ctx);   /* sort/pack on the host, H2D */
int tsq_compute(tsq_ctx *ctx);  /* enqueue all kernels on the context's stream (async) */
int tsq_download(tsq_ctx *ctx); /* synchronise, D2H of scores (+ distances) */
/*
 * Streamed results for the staged calls (tsq_run always streams): with enable != 0 a later tsq_compute lets
 * finished row ranges leave for the host while the packed kernel is still running -- SURVEY.md section 8e's
 * "slabs overlapped with remaining compute".  The kernel writes the distance next to every score and counts
 * finished tasks per row range; a side stream waits for a range's count (a stream memory operation,
 * cuStreamWaitValue32) and copies its scores (+ distances) out: one launch, no launch boundary per range.
 * (Without stream memory operations: a few launches over consecutive row ranges, each followed by finalize +
 * copy-out on the side stream.)  tsq_download then only waits.  Takes effect where the results need no un-sort
 * (fixed-length input, no identity keys, short sequences only); other jobs run as before.
 * The host destination (the library's pinned buffer, or tsq_set_result_buffers) must not change in between.
 */
int tsq_stream_results(tsq_ctx *ctx, int enable);
/* Make tsq_compute enqueue on a caller-owned cudaStream_t (passed as void*); NULL = own. */
int tsq_set_stream(tsq_ctx *ctx, void *cuda_stream);
/* Block until everything enqueued by tsq_compute has finished. */
int tsq_synchronize(tsq_ctx *ctx);

/*
 * upload + compute + download, blocking, cancellable between kernel launches:
 * *cancel != 0 makes it return TSQ_ERR_CANCELLED.  cb/cancel may be NULL.
 */
int tsq_run(tsq_ctx *ctx, tsq_progress_cb cb, void *user, volatile int *cancel);

/* Host results (after tsq_run or tsq_download). count = n*(n-1)/2. */
int tsq_scores(tsq_ctx *ctx, const int32_t **packed_upper, uint64_t *count);
/* With TSQ_FLAG_SCORES_I16: the same matrix as int16. */
int tsq_scores16(tsq_ctx *ctx, const int16_t **packed_upper, uint64_t *count);
int tsq_distances(tsq_ctx *ctx, const double **packed_upper, uint64_t *count);
int tsq_self_scores(tsq_ctx *ctx, const int32_t **self, uint32_t *n);
/* TSQ_FLAG_IDENTITY only: identical residue pairs on the chosen optimal alignment, per pair. */
int tsq_identities(tsq_ctx *ctx, const int32_t **packed_upper, uint64_t *count);

/*
 * Device results (after tsq_compute): pointers into this context's device memory, packed
 * like the host results.  For world > 1 only [*part_begin, *part_end) of the packed index
 * space -- in the library's internal length-sorted order -- is filled on this rank; the
 * host layer gathers the slabs (NCCL) into rank 0 and calls tsq_finalize there.
 */
int tsq_device_scores(tsq_ctx *ctx, void **d_sorted_scores, uint64_t *count);
int tsq_partition(tsq_ctx *ctx, uint64_t *part_begin, uint64_t *part_end);
/* The slab of ANY rank of this context's partition (every rank plans all ranks identically from the
 * same sequences, so no exchange of ranges is needed before the gather).  After tsq_upload. */
int tsq_partition_of(tsq_ctx *ctx, int32_t rank, uint64_t *begin, uint64_t *end);
/* Sorted-order slab complete on this device -> original-order scores (+distances). */
int tsq_finalize(tsq_ctx *ctx);
int tsq_device_results(tsq_ctx *ctx, void **d_scores, void **d_distances, uint64_t *count);
/*
 * The device buffer of sorted-order scores this context really holds: packed indices [*first, *first + *count).
 * The whole triangle on a single-rank context and on rank 0 of a partition whose slabs must be gathered; only the
 * rank's own slab otherwise (a rank of world 8 at 100 000 sequences holds 2.5 GB, not 20 GB).  After tsq_upload.
 */
int tsq_device_slab(tsq_ctx *ctx, void **d_sorted_scores, uint64_t *first, uint64_t *count);
/*
 * How the results of a partitioned job (part_world > 1) come together.  *sharded = 1: the length sort left the
 * submitted order unchanged (fixed-length input, no empty sequence, no identity keys), so a rank's slab is a
 * contiguous piece of the final packed triangle: every rank runs tsq_finalize/tsq_download on its own slab and
 * copies it to the host over its own PCIe link -- no gather at all (point all ranks at one shared host buffer with
 * tsq_set_result_buffers).  *sharded = 0: the un-sort scatters a slab over the triangle: gather the slabs into
 * rank 0's buffer (tsq_device_slab; NCCL send/recv) and finalize/download there.  After tsq_upload.
 */
int tsq_results_sharded(tsq_ctx *ctx, int *sharded);
/*
 * Caller-owned host memory for the results instead of the library's own pinned buffers: `count` = n*(n-1)/2
 * int32 scores and (unless NULL / TSQ_FLAG_NO_DISTANCES) as many doubles, e.g. one POSIX shared-memory segment
 * that every rank of a torchrun job maps.  The library page-locks what it writes (cudaHostRegister) on first use
 * and releases it in tsq_destroy or at the next call of this function; tsq_scores/tsq_distances then return these
 * pointers.  NULL, NULL, 0 restores the library's buffers.
 */
int tsq_set_result_buffers(tsq_ctx *ctx, int32_t *scores, double *distances, uint64_t count);

/*
 * Guide tree (UPGMA, average linkage) from the distance matrix of the last run -- the next
 * consumer of the matrix (what clustalo builds from --distmat-in).
 * n-1 merges in order; node ids: leaves 0..n-1 (submitted order), the node made by merge t is
 * n+t; height = half the distance of the merged clusters.  Ties: smallest first slot, then
 * smallest second slot; the merged cluster keeps the first slot.  Computed on the device.
 */
#define TSQ_GUIDE_TREE_MAX_N 32768u /* tsq_guide_tree / tsq_msa / tsq_write_newick beyond this: TSQ_ERR_RANGE.  The tree
                                       kernel holds a dense n x n fp64 matrix (8.6 GB at the limit) and its n-1 merges are
                                       dependent steps of one persistent CTA (~6 us each: 0.2 s at the limit).  BASELINE
                                       configs[4] (100 000 sequences) is a distance-matrix job; its matrix would be 80 GB. */
typedef struct tsq_merge {
  uint32_t left, right;
  double height;
} tsq_merge;
int tsq_guide_tree(tsq_ctx *ctx, const tsq_merge **merges, uint32_t *count);
/* Newick text of that tree ("(a:0.1,(b:0.05,c:0.05):0.05);"), branch lengths %.6f, to a file
 * (clustalo --guidetree-in).  labels[i] names leaf i; NULL = "s<i>". */
int tsq_write_newick(tsq_ctx *ctx, const char *const *labels, const char *path);

/*
 * Progressive multiple alignment along that guide tree -- what the external aligner hands back and
 * Project::readNewAlignment ingests (tweakseq/Core/Project.cpp:908-1032): equal-length gapped rows.
 * With it the backend needs no clustalo at all.  The n-1 merges are applied in order; merge t aligns
 * the alignments of its two clusters column against column with the recurrence of SURVEY 8c, the
 * score of two columns being the sum of S(a, b) over all residue pairs across them (a residue facing
 * a gap scores 0) and a gap of k columns costing |X| |Y| (gap_open + k gap_extend); ties as in
 * tsq_align_pair, so two sequences align exactly as tsq_align_pair aligns them.  Computed on the device
 * (one CTA per merge, all merges of a tree level in one launch).  Needs tsq_run (distances).
 * *rows: n x *ncols characters, row-major, row r = submitted sequence r, canonical upper-case symbols
 * and '-', no terminators; *tree_order: the n sequence indices left to right in the tree (the row
 * order of `clustalo --output-order=tree-order`, ClustalO.cpp:51).  Library-owned; any out pointer
 * may be NULL.
 */
int tsq_msa(tsq_ctx *ctx, const char **rows, uint32_t *nrows, uint32_t *ncols, const uint32_t **tree_order);
/*
 * That alignment as a FASTA file (60 columns per line).  headers[r]: header line of sequence r with or
 * without its '>' (NULL or headers == NULL: ">s<r>").  residues/lengths (both or neither): the
 * sequences as submitted; their own spelling (case, J/O/U, ...) then replaces the canonical symbols,
 * as an external aligner would echo it.  tree_order != 0: rows in tree order, else as submitted.
 */
int tsq_write_msa_fasta(tsq_ctx *ctx, const char *const *headers, const char *const *residues,
                        const uint32_t *lengths, const char *path, int tree_order);

/*
 * One optimal global alignment of sequences i and j (submitted order) WITH its path -- the
 * "emit pairwise alignments" half of SURVEY.md section 8f-2; what a user would otherwise get by
 * running the external aligner on two sequences (tweakseq/Core/ClustalO.cpp:48-52 argv on a
 * two-record FASTA).  Needs tsq_upload (or tsq_run) to have happened.  row_i / row_j receive the
 * two gapped rows: canonical upper-case symbols (ARNDCQEGHILKMFPSTWYVBZX or ACGTN; the input's
 * own spelling of a residue is not kept) and '-', NUL-terminated; capacity must be at least
 * len_i + len_j + 1 (lengths after gap stripping).  *columns = alignment length, *score = its
 * Gotoh score, equal to the matrix entry S(i, j).  Among equally good alignments the path is
 * fixed by rule (diagonal before gap-in-row-i before gap-in-row-j; a gap run is opened rather
 * than extended on a tie), the same rule the CPU oracle applies.  It is one optimal alignment: its
 * identity count can be below the maximum TSQ_FLAG_IDENTITY reports over all optimal alignments.
 */
int tsq_align_pair(tsq_ctx *ctx, uint32_t i, uint32_t j, char *row_i, char *row_j, uint32_t capacity,
                   uint32_t *columns, int32_t *score);

/*
 * Consensus annotation of an alignment: what Consensus::calculate
 * (tweakseq/Core/Annotations/Consensus.cpp:80-161) computes on the GUI thread in O(cols * rows^2),
 * here per column from a histogram of residue classes in O(rows + 24^2).  rows[r] has ncols
 * characters (what `residues[c].unicode() & 0xff` yields: only 'A'..'Z' are residues, everything
 * else -- gaps, lower case -- is the "not a residue" class 99 of Consensus.cpp:99-105).  out
 * receives ncols characters: the residue of the first row with the best BLOSUM62-weighted sum
 * against all other rows if its count of positively scoring partners reaches `plurality`, else
 * '?'.  plurality < 0 selects the reference's default, nrows / 2 (Consensus.cpp:164-175).
 */
int tsq_consensus(tsq_ctx *ctx, const char *const *rows, uint32_t nrows, uint32_t ncols,
                  double plurality, char *out);

/*
 * Host-only planning (no device needed): the packed-index slab [begins[r], ends[r]) -- in the
 * library's length-sorted order -- that rank r of `world` computes for sequences of the given
 * encoded lengths.  Same arithmetic tsq_upload uses; lets the host layer size its gather.
 */
int tsq_plan_partition(const tsq_params *params, const uint32_t *lengths, uint32_t n,
                       int32_t world, uint64_t *begins, uint64_t *ends);

int tsq_get_stats(tsq_ctx *ctx, tsq_stats *out);
/* Multi-device contexts (n_devices > 1): the figures of device number `index` (0 .. n_devices-1) alone;
 * tsq_get_stats then reports the whole job (sums; kernel_ms = the slowest device). */
int tsq_get_device_stats(tsq_ctx *ctx, int32_t index, tsq_stats *out);

/*
 * The length limits that select a kernel under this context's matrix and gap model (host arithmetic; DESIGN.md
 * section 4).  The packed 16-bit kernels are exact only while every intermediate of the recurrence stays inside an
 * unsigned 16-bit half; these are the bounds the library enforces -- tests feed the kernels inputs AT them.
 */
typedef struct tsq_limits {
  uint32_t max_len_packed;  /* longest sequence the packed 16-bit inter-task kernel takes (0: none) */
  uint32_t max_len_inter;   /* longest sequence the inter-task kernel of this context takes; longer: wavefront */
  int32_t inter_is_32bit;   /* that kernel is the 32-bit one (identity keys, or parameters too wide for 16 bits) */
  int32_t wave_packed;      /* long sequences run on the packed wavefront kernel (else the 32-bit one) */
  int64_t wave_window;      /* span, in score units, of the cells a warp of that kernel holds at one time; the kernel
                               is selected while it is <= wave_window_max */
  int64_t wave_window_max;
  int32_t delta;            /* skew per anti-diagonal, ceil(-min S / 2) */
  int32_t bias_at_limit;    /* BIAS of a job whose longest packed sequence is max_len_packed */
} tsq_limits;
int tsq_get_limits(tsq_ctx *ctx, tsq_limits *out);

/*
 * Integer-pipe issue-rate probe ("measure, don't assume"): thread-level
 * VIADDMNMX.U16x2 / VIMNMX3.U16x2 results per clock per SM on this context's device.
 */
int tsq_measure_dpx_rate(tsq_ctx *ctx, double *ops_per_clk_per_sm, double *sm_mhz);
/*
 * All three live denominators of the roofline bench.py reports (nothing hard-coded): the DPX rate above, the issue
 * ceiling (independent full-rate integer instructions per clock per SM, thread level: 128 on paper) and the packed
 * cells per clock per SM that the inner loop's own instruction mix (3 DPX + 2 adds + 1 LDS) reaches in isolation,
 * dependency-free.  Two-pipe bound of the packed kernels: min(dpx / 3, issue / 6) packed cells per clock per SM.
 */
typedef struct tsq_pipe_rates {
  double dpx_per_clk_sm;
  double issue_per_clk_sm;
  double mix_packed_cells_per_clk_sm;
  double sm_mhz;
} tsq_pipe_rates;
int tsq_measure_pipe_rates(tsq_ctx *ctx, tsq_pipe_rates *out);

/*
 * File-level convenience for the Qt adapter: read the FASTA file tweakseq
 * exported (Project.cpp:870-881, FASTAFile.cpp:149-171), compute, and write a square
 * PHYLIP-style distance matrix (n, then "label d d d ..." rows) that clustalo accepts via
 * --distmat-in, plus the guide tree as <distmat_out>.dnd.  Labels follow FASTAFile::parseComment
 * (FASTAFile.cpp:177-187).  With TSQ_FLAG_MSA_OUT in params->flags the file named by distmat_out
 * instead receives the multiple alignment (FASTA, tree order, header lines and residue spelling as
 * read) -- the file tweakseq reads back at SeqEditMainWin.cpp:836-861; the matrix is then written only
 * with TSQ_FLAG_KEEP_DISTMAT, to <distmat_out>.distmat.
 */
/* The matrix writer of tsq_run_fasta on its own (host only, no device): n, then one "label d d d ..." row per
 * sequence with %.6f distances, from the packed upper triangle (n*(n-1)/2 values; diagonal 0). */
int tsq_write_distmat(const char *path, const char *const *labels, uint32_t n, const double *packed_upper);

int tsq_run_fasta(const char *fasta_in, const char *distmat_out, const tsq_params *params,
                  tsq_log_cb log, void *user, volatile int *cancel);

#ifdef __cplusplus
}
#endif
#endif /* TSQ_B200_H */
